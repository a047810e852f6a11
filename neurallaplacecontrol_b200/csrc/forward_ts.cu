#include "common.cuh"
extern "C" int nlc_model_forward_ts(nlc_model_t, const float*, const float*, const float*, int, int, float*, void*) {
  nlc::set_error("per-sample prediction times are not built in this revision");
  return NLC_ERR_UNSUPPORTED;
}
