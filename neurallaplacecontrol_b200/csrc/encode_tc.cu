#include "common.cuh"
namespace nlc {
int launch_encode_tc(nlc_model_s*, const float*, int, int, int, float*, int, cudaStream_t) {
  set_error("tcgen05 encoder is not built in this revision");
  return NLC_ERR_UNSUPPORTED;
}
}  // namespace nlc
