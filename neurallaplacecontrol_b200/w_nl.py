"""``NeuralLaplaceModel`` with the reference's constructor, ``state_dict`` layout and ``forward`` signature
(``w_nl.py:66-145``), evaluated by the CUDA kernels of ``libnlc_b200.so``.

The ``torch.nn`` sub-modules below only HOLD parameters under the reference's key names
(``action_encoder.gru.*``, ``action_encoder.linear_out.*``, ``laplace_rep_func.linear_tanh_stack.{0,2,4}.*``),
so reference checkpoints load unchanged and random init under a seed matches the reference module's.
They are never called: ``forward`` packs the weights into a device handle and launches the kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib


class _EncoderParams(nn.Module):
    """Parameter container for ReverseGRUEncoder (w_nl.py:14-23)."""

    def __init__(self, dimension_in, latent_dim, hidden_units, encode_obs_time):
        super().__init__()
        if encode_obs_time:
            dimension_in += 1
        self.gru = nn.GRU(dimension_in, hidden_units, 2, batch_first=True)
        self.linear_out = nn.Linear(hidden_units, latent_dim)
        nn.init.xavier_uniform_(self.linear_out.weight)


class _RepFuncParams(nn.Module):
    """Parameter container for LaplaceRepresentationFunc (w_nl.py:35-53)."""

    def __init__(self, s_dim, output_dim, latent_dim, hidden_units):
        super().__init__()
        self.linear_tanh_stack = nn.Sequential(
            nn.Linear(s_dim * 2 + latent_dim, hidden_units), nn.Tanh(),
            nn.Linear(hidden_units, hidden_units), nn.Tanh(),
            nn.Linear(hidden_units, s_dim * 2 * output_dim),
        )
        for m in self.linear_tanh_stack.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)


class NeuralLaplaceModel(nn.Module):
    def __init__(self, state_dim, action_dim, latent_dim, hidden_units=64, s_recon_terms=33, ilt_algorithm="fourier",
                 encode_obs_time=False, state_mean=None, state_std=None, action_mean=None, action_std=None,
                 normalize=False, normalize_time=False, dt=0.05, device=None, math_mode="tc_split3"):
        super().__init__()
        if ilt_algorithm != "fourier":
            raise NotImplementedError("only ilt_algorithm='fourier' is on this path (w_nl.py:86-88 'cme' is not)")
        if hidden_units not in (64, 128):
            raise NotImplementedError("the kernels are built for hidden_units = 128 (config.py:37) and 64 (the class default, "
                                      "w_nl.py:71; fp32 CUDA-core kernels only)")
        self.ilt_algorithm = ilt_algorithm
        self.latent_dim = latent_dim
        self.action_encoder = _EncoderParams(action_dim, 2, hidden_units // 2, encode_obs_time)
        self.laplace_rep_func = _RepFuncParams(s_recon_terms, state_dim, state_dim + 2, hidden_units)
        self.encode_obs_time = encode_obs_time
        self.output_dim = state_dim
        self.action_dim = action_dim
        self.hidden_units = hidden_units
        self.normalize = normalize
        self.normalize_time = normalize_time
        self.s_recon_terms = s_recon_terms
        self.math_mode = math_mode
        for name, val in (("state_mean", state_mean), ("state_std", state_std), ("action_mean", action_mean),
                          ("action_std", action_std), ("dt", dt)):
            # torch.tensor(...) exactly as w_nl.py:107-111 (dt becomes an fp32 buffer there; .double() keeps that value)
            self.register_buffer(name, torch.tensor(val if val is not None else 0.0))
        self._cuda_device = torch.device(device) if device is not None else None
        self._handle = None
        self._handle_key = None
        self._handle_ts = None

    # ---- device handle -------------------------------------------------------------------------------------------
    def _fingerprint(self):
        """(version, address) of every parameter and buffer: changes on ``load_state_dict``, optimizer steps, ``.double()`` ...
        The tensor list is cached (walking the module tree costs ~60 us, this runs once per control step) and dropped whenever
        a tensor attribute is (re)assigned or ``_apply`` (``.to`` / ``.double`` / ``.cuda``) runs."""
        ts = self.__dict__.get("_fp_tensors")
        if ts is None:
            ts = [v for _, v in self.state_dict(keep_vars=True).items()]
            self.__dict__["_fp_tensors"] = ts
        return tuple((v._version, v.data_ptr()) for v in ts)

    def _apply(self, fn, *a, **k):
        self.__dict__["_fp_tensors"] = None
        for m in self.modules():
            m.__dict__["_fp_tensors"] = None
        return super()._apply(fn, *a, **k)

    def __setattr__(self, name, value):
        if isinstance(value, (torch.Tensor, nn.Module)):
            self.__dict__["_fp_tensors"] = None
        super().__setattr__(name, value)

    def _device(self):
        if self._cuda_device is not None:
            return self._cuda_device
        p = next(self.parameters())
        if p.is_cuda:
            return p.device
        if not torch.cuda.is_available():
            raise RuntimeError("neurallaplacecontrol_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    def handle(self):
        """The packed device model (rebuilt when any parameter or buffer changed)."""
        key = (self._fingerprint(), str(self._device()))
        if self._handle is not None and key == self._handle_key:
            return self._handle
        self._free()
        lib = _lib.load()
        sd = {k: v.detach().to("cpu", torch.float64).contiguous().numpy() for k, v in self.state_dict().items()}
        keep = []

        def dp(name):
            a = np.ascontiguousarray(sd[name].reshape(-1))
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_double))

        d = _lib.ModelDesc()
        d.state_dim, d.action_dim, d.hidden_units, d.s_terms = self.output_dim, self.action_dim, self.hidden_units, self.s_recon_terms
        d.encode_obs_time, d.normalize, d.normalize_time = int(self.encode_obs_time), int(self.normalize), int(self.normalize_time)
        d.action_std_len = int(sd["action_std"].size)
        d.dt = float(sd["dt"])
        d.state_mean, d.state_std = dp("state_mean"), dp("state_std")
        d.action_mean, d.action_std = dp("action_mean"), dp("action_std")
        g = "action_encoder.gru."
        d.gru_w_ih_l0, d.gru_w_hh_l0, d.gru_b_ih_l0, d.gru_b_hh_l0 = dp(g + "weight_ih_l0"), dp(g + "weight_hh_l0"), dp(g + "bias_ih_l0"), dp(g + "bias_hh_l0")
        d.gru_w_ih_l1, d.gru_w_hh_l1, d.gru_b_ih_l1, d.gru_b_hh_l1 = dp(g + "weight_ih_l1"), dp(g + "weight_hh_l1"), dp(g + "bias_ih_l1"), dp(g + "bias_hh_l1")
        d.enc_out_w, d.enc_out_b = dp("action_encoder.linear_out.weight"), dp("action_encoder.linear_out.bias")
        s = "laplace_rep_func.linear_tanh_stack."
        d.mlp_w0, d.mlp_b0, d.mlp_w2, d.mlp_b2, d.mlp_w4, d.mlp_b4 = (dp(s + "0.weight"), dp(s + "0.bias"), dp(s + "2.weight"),
                                                                      dp(s + "2.bias"), dp(s + "4.weight"), dp(s + "4.bias"))
        if sd["action_mean"].size != self.action_dim:
            raise ValueError("action_mean must have action_dim entries")
        h = C.c_void_p()
        dev = self._device()
        _lib.check(lib.nlc_model_create(C.byref(h), C.byref(d), dev.index if dev.index is not None else 0), "nlc_model_create")
        self._handle, self._handle_key, self._handle_ts = h, key, float(sd["dt"])
        return h

    def _free(self):
        if self._handle is not None:
            _lib.load().nlc_model_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def set_prediction_time(self, ts: float):
        """Fold the constants of a fixed prediction time (the planner passes dt, mppi_with_model.py:74)."""
        h = self.handle()
        if self._handle_ts != float(ts):
            _lib.check(_lib.load().nlc_model_set_prediction_time(h, float(ts)), "nlc_model_set_prediction_time")
            self._handle_ts = float(ts)
        return h

    # ---- forward -------------------------------------------------------------------------------------------------
    def forward(self, in_batch_obs, in_batch_action, ts_pred):
        """Predicted state difference, ``w_nl.py:117-145``.  ``in_batch_obs`` (K,nx), ``in_batch_action`` (K,B,nu)
        or (K,nu), ``ts_pred`` (K,1) seconds.  CUDA tensors in, tensor of the input dtype out."""
        dev = self._device()
        out_dtype = in_batch_obs.dtype
        obs = in_batch_obs.to(device=dev, dtype=torch.float32).contiguous()
        act = in_batch_action.to(device=dev, dtype=torch.float32)
        if act.dim() == 2:
            act = act.unsqueeze(1)
        act = act.contiguous()
        K, B = act.shape[0], act.shape[1]
        ts = torch.as_tensor(ts_pred).to(device=dev, dtype=torch.float64).reshape(-1)
        if ts.numel() not in (1, K):
            raise ValueError("ts_pred must hold one time per sample")
        lib = _lib.load()
        out = torch.empty((K, self.output_dim), dtype=torch.float32, device=dev)
        t0 = float(ts[0])
        with torch.cuda.device(dev):
            if bool((ts == ts[0]).all()):
                h = self.set_prediction_time(t0)
                p_action = torch.empty((K, 2), dtype=torch.float32, device=dev)
                _lib.check(lib.nlc_model_forward(h, obs.data_ptr(), act.data_ptr(), K, B, out.data_ptr(), p_action.data_ptr(),
                                                 _lib.MATH_MODES[self.math_mode], _lib.current_stream_ptr()), "nlc_model_forward")
                self.last_p_action = p_action
            else:
                h = self.handle()
                ts32 = ts.to(torch.float32).contiguous()
                scratch = torch.empty((K, 4 + self.hidden_units), dtype=torch.float32, device=dev)
                _lib.check(lib.nlc_model_forward_ts(h, obs.data_ptr(), act.data_ptr(), ts32.data_ptr(), K, B, out.data_ptr(),
                                                    scratch.data_ptr(), _lib.MATH_MODES[self.math_mode], _lib.current_stream_ptr()),
                           "nlc_model_forward_ts")
                self.last_p_action = scratch.view(-1)[:2 * K].view(K, 2)
        return torch.squeeze(out.to(out_dtype))
