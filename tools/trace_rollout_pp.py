#!/usr/bin/env python
"""Timeline of the ping-pong rollout kernel's first CTA (clock64 per warp at the phase boundaries).  Measurement tool only.
    NLC_ROLLOUT_TILES=3 python tools/trace_rollout_pp.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurallaplacecontrol_b200 as nlc  # noqa: E402
from neurallaplacecontrol_b200 import _lib  # noqa: E402
from _util import DT, S_TERMS, weights  # noqa: E402
from oracle import costs  # noqa: E402

ORDER = [(11, "step start"), (0, "E1x go"), (1, "E1y go"), (12, "E1y end"), (2, "E2x go"), (3, "E2y go"), (13, "E2y end"), (4, "E3ax go"),
         (5, "E3ay go"), (14, "E3ay end"), (6, "E3bx go"), (7, "E3by go"), (15, "E3by end"), (8, "U go"), (9, "U end"), (10, "step end")]


def main():
    os.environ["NLC_ROLLOUT_TILES"] = "3"
    torch.set_grad_enabled(False)
    dev = torch.device("cuda", 0)
    env = os.environ.get("ENV", "oderl-acrobot")
    nx, nu = costs.ENV_DIMS[env]
    model = nlc.NeuralLaplaceModel(nx, nu, nx, hidden_units=128, s_recon_terms=S_TERMS, state_mean=np.zeros(nx), state_std=np.ones(nx),
                                   action_mean=np.array([0] * nu), action_std=np.array([1.0]), normalize=True, normalize_time=True,
                                   dt=DT, device=dev).double()
    model.load_state_dict(weights(env, calibrated=True))
    mh = model.set_prediction_time(DT)
    lib = _lib.load()
    lib.nlc_debug_set_rollout_trace.argtypes = [C.c_void_p]
    lib.nlc_debug_set_rollout_trace.restype = None
    K, H, B = int(os.environ.get("K", 65536)), 50, 4
    hist = torch.randn(K, B - 1 + H, nu, device=dev, dtype=torch.float32)
    p = torch.empty(K, H, 2, device=dev, dtype=torch.float32)
    mode = _lib.MATH_MODES["tc_split3"]
    _lib.check(lib.nlc_encode_history(mh, hist.data_ptr(), K, H, B, p.data_ptr(), mode, _lib.current_stream_ptr()))
    st = torch.tensor(costs_start(env), dtype=torch.float32, device=dev)
    ro = _lib.RolloutOpts()
    ro.env, ro.state_constraint, ro.goal_x, ro.dynamics, ro.delay, ro.dt = _lib.ENV_IDS[env], 0, 0.0, 0, 0, DT
    cost = torch.empty(K, device=dev, dtype=torch.float32)
    states = torch.empty(K, H, nx, device=dev, dtype=torch.float32)
    trace = torch.zeros(2 * 52 * 16 * 16 + 4 * 148, dtype=torch.int64, device=dev)
    lib.nlc_debug_set_rollout_trace(trace.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(2):
        if os.environ.get("FLUSH"):
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.nlc_rollout_cost(mh, C.byref(ro), st.data_ptr(), 0, p.data_ptr(), hist.data_ptr(), None, K, H, B, nu,
                                        cost.data_ptr(), states.data_ptr(), mode, _lib.current_stream_ptr()))
        e1.record()
    torch.cuda.synchronize()
    ev_ms = e0.elapsed_time(e1)
    lib.nlc_debug_set_rollout_trace(None)
    allv = trace.cpu().numpy().astype(np.int64)
    both = allv[:2 * 52 * 256].reshape(2, 52, 16, 16)
    spans = allv[2 * 52 * 256:].reshape(148, 4)
    act = spans[spans[:, 0] > 100000]
    ns = act[:, 3].max() - act[:, 2].min()
    print(f"globaltimer: first CTA start -> last CTA end {ns / 1e3:.1f} us; CUDA-event time of the traced launch {ev_ms * 1e3:.1f} us; "
          f"slowest CTA {act[:, 0].max() / 1e3:.0f} k clk in {(act[:, 3] - act[:, 2]).max() / 1e3:.1f} us = {act[:, 0].max() / (act[:, 3] - act[:, 2]).max():.3f} GHz; CTA start spread {(act[:, 2].max() - act[:, 2].min()) / 1e3:.1f} us")
    print("per-CTA spans (k clk) by CTA index:", [int(v // 1000) for v in spans[:, 0]])
    print("per-CTA spans (us) by CTA index:", [int((spans[i, 3] - spans[i, 2]) // 1000) for i in range(148)])
    print("per-CTA SM id:", [int(v) for v in spans[:, 1]])
    print(f"per-CTA kernel spans (k clk): min {act[:, 0].min() / 1e3:.0f} median {np.median(act[:, 0]) / 1e3:.0f} max {act[:, 0].max() / 1e3:.0f}; "
          f"slowest CTAs (cta, sm, k clk): {[(int(i), int(spans[i, 1]), int(spans[i, 0] // 1000)) for i in np.argsort(-spans[:, 0])[:6]]}")
    tr = both[0]
    k0, k1, k2 = tr[51, :, 0].min(), tr[51, :, 1].max(), tr[51, :, 2].max()
    print(f"CTA 0: kernel start -> epilogue warps running {k1 - k0} clk; -> step 0 start {tr[0, :, 11].min() - k0}; -> last step end "
          f"{tr[49, :, 10].max() - k0}; -> epilogue warps done {k2 - k0}")
    mm = both[1].reshape(52, 256)  # [MMA-warp step][2 * product + {start, end}]
    print("per event: min / median / max over the 16 warps, clocks relative to the step's first warp; then per-cg medians")
    for s in (5, 6, 20, 21):
        t0 = tr[s, :, 11].min()
        print(f"step {s}: length {tr[s + 1, :, 11].min() - t0}")
        for e, name in ORDER:
            v = tr[s, :, e] - t0
            cgm = [int(np.median(v[4 * c:4 * c + 4])) for c in range(4)]
            print(f"   {name:10s} {v.min():6d} / {int(np.median(v)):6d} / {v.max():6d}    cg medians {cgm}")
        # MMA warp: products of step s are (M2x M2y M3ax M3ay M3bx M3by M1x M1y) = its step s + 1 (step 0 there = the prologue's M1 pair)
        names = ["M2x", "M2y", "M3ax", "M3ay", "M3bx", "M3by", "M1x'", "M1y'"]
        print("   MMA warp (issue start -> issue end): " + "  ".join(f"{n} {mm[s + 1, 2 * j] - t0}->{mm[s + 1, 2 * j + 1] - t0}" for j, n in enumerate(names)))
    starts = [int(tr[s, :, 11].min()) for s in range(50)]
    print("step lengths (k clk):", [round((starts[i + 1] - starts[i]) / 1e3, 1) for i in range(49)])


def costs_start(env):
    return {"oderl-pendulum": [-1.0, 0.0, 1.0], "oderl-cartpole": [0.0, 0.0, -1.0, 0.0, 0.0], "oderl-acrobot": [1.0, 0.0, 1.0, 0.0, 0.0, 0.0]}[env]


if __name__ == "__main__":
    main()
