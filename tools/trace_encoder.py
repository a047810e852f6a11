#!/usr/bin/env python
"""Timeline of the history encoder's first CTA (clock64 per warp at the phase boundaries), to see which waits are
exposed.  Measurement tool only.   python tools/trace_encoder.py [--math tc_split3]"""
import argparse
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

EV = ["eA start", "bar_a ok", "eA math done", "bar_b ok", "h0 published(+A issue)", "eB math done", "h1 published(+B issue)"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--math", default="tc_split3")
    args = ap.parse_args()
    torch.set_grad_enabled(False)
    dev = torch.device("cuda", 0)
    env = "oderl-acrobot"
    nx, nu = costs.ENV_DIMS[env]
    model = nlc.NeuralLaplaceModel(nx, nu, nx, hidden_units=128, s_recon_terms=S_TERMS, state_mean=np.zeros(nx), state_std=np.ones(nx),
                                   action_mean=np.array([0] * nu), action_std=np.array([1.0]), normalize=True, normalize_time=True,
                                   dt=DT, device=dev).double()
    model.load_state_dict(weights(env, calibrated=True))
    mh = model.set_prediction_time(DT)
    lib = _lib.load()
    lib.nlc_debug_set_encoder_trace.argtypes = [C.c_void_p]
    lib.nlc_debug_set_encoder_trace.restype = None
    K, H, B = 65536, 50, 4
    hist = torch.randn(K, B - 1 + H, nu, device=dev, dtype=torch.float32)
    p = torch.empty(K, H, 2, device=dev, dtype=torch.float32)
    trace = torch.zeros(32 * 16 * 8, dtype=torch.int64, device=dev)
    lib.nlc_debug_set_encoder_trace(trace.data_ptr())
    for _ in range(3):
        _lib.check(lib.nlc_encode_history(mh, hist.data_ptr(), K, H, B, p.data_ptr(), _lib.MATH_MODES[args.math], _lib.current_stream_ptr()))
    torch.cuda.synchronize()
    tr = trace.cpu().numpy().reshape(32, 16, 8).astype(np.int64)
    t0 = tr[4, :, 0][tr[4, :, 0] > 0].min()
    print("step: per event  min / median / max over the 16 warps, clocks relative to step 4's first warp;  (dur = step length)")
    prev = None
    for s in range(4, 20):
        row = []
        for e in range(7):
            v = tr[s, :, e]
            v = v[v > 0]
            if len(v) == 0:
                row.append("      -      ")
                continue
            row.append(f"{v.min() - t0:6d}/{int(np.median(v)) - t0:6d}/{v.max() - t0:6d}")
        start = tr[s, :, 0][tr[s, :, 0] > 0]
        start = start.min() if len(start) else tr[s, :, 2][tr[s, :, 2] > 0].min()
        print(f"step {s:2d} (st={s % 4}) dur={'' if prev is None else start - prev:>6} | " + " | ".join(row))
        prev = start
    print("events:", ", ".join(f"{i}={n}" for i, n in enumerate(EV)))


if __name__ == "__main__":
    main()
