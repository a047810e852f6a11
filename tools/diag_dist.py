#!/usr/bin/env python
"""Distribution of per-call wall times: C entry point vs MPPIDelay.command with host buffers.  Measurement tool."""
import ctypes as C
import gc
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from neurallaplacecontrol_b200 import _lib  # noqa: E402


class A:
    gpus, math = 1, "tc_split3"


def stats(v):
    v = sorted(v)
    return f"min {1e3 * v[0]:.4f} med {1e3 * v[len(v) // 2]:.4f} mean {1e3 * sum(v) / len(v):.4f} p90 {1e3 * v[int(0.9 * len(v))]:.4f} max {1e3 * v[-1]:.4f}"


def main():
    torch.set_grad_enabled(False)
    ctx = bench.Ctx(A)
    env, K, H, _ = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
    inp, model, planner = bench.make_planner(ctx, env, K, H)
    lib = ctx.lib
    for _ in range(5):
        planner.command(inp["state"], inp["buffer"])
    h, st = planner._handle, _lib.current_stream_ptr()
    sp, k1 = _lib.as_double_array(inp["state"])
    bp, k2 = _lib.as_double_array(inp["buffer"].numpy())
    out = np.empty(inp["nu"], dtype=np.float64)
    op = out.ctypes.data_as(C.POINTER(C.c_double))
    for name, fn in (("command_host (C)", lambda: lib.nlc_planner_command_host(h, sp, bp, None, op, st)),
                     ("MPPIDelay.command", lambda: planner.command(inp["state"], inp["buffer"])),
                     ("MPPIDelay.command, gc off", lambda: planner.command(inp["state"], inp["buffer"]))):
        if "gc off" in name:
            gc.disable()
        ts = []
        for _ in range(40):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        print(f"{env} {name}: {stats(ts)}")


if __name__ == "__main__":
    main()
