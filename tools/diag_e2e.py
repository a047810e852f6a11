#!/usr/bin/env python
"""Where the end-to-end control step (host buffers) spends its time beyond the device-timed step.  Measurement tool."""
import ctypes as C
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


def main():
    torch.set_grad_enabled(False)
    ctx = bench.Ctx(A)
    env, K, H, _ = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
    inp, model, planner = bench.make_planner(ctx, env, K, H)
    state_dev = torch.tensor(inp["state"], dtype=torch.float64, device=ctx.dev)
    buf_dev = inp["buffer"].to(ctx.dev)
    lib = ctx.lib

    def wall(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return 1e3 * sorted(ts)[len(ts) // 2]

    print("command(device tensors) wall ms", wall(lambda: planner.command(state_dev, buf_dev)))
    print("command(host buffers)   wall ms", wall(lambda: planner.command(inp["state"], inp["buffer"]).cpu()))
    h = planner._handle
    sp, k1 = _lib.as_double_array(inp["state"])
    bp, k2 = _lib.as_double_array(inp["buffer"].numpy())
    out = np.empty(inp["nu"], dtype=np.float64)
    op = out.ctypes.data_as(C.POINTER(C.c_double))
    st = _lib.current_stream_ptr()
    print("nlc_planner_command_host wall ms", wall(lambda: lib.nlc_planner_command_host(h, sp, bp, None, op, st)))
    print("nlc_planner_step         wall ms", wall(lambda: lib.nlc_planner_step(h, st)))
    # the same measurements again in the opposite order: a difference between the rounds is the box (power cap, clocks), not the path
    print("nlc_planner_command_host wall ms", wall(lambda: lib.nlc_planner_command_host(h, sp, bp, None, op, st)))
    print("command(host buffers)   wall ms", wall(lambda: planner.command(inp["state"], inp["buffer"]).cpu()))
    print("command(device tensors) wall ms", wall(lambda: planner.command(state_dev, buf_dev)))
    time.sleep(2.0)
    print("after 2 s idle: nlc_planner_step wall ms", wall(lambda: lib.nlc_planner_step(h, st)))
    time.sleep(2.0)
    print("after 2 s idle: command(host)    wall ms", wall(lambda: planner.command(inp["state"], inp["buffer"]).cpu()))
    t0 = time.perf_counter()
    for _ in range(200):
        planner._ensure(4)
    print("_ensure us", (time.perf_counter() - t0) / 200 * 1e6)


if __name__ == "__main__":
    main()
