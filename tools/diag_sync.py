#!/usr/bin/env python
"""Launch / synchronisation overhead of one control step: graph replay with and without a sync per step, direct launches
(NLC_NO_GRAPH=1 in the environment), host entry point.  Measurement tool."""
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
    planner.set_inputs(torch.tensor(inp["state"], device=ctx.dev), inp["buffer"].to(ctx.dev))
    lib, h, st = ctx.lib, planner._handle, _lib.current_stream_ptr()
    sp, k1 = _lib.as_double_array(inp["state"])
    bp, k2 = _lib.as_double_array(inp["buffer"].numpy())
    out = np.empty(inp["nu"], dtype=np.float64)
    op = out.ctypes.data_as(C.POINTER(C.c_double))
    for _ in range(5):
        lib.nlc_planner_step(h, st)
        lib.nlc_planner_command_host(h, sp, bp, None, op, st)
    torch.cuda.synchronize()
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        lib.nlc_planner_step(h, st)
    torch.cuda.synchronize()
    back = (time.perf_counter() - t0) / n * 1e3
    t0 = time.perf_counter()
    for _ in range(n):
        lib.nlc_planner_step(h, st)
        torch.cuda.synchronize()
    synced = (time.perf_counter() - t0) / n * 1e3
    t0 = time.perf_counter()
    for _ in range(n):
        lib.nlc_planner_command_host(h, sp, bp, None, op, st)
    host = (time.perf_counter() - t0) / n * 1e3
    t0 = time.perf_counter()
    for _ in range(n):
        planner.command(inp["state"], inp["buffer"])
    py = (time.perf_counter() - t0) / n * 1e3
    print(f"{env} K={K} H={H} graph={'off' if os.environ.get('NLC_NO_GRAPH') == '1' else 'on'}: back-to-back {back:.4f}  step+sync {synced:.4f}  "
          f"command_host {host:.4f}  MPPIDelay.command(host) {py:.4f} ms")


if __name__ == "__main__" and os.environ.get("NLC_DIAG_BREAKDOWN") != "1":
    main()


def breakdown():
    """Where MPPIDelay.command(host buffers) spends its host time: python before the C call, the C call, python after."""
    torch.set_grad_enabled(False)
    ctx = bench.Ctx(A)
    env, K, H, _ = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
    inp, model, planner = bench.make_planner(ctx, env, K, H)
    lib = planner._lib
    real = lib.nlc_planner_command_host
    marks = {}

    class Wrap:
        def __call__(self, *a):
            marks["c0"] = time.perf_counter()
            r = real(*a)
            marks["c1"] = time.perf_counter()
            return r

    class LibProxy:
        def __getattr__(self, n):
            return Wrap() if n == "nlc_planner_command_host" else getattr(lib, n)

    planner._lib = LibProxy()
    for _ in range(5):
        planner.command(inp["state"], inp["buffer"])
    pre, cc, post = [], [], []
    for _ in range(30):
        t0 = time.perf_counter()
        planner.command(inp["state"], inp["buffer"])
        t1 = time.perf_counter()
        pre.append(marks["c0"] - t0); cc.append(marks["c1"] - marks["c0"]); post.append(t1 - marks["c1"])
    med = lambda v: 1e3 * sorted(v)[len(v) // 2]
    print(f"{env} K={K}: python before C call {med(pre):.4f}  C call {med(cc):.4f}  python after {med(post):.4f} ms")


if __name__ == "__main__" and os.environ.get("NLC_DIAG_BREAKDOWN") == "1":
    breakdown()
