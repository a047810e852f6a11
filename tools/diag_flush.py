#!/usr/bin/env python
"""End-to-end control step after an L2 flush (as bench.py times it) against the same call without the flush, and against the
device-timed step.  Measurement tool."""
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


def med(v):
    return 1e3 * sorted(v)[len(v) // 2]


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
        lib.nlc_planner_command_host(h, sp, bp, None, op, st)
        planner.step()
    torch.cuda.synchronize()
    res = {}
    for flush in (False, True):
        for name, fn in (("C command_host", lambda: lib.nlc_planner_command_host(h, sp, bp, None, op, st)),
                         ("py command", lambda: planner.command(inp["state"], inp["buffer"]))):
            ts = []
            for _ in range(20):
                if flush:
                    ctx.flush.fill_(1)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            res[(name, flush)] = med(ts)
        ev = []
        for _ in range(20):
            if flush:
                ctx.flush.fill_(1)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); planner.step(); e.record(); e.synchronize()
            ev.append(s.elapsed_time(e) * 1e-3)
        res[("device step (events)", flush)] = med(ev)
    for k, v in res.items():
        print(f"{env}  {k[0]:24s} flush={k[1]!s:5s}  {v:.4f} ms")


if __name__ == "__main__":
    main()
