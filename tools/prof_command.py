#!/usr/bin/env python
"""cProfile of MPPIDelay.command with host buffers (the e2e path) at a small plan: where the Python side spends its time.
Measurement tool.   python tools/prof_command.py [cfg3]"""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


class A:
    gpus, math = 1, "tc_split3"


def main():
    torch.set_grad_enabled(False)
    ctx = bench.Ctx(A)
    env, K, H, _ = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
    inp, model, planner = bench.make_planner(ctx, env, K, H)
    for _ in range(20):
        planner.command(inp["state"], inp["buffer"])
    torch.cuda.synchronize()
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n):
        planner.command(inp["state"], inp["buffer"])
    torch.cuda.synchronize()
    print("wall us per command:", (time.perf_counter() - t0) / n * 1e6)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        planner.command(inp["state"], inp["buffer"])
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(18)


if __name__ == "__main__":
    main()
