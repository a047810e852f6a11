#!/usr/bin/env python
"""One shard of the 8-way strong-scaled config 4 (acrobot K=8192 H=50) on one GPU: device-timed control step and the in-step
split, with the encoder beside the rollout (default) or, under NLC_NO_OVERLAP=1, in sequence.  Measurement tool."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


class A:
    gpus, math, steps = 1, "tc_split3", 10


def main():
    torch.set_grad_enabled(False)
    ctx = bench.Ctx(A)
    env = sys.argv[1] if len(sys.argv) > 1 else "oderl-acrobot"
    K, H = int(sys.argv[2]) if len(sys.argv) > 2 else 8192, int(sys.argv[3]) if len(sys.argv) > 3 else 50
    r = bench.time_plan(ctx, env, K, H, None, 20, 5)
    print("overlap off" if os.environ.get("NLC_NO_OVERLAP") == "1" else "overlap on", env, K, H, "step ms", round(r["ms_dev"], 4),
          "e2e ms", round(r["ms_e2e"], 4), "launches/step", r["launches"] / 20)
    print("  in-step:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in bench.in_step_split(ctx, r, 10).items()})
    print("  overlap status", r["planner"].overlap_status())


if __name__ == "__main__":
    main()
