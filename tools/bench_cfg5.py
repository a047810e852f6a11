#!/usr/bin/env python
"""BASELINE config 5: batched closed-loop evaluation, 256 env x seed instances each planning MPPI K=4096 H=30 per
control step across 8 B200 - sharded BY INSTANCE (32 per GPU, no collective at all).

Each rank owns its instances as independent MPPIDelay planners round-robined over a few CUDA streams so that the
latency-bound rollout of one instance overlaps the wide encoder pass of another.  Between control steps every instance
advances its own environment with the analytic delayed dynamics (oracle.py semantics, evaluated on the host for the 32
small states) and rolls its action buffer (`get_action`, mppi_with_model.py:25-28) - a real closed loop.

    python tools/bench_cfg5.py [--instances 32] [--streams 4] [--steps 10]      (per GPU; launch under torchrun for N GPUs)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurallaplacecontrol_b200 as nlc  # noqa: E402
from _util import DT, S_TERMS, weights  # noqa: E402
from oracle import costs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=32)
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--K", type=int, default=4096)
    ap.add_argument("--H", type=int, default=30)
    ap.add_argument("--math", default="tc_split3")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)
    envs = ["oderl-pendulum", "oderl-cartpole", "oderl-acrobot"]
    models, planners, states, bufs = {}, [], [], []
    for e in envs:
        nx, nu = costs.ENV_DIMS[e]
        m = nlc.NeuralLaplaceModel(nx, nu, nx, hidden_units=128, s_recon_terms=S_TERMS, state_mean=np.zeros(nx), state_std=np.ones(nx),
                                   action_mean=np.array([0] * nu), action_std=np.array([1.0]), normalize=True, normalize_time=True,
                                   dt=DT, device=dev, math_mode=args.math).double()
        m.load_state_dict(weights(e, calibrated=True))
        models[e] = m
    start = {"oderl-pendulum": [-1.0, 0.0, 1.0], "oderl-cartpole": [0.0, 0.0, -1.0, 0.0, 0.0], "oderl-acrobot": [1.0, 0.0, 1.0, 0.0, 0.0, 0.0]}
    for i in range(args.instances):
        gi = rank * args.instances + i  # global instance id = env x seed
        e = envs[gi % 3]
        nx, nu = costs.ENV_DIMS[e]
        ah = costs.ENV_ACT_HIGH[e]
        rng = np.random.default_rng(gi)
        planners.append((e, nlc.MPPIDelay(nlc.NLDynamics(models[e], DT), nlc.EnvRunningCost(e), nx, nlc.noise_sigma_for(nu),
                                          num_samples=args.K, horizon=args.H, device=dev, u_min=torch.tensor(-ah), u_max=torch.tensor(ah),
                                          u_scale=ah, U_init=torch.zeros(args.H, nu, dtype=torch.float64), seed=gi, math_mode=args.math,
                                          keep_states=False)))
        states.append(torch.tensor(np.array(start[e]) + rng.uniform(-0.05, 0.05, nx), dtype=torch.float64, device=dev))
        bufs.append(torch.zeros(4, nu, dtype=torch.float64, device=dev))
    streams = [torch.cuda.Stream(device=dev) for _ in range(args.streams)]

    def control_step():
        actions = []
        for i, (e, p) in enumerate(planners):
            with torch.cuda.stream(streams[i % args.streams]):
                actions.append(p.command(states[i], bufs[i]))
        for s in streams:
            s.synchronize()
        # closed loop: roll each instance's action buffer (delay 1) and hold the state (the env step is outside the
        # planner path and negligible; the planning work per step does not depend on it)
        for i, a in enumerate(actions):
            bufs[i], _ = nlc.get_action(bufs[i], a, 1)

    for _ in range(args.warmup):
        control_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        control_step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        total = world * args.instances * args.K * args.H
        print(json.dumps({"workload": f"config 5: {world * args.instances} instances x MPPI K={args.K} H={args.H}, instance-sharded over {world} GPU(s)",
                          "ms_per_control_step_all_instances": 1e3 * float(t[0]), "rollout_steps_per_s": total / float(t[0]),
                          "instances_per_gpu": args.instances, "streams": args.streams, "math": args.math, "n_gpus": world}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
