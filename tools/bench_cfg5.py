#!/usr/bin/env python
"""BASELINE config 5: batched closed-loop evaluation, 256 env x seed instances each planning MPPI K=4096 H=30 per
control step across 8 B200 - sharded BY INSTANCE (32 per GPU, no collective at all).

Default (--mode batched): the instances of each environment on a rank form one `BatchedMPPIDelay` - one encoder launch
and one rollout launch over all their samples per control step - and advance through `env_step` on the device
(`run_closed_loop`'s body: a real closed loop, nothing crosses PCIe inside it).  The three environments' batches run on
three CUDA streams.  --mode streams is the earlier form: one `MPPIDelay` per instance round-robined over a few streams.

    python tools/bench_cfg5.py [--instances 32] [--steps 10] [--mode batched|streams]     (per GPU; torchrun for N GPUs)
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
    ap.add_argument("--mode", default="batched", choices=["batched", "streams"])
    ap.add_argument("--delay", type=int, default=1)
    ap.add_argument("--hold-state", action="store_true", help="skip env_step (open loop): isolates the planning cost")
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
    per_env = {e: [] for e in envs}
    for i in range(args.instances):
        gi = rank * args.instances + i  # global instance id = env x seed
        per_env[envs[gi % 3]].append(gi)
    if args.mode == "batched":
        batches = []
        for e in envs:
            ids = per_env[e]
            if not ids:
                continue
            nx, nu = costs.ENV_DIMS[e]
            ah = costs.ENV_ACT_HIGH[e]
            bp = nlc.BatchedMPPIDelay(nlc.NLDynamics(models[e], DT), nlc.EnvRunningCost(e), nx, nlc.noise_sigma_for(nu), len(ids), seeds=ids,
                                      num_samples=args.K, horizon=args.H, device=dev, u_min=torch.tensor(-ah), u_max=torch.tensor(ah),
                                      u_scale=ah, math_mode=args.math)
            st = torch.stack([torch.tensor(np.array(start[e]) + np.random.default_rng(gi).uniform(-0.05, 0.05, nx), dtype=torch.float32)
                              for gi in ids]).to(dev).contiguous()
            batches.append({"env": e, "planner": bp, "states": st, "bufs": torch.zeros(len(ids), 4, nu, device=dev),
                            "reward": torch.zeros(len(ids), device=dev), "total": torch.zeros(len(ids), device=dev),
                            "stream": torch.cuda.Stream(device=dev)})

        def control_step():
            for b in batches:
                with torch.cuda.stream(b["stream"]):
                    a = b["planner"].command(b["states"], b["bufs"])
                    if not args.hold_state:
                        nlc.env_step(b["env"], b["states"], b["bufs"], a.float().contiguous(), args.delay, DT, b["reward"])
                        b["total"] += b["reward"]
            for b in batches:
                b["stream"].synchronize()
    else:
        for e in envs:
            for gi in per_env[e]:
                nx, nu = costs.ENV_DIMS[e]
                ah = costs.ENV_ACT_HIGH[e]
                rng = np.random.default_rng(gi)
                planners.append((e, nlc.MPPIDelay(nlc.NLDynamics(models[e], DT), nlc.EnvRunningCost(e), nx, nlc.noise_sigma_for(nu),
                                                  num_samples=args.K, horizon=args.H, device=dev, u_min=torch.tensor(-ah), u_max=torch.tensor(ah),
                                                  u_scale=ah, U_init=torch.zeros(args.H, nu, dtype=torch.float64), seed=gi, math_mode=args.math,
                                                  keep_states=False)))
                states.append(torch.tensor(np.array(start[e]) + rng.uniform(-0.05, 0.05, nx), dtype=torch.float32, device=dev).reshape(1, nx))
                bufs.append(torch.zeros(1, 4, nu, dtype=torch.float32, device=dev))
        streams = [torch.cuda.Stream(device=dev) for _ in range(args.streams)]

        def control_step():
            for i, (e, p) in enumerate(planners):
                with torch.cuda.stream(streams[i % args.streams]):
                    a = p.command(states[i][0], bufs[i][0])
                    nlc.env_step(e, states[i], bufs[i], a.float().reshape(1, -1).contiguous(), args.delay, DT)
            for s in streams:
                s.synchronize()

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
    if rank == 0 and args.mode == "batched":
        for b in batches:
            print(b["env"], "finite states:", bool(torch.isfinite(b["states"]).all()), "max|state|", float(b["states"].abs().max()),
                  "mean total reward", float(b["total"].mean()), file=sys.stderr)
    if rank == 0:
        total = world * args.instances * args.K * args.H
        print(json.dumps({"workload": f"config 5: {world * args.instances} instances x MPPI K={args.K} H={args.H}, instance-sharded over {world} GPU(s)",
                          "ms_per_control_step_all_instances": 1e3 * float(t[0]), "rollout_steps_per_s": total / float(t[0]),
                          "instances_per_gpu": args.instances, "mode": args.mode, "streams": args.streams if args.mode == "streams" else 3, "math": args.math, "n_gpus": world}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
