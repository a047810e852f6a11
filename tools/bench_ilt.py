#!/usr/bin/env python
"""BASELINE config 2: Fourier-series inverse Laplace microbenchmark, 1M trajectories x 100 time points, S in {33,65,129}
s-points, on one B200.  Prints one JSON line per S: achieved algorithmic GB/s (8*S + 8 bytes per output point) against the
measured HBM copy bandwidth of MEASURED_PEAKS.json, plus a CPU oracle timing on a bounded sample."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neurallaplacecontrol_b200 import fourier_ilt  # noqa: E402
from oracle import ilt  # noqa: E402


def main():
    N = int(os.environ.get("ILT_N", 1_000_000))
    n_t = 100
    peak = 6542.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = json.load(open(p))["hbm_gbs"]
    t = (torch.arange(n_t, dtype=torch.float32, device="cuda") + 1) * 0.05
    for S in (33, 65, 129):
        g = torch.Generator(device="cuda").manual_seed(2)
        F = torch.empty((N, n_t, S, 2), dtype=torch.float32, device="cuda")
        for i in range(0, N, 100_000):  # U(-1,1), chunked to bound the temporaries
            F[i:i + 100_000].uniform_(-1, 1, generator=g)
        Fc = torch.view_as_complex(F)
        out = torch.empty((N, n_t), dtype=torch.float32, device="cuda")
        for _ in range(3):
            fourier_ilt(Fc, t, out=out)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for s, e in ev:  # the input (>= 26 GB) is far larger than L2: no flush needed
            s.record(); fourier_ilt(Fc, t, out=out); e.record()
        torch.cuda.synchronize()
        ms = sum(s.elapsed_time(e) for s, e in ev) / len(ev)
        nbytes = N * n_t * (8 * S + 8)
        # parity on a sample against the oracle
        idx = torch.randint(0, N, (64,), device="cuda")
        Fs = Fc[idx].cpu()
        t64 = t.cpu().double().expand(64, n_t)
        ref = ilt.fourier_line_integrate(Fs.real.double(), Fs.imag.double(), t64, ilt.SCALE * (t64 + ilt.EPS))
        err = float((ref - out[idx].cpu().double()).abs().max() / ref.abs().max())
        # CPU oracle on a bounded sample
        Ns = 20000
        Fcpu = Fc[:Ns].cpu()
        tc = t.cpu().double().expand(Ns, n_t)
        torch.set_num_threads(os.cpu_count())
        t0 = time.perf_counter()
        ilt.fourier_line_integrate(Fcpu.real.double(), Fcpu.imag.double(), tc, ilt.SCALE * (tc + ilt.EPS))
        cpu_s = time.perf_counter() - t0
        print(json.dumps({"workload": f"ILT microbench N={N} n_t={n_t} S={S}", "ms": ms, "points_per_s": N * n_t / (ms * 1e-3),
                          "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": nbytes / (ms * 1e-3) / 1e9 / peak, "bytes_per_point": 8 * S + 8},
                          "relerr_vs_oracle_sample": err,
                          "cpu_baseline": {"points_per_s": Ns * n_t / cpu_s, "cores": os.cpu_count(), "kind": "port",
                                           "sample": f"{Ns} trajectories, fp64"}}), flush=True)
        del F, Fc, out


if __name__ == "__main__":
    main()
