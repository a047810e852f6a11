"""CPU-side checks of the boundary: the shared library loads, exports every symbol ``include/nlc_b200.h`` declares,
refuses to compute without an sm_100 device (no CPU fallback), and the K-sharding host logic works over gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nlc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nlc_[A-Za-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from neurallaplacecontrol_b200 import _lib

    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nlc_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes prototypes and header disagree"
    assert lib.nlc_version() == 100


def test_struct_layouts_match_header_sizes():
    """sizeof() of the ctypes mirrors against a C program compiled from the header."""
    from neurallaplacecontrol_b200 import _lib

    prog = ('#include <stdio.h>\n#include "nlc_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(nlc_model_desc), '
            'sizeof(nlc_mppi_params), sizeof(nlc_rollout_opts), sizeof(nlc_planner_desc));return 0;}\n')
    exe = os.path.join(ROOT, "tests", "_sizes.out")
    try:
        subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=prog.encode(), check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, check=True).stdout.split()]
    finally:
        if os.path.exists(exe):
            os.remove(exe)
    assert got == [C.sizeof(_lib.ModelDesc), C.sizeof(_lib.MppiParams), C.sizeof(_lib.RolloutOpts), C.sizeof(_lib.PlannerDesc)]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import numpy as np

    import neurallaplacecontrol_b200 as nlc
    from neurallaplacecontrol_b200 import _lib

    lib = _lib.load()
    assert lib.nlc_device_check(0) == -3  # NLC_ERR_ARCH
    assert b"no CPU path" in lib.nlc_last_error()
    h = C.c_void_p()
    assert lib.nlc_model_create(C.byref(h), C.byref(_lib.ModelDesc()), 0) == -3
    assert lib.nlc_planner_create(C.byref(h), None, C.byref(_lib.PlannerDesc()), 0) == -3
    m = nlc.NeuralLaplaceModel(3, 1, 3, hidden_units=128, s_recon_terms=17, state_mean=np.zeros(3), state_std=np.ones(3),
                               action_mean=np.zeros(1), action_std=np.ones(1), normalize=True, normalize_time=True)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 3), torch.zeros(2, 4, 1), torch.full((2, 1), 0.05))
    with pytest.raises(RuntimeError):
        nlc.MPPIDelay(nlc.NLDynamics(m), nlc.EnvRunningCost("oderl-pendulum"), 3, nlc.noise_sigma_for(1))
    with pytest.raises(RuntimeError):
        nlc.fourier_ilt(torch.zeros(1, 1, 3, dtype=torch.complex64), torch.ones(1))


def test_state_dict_layout_matches_reference():
    import numpy as np

    import neurallaplacecontrol_b200 as nlc
    from oracle.nl_model import STATE_DICT_KEYS

    from _util import weights

    m = nlc.NeuralLaplaceModel(5, 1, 5, hidden_units=128, s_recon_terms=17, state_mean=np.zeros(5), state_std=np.ones(5),
                               action_mean=np.array([0]), action_std=np.array([1.5]), normalize=True, normalize_time=True)
    assert tuple(m.state_dict().keys()) == STATE_DICT_KEYS
    assert m.state_dict()["dt"].dtype == torch.float32  # torch.tensor(0.05), as w_nl.py:111
    res = m.double().load_state_dict(weights("oderl-cartpole"))
    assert not res.missing_keys and not res.unexpected_keys


def test_shard_range():
    from neurallaplacecontrol_b200 import sharding

    assert [sharding.shard_range(64, 4, r) for r in range(4)] == [(0, 16), (16, 16), (32, 16), (48, 16)]
    with pytest.raises(ValueError):
        sharding.shard_range(10, 4, 0)
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from neurallaplacecontrol_b200 import sharding
from oracle import mppi
rank, G = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=G)
g = torch.Generator().manual_seed(0)
K, T, nu, lam = 96, 5, 2, 0.7
cost = torch.rand(K, generator=g, dtype=torch.float64) * 30
noise = torch.randn(K, T, nu, generator=g, dtype=torch.float64)
U = torch.randn(T, nu, generator=g, dtype=torch.float64)
off, n = sharding.shard_range(K, G, rank)
b, e, W = mppi.shard_triple(cost[off:off + n], noise[off:off + n], lam)
triple = torch.cat([b.view(1), e.view(1), W.reshape(-1)])
allt = sharding.gather_triples(triple)
triples = [(allt[i, 0], allt[i, 1], allt[i, 2:].view(T, nu)) for i in range(G)]
U_sh, _, _ = mppi.combine_shards(U, triples, lam)
U_ref, _, _ = mppi.softmax_update(U, cost, noise, lam)
assert (U_sh - U_ref).abs().max() < 1e-12, (U_sh - U_ref).abs().max()
gathered = [torch.zeros_like(U_sh) for _ in range(G)]
dist.all_gather(gathered, U_sh)
assert all(torch.equal(gathered[0], x) for x in gathered)  # replicated U stays bit-identical without a broadcast
dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_combine_over_gloo_world_size_2():
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    code = _WORKER.format(root=ROOT, port=port)
    procs = [subprocess.Popen([sys.executable, "-c", code, str(r), "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err.decode()[-2000:]
        assert b"ok" in out


def test_overlap_schedule_host_logic():
    """The planner's schedule for the encoder beside the rollout (``csrc/planner.cu`` ``overlap_schedule``; host arithmetic
    only, no device): the rollout's form and SM count and the encoder tiles that run on all SMs before the fork, at the shard
    sizes of config 4 (strong scaling over 8 / 4 / 2 GPUs), at configs 1 and 3, and at the form crossover."""
    from neurallaplacecontrol_b200 import _lib

    lib = _lib.load()
    fn = lib.nlc_debug_overlap_schedule
    fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    fn.restype = C.c_int

    def sched(K, T, nx=6, S=17):
        out = (C.c_longlong * 4)()
        assert fn(K, T, nx, S, out) == 0
        return tuple(out)

    # 8 GPUs: 64 tiles, one tile per CTA on 64 SMs, the first 6 of 50 steps encoded before the fork
    assert sched(8192, 50) == (0, 64, 6 * 64, 50 * 64)
    # 4 GPUs: 128 tiles, ping-pong form on 64 SMs, 19 steps before the fork
    pp, sms, split, total = sched(16384, 50)
    assert (pp, sms, total) == (1, 64, 50 * 128) and split == 19 * 128
    # 2 GPUs: 256 tiles, ping-pong form on 128 SMs, only the last steps beside the rollout on the 20 spare SMs
    pp, sms, split, total = sched(32768, 50)
    assert (pp, sms) == (1, 128) and 0 < total - split <= 4 * 256
    # form crossover at 89 tiles
    assert sched(88 * 128, 50)[:2] == (0, 88) and sched(89 * 128, 50)[:2] == (1, 45)
    # config 1 (8 tiles): everything beside the rollout; config 3 (64 tiles, T = 30): a few steps first
    assert sched(1000, 20, nx=3) == (0, 8, 0, 20 * 8)
    pp, sms, split, total = sched(8192, 30, nx=5)
    assert (pp, sms) == (0, 64) and split % 64 == 0 and 0 < split < total // 4
    # S = 33 with nx >= 5 has no ping-pong instantiation: one tile per CTA whatever the size
    assert sched(16384, 50, nx=6, S=33)[:2] == (0, 128)
