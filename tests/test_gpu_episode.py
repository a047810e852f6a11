"""Closed loop (BASELINE config 5): the on-device environment step against golden vectors produced with the
reference's own oracle.py dynamics, the instance-batched planner against stand-alone planners, and the loop itself."""
import numpy as np
import pytest
import torch

from _util import DT, load, relerr, short

pytestmark = pytest.mark.gpu

ENVS = ["oderl-pendulum", "oderl-cartpole", "oderl-acrobot"]


@pytest.mark.parametrize("env", ENVS)
@pytest.mark.parametrize("delay", [0, 1, 3])
def test_env_step_matches_reference_dynamics(env, delay):
    """step_env (mppi_with_model.py:193-216): buffer roll + delayed action (get_action), Euler step (reference oracle.py),
    reward of the new state - six consecutive steps of 48 instances."""
    import neurallaplacecontrol_b200 as nlc

    g = load("env_step_" + short(env))
    states = torch.tensor(g[f"d{delay}_state0"], dtype=torch.float32).cuda().contiguous()
    bufs = torch.tensor(g[f"d{delay}_buf0"], dtype=torch.float32).cuda().contiguous()
    rewards = torch.zeros(states.shape[0], device="cuda")
    for it in range(g[f"d{delay}_actions"].shape[0]):
        a = torch.tensor(g[f"d{delay}_actions"][it], dtype=torch.float32).cuda().contiguous()
        nlc.env_step(env, states, bufs, a, delay, DT, rewards)
        torch.cuda.synchronize()
        assert torch.equal(bufs.cpu(), torch.tensor(g[f"d{delay}_bufs"][it], dtype=torch.float32))  # pure data movement
        assert relerr(g[f"d{delay}_states"][it], states) < 2e-5, (it, relerr(g[f"d{delay}_states"][it], states))
        assert relerr(g[f"d{delay}_rewards"][it], rewards) < 2e-5, (it, relerr(g[f"d{delay}_rewards"][it], rewards))


def test_env_step_rejects_bad_arguments():
    import neurallaplacecontrol_b200 as nlc

    s = torch.zeros(2, 3, device="cuda")
    b = torch.zeros(2, 4, 1, device="cuda")
    a = torch.zeros(2, 1, device="cuda")
    with pytest.raises(RuntimeError):
        nlc.env_step("oderl-pendulum", s, b, a, 4)  # delay outside the 4-entry buffer
    with pytest.raises(TypeError):
        nlc.env_step("oderl-pendulum", s.cpu(), b, a, 1)  # no CPU path
    with pytest.raises(RuntimeError):
        nlc.env_step("oderl-acrobot", torch.zeros(2, 6, device="cuda"), b, a, 1)  # nu mismatch


def _make(env, mode, **kw):
    import neurallaplacecontrol_b200 as nlc
    from oracle import costs
    from test_gpu_parity import make_model

    nx, nu = costs.ENV_DIMS[env]
    ah = np.float32(costs.ENV_ACT_HIGH[env])
    m = make_model(env, True, math_mode=mode)
    common = dict(num_samples=kw.pop("K"), horizon=kw.pop("T"), device="cuda:0", lambda_=1.0, u_min=torch.tensor(-ah), u_max=torch.tensor(ah),
                  u_scale=ah, math_mode=mode)
    return nlc, m, nx, nu, ah, common


@pytest.mark.parametrize("env", ["oderl-pendulum", "oderl-acrobot"])
@pytest.mark.parametrize("mode", ["fp32", "tc_split3"])
def test_batched_planner_reproduces_standalone_planners(env, mode, monkeypatch):
    """Instance i of a batch == MPPIDelay(seed=seeds[i]) on the same state / buffer / U: bit-identical samples, costs
    and actions over two consecutive control steps (same kernels, row-independent arithmetic)."""
    monkeypatch.setenv("NLC_ROLLOUT_TILES", "1")  # both sides on the same rollout form regardless of I*K
    I, K, T, B = 3, 200, 6, 4
    nlc, m, nx, nu, ah, common = _make(env, mode, K=K, T=T)
    gen = torch.Generator().manual_seed(5)
    U0 = torch.randn(I, T, nu, generator=gen, dtype=torch.float64) * 0.3
    states = torch.tensor(load("env_step_" + short(env))["d1_state0"][:I], dtype=torch.float64)
    bufs = (torch.rand(I, B, nu, generator=gen, dtype=torch.float64) * 2 - 1) * float(ah)
    seeds = [11, 12, 13]
    batch = nlc.BatchedMPPIDelay(nlc.NLDynamics(m, DT), nlc.EnvRunningCost(env), nx, nlc.noise_sigma_for(nu), I, seeds=seeds, U_init=U0, **common)
    singles = [nlc.MPPIDelay(nlc.NLDynamics(m, DT), nlc.EnvRunningCost(env), nx, nlc.noise_sigma_for(nu), U_init=U0[i], seed=seeds[i],
                             keep_states=False, **common) for i in range(I)]
    for call in range(2):
        act_b = batch.command(states, bufs)
        torch.cuda.synchronize()
        for i, p in enumerate(singles):
            a = p.command(states[i].cuda(), bufs[i].cuda())  # device inputs: the on-device sampler path, like the batch
            torch.cuda.synchronize()
            assert torch.equal(p.noise, batch.noise[i]), (call, i)
            assert torch.equal(p.cost_total, batch.cost_total[i]), (call, i)
            assert torch.equal(a.float(), act_b[i].float()), (call, i)
            assert torch.equal(p.U.float(), batch.U[i]), (call, i)
        for i in range(I):
            bufs[i], _ = nlc.get_action(bufs[i], act_b[i].cpu().double(), 1)
    assert abs(float(batch.omega.sum()) - I) < 1e-4


def test_batched_planner_with_injected_noise_matches_reference_plan():
    """The golden single-planner plan (reference MPPIDelay + NeuralLaplaceModel) replicated in 2 of 3 instances."""
    env = "oderl-cartpole"
    g = load("plan_cal_cartpole_calls1")
    noise = torch.from_numpy(g["in_noise"])
    noise = noise[0] if noise.dim() == 4 else noise
    K, T, _ = noise.shape
    nlc, m, nx, nu, ah, common = _make(env, "fp32", K=K, T=T)
    I = 3
    batch = nlc.BatchedMPPIDelay(nlc.NLDynamics(m, DT), nlc.EnvRunningCost(env), nx, nlc.noise_sigma_for(nu), I,
                                 U_init=torch.from_numpy(g["in_U"]), keep_states=True, **common)
    states = torch.from_numpy(np.asarray(g["in_state"])).reshape(1, nx).repeat(I, 1)
    states[1] += 0.1  # the odd one out must not disturb its neighbours
    bufs = torch.from_numpy(g["in_buffer"]).unsqueeze(0).repeat(I, 1, 1)
    nz = noise.unsqueeze(0).repeat(I, 1, 1, 1)
    act = batch.command(states, bufs, nz)
    torch.cuda.synchronize()
    for i in (0, 2):
        assert relerr(g["cost_total"], batch.cost_total[i]) < 1e-4
        assert relerr(g["states"], batch.states[i]) < 1e-4
        assert relerr(g["omega"], batch.omega[i]) < 1e-4
        assert relerr(g["U"], batch.U[i]) < 1e-4
    assert not torch.equal(batch.cost_total[1], batch.cost_total[0])
    assert torch.equal(act[0], act[2])


@pytest.mark.parametrize("dynamics", ["analytic", "neural_laplace"])
def test_closed_loop_batch_equals_single_instance_loops(dynamics, monkeypatch):
    """run_closed_loop over I instances == I single-instance loops (MPPIDelay.command + env_step with I = 1), and with
    the analytic dynamics the pendulum swings towards upright (reward improves over a random policy)."""
    monkeypatch.setenv("NLC_ROLLOUT_TILES", "1")
    env, delay, I, K, T, n_steps = "oderl-pendulum", 1, 4, 256, 10, 8
    nlc, m, nx, nu, ah, common = _make(env, "fp32", K=K, T=T)
    dyn = nlc.AnalyticDelayDynamics(env, delay, DT) if dynamics == "analytic" else nlc.NLDynamics(m, DT)
    seeds = [3, 4, 5, 6]
    st0 = torch.tensor(load("env_step_pendulum")["d1_state0"][:I], dtype=torch.float64)
    batch = nlc.BatchedMPPIDelay(dyn, nlc.EnvRunningCost(env), nx, nlc.noise_sigma_for(nu), I, seeds=seeds, **common)
    res = nlc.run_closed_loop(batch, env, st0, delay, n_steps, DT)
    assert res["total_reward_raw"].shape == (I,) and np.isfinite(res["total_reward_raw"]).all()
    for i in range(I):
        p = nlc.MPPIDelay(dyn, nlc.EnvRunningCost(env), nx, nlc.noise_sigma_for(nu), U_init=torch.zeros(T, nu, dtype=torch.float64),
                          seed=seeds[i], keep_states=False, **common)
        s = st0[i:i + 1].float().cuda().contiguous()
        b = torch.zeros(1, 4, nu, device="cuda")
        r = torch.zeros(1, device="cuda")
        tot = 0.0
        for it in range(n_steps):
            a = p.command(s[0], b[0])
            nlc.env_step(env, s, b, a.float().reshape(1, nu).contiguous(), delay, DT, r)
            tot += float(r[0])
        assert abs(tot - res["total_reward_raw"][i]) <= 1e-5 * max(1.0, abs(tot)), (i, tot, res["total_reward_raw"][i])
        assert relerr(s.double().cpu().numpy()[0], res["final_states"][i]) < 1e-6
