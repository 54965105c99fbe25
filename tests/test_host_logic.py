"""Host-side logic (no GPU): config -> problem mapping, parameter maps, Algorithm-2 updates, sharding, workload."""
import numpy as np
import pytest
import torch

from differentiable_tube_mpc_b200 import (AuxiliaryTheta, DBaSConfig, NominalTheta, TubeProblem, default_dubins_config,
                                          is_paper_mode)
from differentiable_tube_mpc_b200.adaptation import (NAMES_THETA, NAMES_THETA_BAR, PaperAdapter, apply_update,
                                                     split_raw_grads)
from differentiable_tube_mpc_b200.engine import shard_bounds
from differentiable_tube_mpc_b200.workload import make_instances
from helpers import load_golden, rel_err


@pytest.mark.parametrize("mode", ["paper", "full"])
def test_problem_from_config_matches_the_oracle_tables(oracle, mode):
    """Two independent readings of core/tube_mpc.py:51-165 / :674-768 (product spec.py vs oracle) agree, and both
    agree with what the reference configured in the goldens (checked in test_oracle_golden.py)."""
    cfg = default_dubins_config()
    prob = TubeProblem.from_config(cfg, mode=mode)
    env, nom, aux = oracle.env_from_cfg(cfg), *oracle.ocps_from_cfg(cfg, mode)
    e = prob.system.to_c()
    assert (e.dt, list(e.u_min), list(e.u_max), e.active_tol) == (env.dt, list(env.u_min), list(env.u_max), env.active_tol)
    assert (e.n_obs, e.obs_agg, e.obs_beta) == (env.n_obs, env.obs_agg, env.obs_beta)
    assert list(e.obs_cx)[:5] == list(env.obs_cx)[:5] and list(e.obs_r)[:5] == list(env.obs_r)[:5]
    for spec, o in ((prob.nominal, nom), (prob.auxiliary, aux)):
        c = spec.to_c()
        assert (c.horizon, c.max_iter, c.n_alpha, c.angle_wrap) == (o.T, o.max_iter, o.n_alpha, o.angle_wrap)
        assert (c.tol, c.reg, c.eps_dyn, c.eps_jac) == (o.tol, o.reg, o.eps_dyn, o.eps_jac)
        assert list(c.alphas)[: c.n_alpha] == list(o.alphas)[: o.n_alpha]
    if mode == "paper":
        assert np.array_equal(np.asarray(prob.theta_nom0), nom.theta())
        assert np.array_equal(np.asarray(prob.theta_aux0), aux.theta())
    else:
        th_n, th_a = prob.make_thetas("cpu", torch.float64)
        assert rel_err(th_n.constrained().numpy(), nom.theta()) < 1e-15
        assert rel_err(th_a.constrained().numpy(), aux.theta()) < 1e-15
        # full mode: alpha = softplus(0) + 1e-6, tight = softplus(0): the relaxed branch is live (App. B Q3)
        assert abs(float(th_n.alpha()) - (np.log(2.0) + 1e-6)) < 1e-15 and abs(float(th_n.tight()) - np.log(2.0)) < 1e-15


def test_mode_dispatch_and_single_obstacle_fallback():
    cfg = default_dubins_config()
    assert is_paper_mode(cfg)
    cfg["adaptation"]["adapt_nominal"] = True
    assert not is_paper_mode(cfg)
    cfg["paper_dubins_mode"] = False
    cfg["adaptation"]["adapt_nominal"] = False
    assert not is_paper_mode(cfg)
    cfg = default_dubins_config()
    del cfg["environment"]["obstacles"]
    prob = TubeProblem.from_config(cfg)
    assert prob.system.obs_agg == "single" and prob.system.obstacles[0].radius == 1.5
    cfg["system"]["name"] = "quadrotor"
    with pytest.raises(NotImplementedError):
        TubeProblem.from_config(cfg)
    with pytest.raises(ValueError):
        DBaSConfig(barrier_type="log")
    with pytest.raises(ValueError):
        DBaSConfig(alpha=-1.0)
    with pytest.raises(ValueError):
        DBaSConfig(gamma=1.5)


def test_theta_chain_rule_matches_autograd(oracle):
    g = torch.Generator().manual_seed(3)
    raw = torch.randn(12, dtype=torch.float64, generator=g)
    th = NominalTheta(raw[0:3].clone(), raw[3:5].clone(), raw[5:8].clone(), raw[8].clone(), raw[9].clone(),
                      raw[10].clone(), raw[11].clone())
    assert rel_err(th.constrained().numpy(), oracle.constrain_theta(raw.numpy())) < 1e-15
    w = torch.randn(12, dtype=torch.float64, generator=g)
    for t in th.tensors():
        t.requires_grad_(True)
    c = torch.cat([th.Q(), th.R(), th.Qf(), th.qb().view(1), th.alpha().view(1), th.gamma().view(1), th.tight().view(1)])
    grads = torch.autograd.grad((c * w).sum(), th.tensors())
    want = torch.cat([gg.reshape(-1) for gg in grads])
    assert rel_err((th.chain() * w).numpy(), want.numpy()) < 1e-14
    assert rel_err(oracle.chain_theta(raw.numpy(), w.numpy()), want.numpy()) < 1e-14
    ta = AuxiliaryTheta(raw[0:3].clone(), raw[3:5].clone(), raw[5:8].clone(), raw[8].clone(), raw[9].clone(),
                        raw[10].clone())
    assert ta.constrained()[11] == 0 and ta.chain()[11] == 0 and len(ta.tensors()) == 6


def test_paper_adapter_reproduces_the_reference_update():
    """v <- mu v + g ; clamp(theta - eta v)  (reference core/tube_mpc.py:978-984) against the closed-loop golden:
    the weights the reference solved with at step t+1 are the update of those at step t with its own gradient."""
    g = load_golden("closedloop_paper")
    cfg = default_dubins_config()
    th = g["aux_theta"][0]
    ad = PaperAdapter(Qa=torch.tensor(th[0, 0:3]), Ra=torch.tensor(th[0, 3:5]), qba=torch.tensor(th[0, 8]),
                      eta=cfg["adaptation"]["lr_eta"], mom=cfg["adaptation"]["momentum"])
    for t in range(th.shape[0] - 1):
        g12 = torch.zeros(12, dtype=torch.float64)
        g12[0:3], g12[3:5], g12[8] = torch.tensor(g["gQ"][0, t]), torch.tensor(g["gR"][0, t]), torch.tensor(g["gqb"][0, t])
        ad.update(g12)
        assert rel_err(ad.theta.numpy(), th[t + 1]) < 1e-15, t
    gQ, gR, gqb = PaperAdapter.grads_from_g12(torch.arange(12.0))
    assert gQ.tolist() == [5.0, 7.0, 9.0] and gR.tolist() == [3.0, 4.0] and float(gqb) == 8.0


def test_full_mode_update_reproduces_the_reference():
    """_apply_update on the RAW tensors (reference core/tube_mpc.py:232-249) against the closed-loop golden."""
    g = load_golden("closedloop_full")
    cfg = default_dubins_config()
    prob = TubeProblem.from_config(cfg, mode="full")
    th_n, th_a = prob.make_thetas("cpu", torch.float64)
    v_a = [torch.zeros_like(p) for p in th_a.tensors()]
    v_n = [torch.zeros_like(p) for p in th_n.tensors()]
    ad = cfg["adaptation"]
    kw = dict(lr=ad["lr_eta"], momentum=ad["momentum"], clip_norm=0.0, project_params=ad["project_params"])
    H = g["aux_theta"].shape[1]
    for t in range(H):
        assert rel_err(th_a.constrained().numpy(), g["aux_theta"][0, t]) < 1e-14, t
        assert rel_err(th_n.constrained().numpy(), g["nom_theta"][0, t]) < 1e-14, t
        ga = torch.zeros(12, dtype=torch.float64)
        ga[:11] = torch.tensor(g["g_theta_raw"][0, t])
        apply_update(th_a.tensors(), split_raw_grads(ga, False), v_a, NAMES_THETA, **kw)
        apply_update(th_n.tensors(), split_raw_grads(torch.tensor(g["g_theta_bar_raw"][0, t]), True), v_n,
                     NAMES_THETA_BAR, **kw)
    # projection bounds and norm clipping
    p = [torch.tensor([5.0, -3.0])]
    apply_update(p, [torch.tensor([300.0, 400.0])], [torch.zeros(2)], ["R_raw"], lr=1.0, clip_norm=5.0, project_params=True)
    assert torch.allclose(p[0], torch.tensor([2.0, 1e-4]))


def test_shard_bounds_partition_the_batch():
    for total, world in ((10, 3), (65536 * 8, 8), (7, 8), (0, 2)):
        spans = [shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_workload_is_a_pure_function_of_seed_and_index():
    cfg = default_dubins_config()
    x_all, xb_all = make_instances(20000, 2024, cfg)
    x, xb = make_instances(300, 2024, cfg)
    assert np.array_equal(x, x_all[:300]) and np.array_equal(xb, xb_all[:300])
    assert np.abs(x - xb).max() <= 0.05
    x2, _ = make_instances(300, 2025, cfg)
    assert not np.array_equal(x, x2)


def test_operators_refuse_cpu_tensors():
    """No CPU fallback: host tensors are rejected, never silently computed."""
    from differentiable_tube_mpc_b200 import ops

    prob = TubeProblem.from_config(default_dubins_config(), mode="paper")
    x0 = torch.zeros(2, 4, dtype=torch.float64)
    V = torch.zeros(2, 50, 2, dtype=torch.float64)
    th = torch.tensor(prob.theta_nom0, dtype=torch.float64)
    with pytest.raises(ValueError, match="CUDA"):
        ops.ilqr_solve_batched(prob.system, prob.nominal, x0, V, th)
    with pytest.raises(ValueError, match="CUDA"):
        ops.rollout_batched(prob.system, prob.nominal, x0, V, th)
    with pytest.raises(ValueError, match="CUDA"):
        ops.reduce_columns(torch.zeros(4, 3, dtype=torch.float64))
    from differentiable_tube_mpc_b200.engine import BatchedTubeMPC

    with pytest.raises(ValueError, match="CUDA"):
        BatchedTubeMPC(prob, device="cpu")


def test_packed_adapters_reproduce_the_reference_updates():
    """adaptation.RawAdapter (the packed, in-place, graph-capturable form the closed loop uses) against the closed-loop
    golden: raw parameters and momentum of theta and theta_bar after every reference step, from the constrained-parameter
    gradient the kernels produce (raw gradient = constrained gradient x chain factor)."""
    from differentiable_tube_mpc_b200.adaptation import RawAdapter, chain_factor, constrain

    g = load_golden("closedloop_full")
    cfg = default_dubins_config()
    prob = TubeProblem.from_config(cfg, mode="full")
    ad = cfg["adaptation"]
    kw = dict(lr=ad["lr_eta"], momentum=ad["momentum"], clip_norm=0.0, project_params=ad["project_params"])
    t64 = lambda v: torch.tensor(v, dtype=torch.float64)
    th_a = RawAdapter(t64(list(prob.theta_aux0) + [0.0]), with_tight=False, **kw)
    th_n = RawAdapter(t64(prob.theta_nom0), with_tight=True, **kw)
    keep = (th_a.theta, th_n.theta, th_a.raw, th_n.vel)
    for t in range(g["aux_theta"].shape[1]):
        assert rel_err(th_a.theta.numpy(), g["aux_theta"][0, t]) < 1e-14, t
        assert rel_err(th_n.theta.numpy(), g["nom_theta"][0, t]) < 1e-14, t
        ga_raw = torch.zeros(12, dtype=torch.float64)
        ga_raw[:11] = t64(g["g_theta_raw"][0, t])
        gn_raw = t64(g["g_theta_bar_raw"][0, t])
        # the kernels hand over d/d(constrained); divide the reference's raw gradient by the chain factor to get it
        ca, cn = chain_factor(th_a.raw, False), chain_factor(th_n.raw, True)
        th_a.update(torch.where(ca != 0, ga_raw / torch.where(ca != 0, ca, torch.ones_like(ca)), torch.zeros_like(ca)))
        th_n.update(torch.where(cn != 0, gn_raw / torch.where(cn != 0, cn, torch.ones_like(cn)), torch.zeros_like(cn)))
    assert keep[0] is th_a.theta and keep[1] is th_n.theta and keep[2] is th_a.raw and keep[3] is th_n.vel   # in place
    assert torch.equal(th_a.theta, constrain(th_a.raw, False)) and float(th_a.theta[11]) == 0.0
    # per-group norm clip and projection, on the device tensor (no host round trip): groups R (2 entries) and gamma
    a = RawAdapter(t64([1, 1, 0, 5, -3, 1, 1, 1, 0.5, 0.2, 0.0, 0.3]), with_tight=True, lr=1.0, clip_norm=5.0, project_params=True)
    g12 = torch.zeros(12, dtype=torch.float64)
    g12[3:5] = t64([300.0, 400.0]) / chain_factor(a.raw, True)[3:5]
    a.update(g12)
    assert torch.allclose(a.raw[3:5], t64([2.0, 1e-4])) and torch.equal(a.raw[0:3], t64([1, 1, 0]))


def test_executed_work_model_reads_the_committed_counters():
    """bench.py's executed-work roofline: the committed ncu counters (profiles/r02_step_counters.json) times the launches
    a step ran reproduce the step totals they were reduced from (same batch, same iteration counts)."""
    import json
    import os

    import bench

    c = bench.load_counters()
    assert c is not None and {"init_n", "ric_n", "fwdc_n", "ric_a", "fwdc_a", "sens_a", "sens_n"} <= set(c["per_tile_step"])
    rows = [r for r in open(os.path.join(os.path.dirname(bench.COUNTERS), "r02m_step_launches_full_B65536.csv")) if r.startswith('"')]
    assert len(rows) > 100
    # iterations of that profiled step: full mode, 4 evaluated step sizes per OCP; nominal 9.918 / auxiliary 14.714 per solve
    tiles, T, B = (c["batch"] + 31) // 32, c["horizon"], c["batch"]
    ex = bench.executed_work(c, T, tiles, [(4, 9.917648 * B / 32), (4, 14.714111 * B / 32)], 2)
    for k in ("fp64_inst", "inst", "dram_bytes"):
        assert abs(ex[k] - c["step_totals"][k]) < 0.06 * c["step_totals"][k], (k, ex[k], c["step_totals"][k])
    assert 50e3 < c["per_solve"]["fp64_inst"] < 65e3 and 0.6e6 < c["per_solve"]["dram_bytes"] < 0.9e6
