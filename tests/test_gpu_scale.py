"""Parity at the size that is benchmarked (BASELINE configs[2]): the CUDA path against the C oracle on the WHOLE
bench batch -- full mode, T = 50, B = 65,536, per-instance theta, seed 2024 (bench.make_workload) -- through the same
library call bench.py times (`dtmpc_tube_step`).  The oracle needs a few seconds per pass on the box's host cores.

Bar: 1e-9 relative (tensor infinity norm per instance) on states, controls, sensitivities and gradients; identical
discrete decisions.  Two classes of instances are exempt from the bitwise-decision / 1e-9 demand, and their MEASURED
shares are printed and bounded at what they really are, not at a comfortable allowance:
  * ill-conditioned instances: the oracle's own answer moves by more than 1e-10 under a 1e-14 relative perturbation of
    x0 (trajectories grazing an obstacle, jacobian entries ~1e11); compared at 10 x that measured sensitivity;
  * decision ties: accepted-step / stopping decisions whose margin in the oracle's own costs is a roundoff tie
    (SURVEY App. B Q12).
"""
import os

import numpy as np
import pytest
import torch

import bench
from helpers import per_instance_rel

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-9
B_BENCH = int(os.environ.get("DTMPC_SCALE_B", bench.BATCH_PER_GPU))


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)


@pytest.mark.parametrize("mode", ["full"])
def test_bench_batch_matches_the_oracle(oracle, mode):
    _batch_vs_oracle(oracle, mode, B_BENCH, bench.T_HORIZON, max_mismatch=0.002, max_illc=0.002)


# BASELINE configs[3] (the horizon sweep of bench.py / tools/sweep.py) at its own horizons; the bounds are the measured
# shares (printed), rounded up
# measured on B200: T = 20: 9 tie mismatches of 16,384, ill-conditioned 0.16 %; T = 100: 0 / 0.95 %; T = 200: 0 / 5.3 %
# (a 1e-14 perturbation of x0 grows with the horizon: more instances pass the 1e-10 mark and are compared at 10 x their
# own measured sensitivity instead of 1e-9)
@pytest.mark.parametrize("T,B,max_mismatch,max_illc", [(20, 16384, 0.002, 0.003), (100, 4096, 0.002, 0.012),
                                                       (200, 2048, 0.002, 0.06)])
def test_sweep_horizons_match_the_oracle(oracle, T, B, max_mismatch, max_illc):
    _batch_vs_oracle(oracle, "full", B, T, max_mismatch=max_mismatch, max_illc=max_illc)


def _batch_vs_oracle(oracle, mode, B, T, max_mismatch, max_illc):
    from differentiable_tube_mpc_b200 import default_dubins_config, ops

    cfg = default_dubins_config()
    cfg["system"]["horizon_N"] = T
    prob, w = bench.make_workload(cfg, mode, 0, B, B)
    env = oracle.env_from_cfg(cfg)
    nom, aux = oracle.ocps_from_cfg(cfg, mode)
    full = mode == "full"
    th_n, th_a = w["th_n"], w["th_a"]
    # initial barrier states from the product (b0 parity is pinned by test_gpu_parity.py); both sides get the same bits
    eng_b0 = lambda x, th, ocp: ops.primitives_batched(prob.system, ocp, cu(np.concatenate([x, np.zeros((B, 1))], 1)),
                                                        cu(np.zeros((B, 2))), cu(th))["b0"].cpu().numpy()
    x0n = np.concatenate([w["xb"], eng_b0(w["xb"], th_n, prob.nominal)[:, None]], 1)
    x0a = np.concatenate([w["x"], eng_b0(w["x"], th_a, prob.auxiliary)[:, None]], 1)
    V0 = np.zeros((B, T, 2))
    threads = os.cpu_count() or 8
    want = oracle.tube_solve_batch(env, nom, aux, x0n, x0a, V0, V0, theta_n=th_n, theta_a=th_a, mode=mode, n_threads=threads)
    pert = oracle.tube_solve_batch(env, nom, aux, x0n * (1 + 1e-14), x0a * (1 + 1e-14), V0, V0, theta_n=th_n,
                                   theta_a=th_a, mode=mode, n_threads=threads)
    # k = -Q_uu^-1 Q_u is a residual (~0 at the optimum) left by cancellation in Q_u = l_u + B^T V_x, and V_x sums one
    # term per remaining step: its absolute roundoff grows with the horizon, so does the floor it is compared on
    k_floor = 1e-3 * max(1.0, T / 50.0)
    cond = np.zeros(B)
    for k in ("Xn", "Un", "Xa", "Ua", "dXa", "dLa", "Kn", "Ka"):
        cond = np.maximum(cond, per_instance_rel(pert[k], want[k], floor=1e-6))
    for k in ("kn", "ka"):
        cond = np.maximum(cond, per_instance_rel(pert[k], want[k], floor=k_floor))
    del pert

    rn, ra, sa, sn = ops.tube_step_batched(prob.system, prob.nominal, prob.auxiliary, cu(x0n), cu(x0a), cu(V0), cu(V0),
                                           cu(th_n), cu(th_a), full=full, gains=True, trace=True)
    torch.cuda.synchronize()
    status = rn.status.cpu().numpy()
    assert np.array_equal((status & 0xff) != 0, want["status"] != 0)
    ok = want["status"] == 0
    assert ok.mean() > 0.999, ok.mean()

    # ---- discrete decisions -----------------------------------------------------------------------------------
    same = np.ones(B, dtype=bool)
    tie = np.zeros(B, dtype=bool)
    for tag, r, ocp in (("n", rn, nom), ("a", ra, aux)):
        it_w, it_g = want[f"iters_{tag}"], r.iters.cpu().numpy()
        ai_w, ai_g = want[f"aidx_{tag}"], r.alpha_idx.cpu().numpy()
        n_it = ai_w.shape[1]
        col = np.arange(n_it)[None, :]
        live = col < np.minimum(it_w, it_g)[:, None]
        s = (it_w == it_g) & np.all((ai_w == ai_g) | ~live, axis=1)
        same &= s
        # a mismatch is a roundoff tie if, up to the first differing iteration, the oracle's best/second-best margin or
        # its distance to the stopping tolerance is below 1e-10 relative
        marg, costs = want[f"marg_{tag}"], want[f"costs_{tag}"]
        with np.errstate(invalid="ignore", divide="ignore"):
            rel_m = np.where(live, np.abs(marg) / np.maximum(np.abs(costs), 1e-300), np.inf).min(axis=1)
            dc = np.abs(np.diff(costs, axis=1))
            near_tol = np.where(col[:, 1:] < it_w[:, None], np.abs(dc - ocp.tol) / np.maximum(np.abs(costs[:, :1]), 1e-300),
                                np.inf).min(axis=1)
        tie |= ~s & ((rel_m < 1e-10) | (near_tol < 1e-9))
    mism = ok & ~same
    illc = ok & (cond > 1e-10)
    unexplained = mism & ~tie & ~(cond > 1e-12)
    print(f"[scale {mode}] T={T} B={B}: decisions identical on {(ok & same).mean():.4%}; mismatches {mism.sum()} "
          f"(ties {int((mism & tie).sum())}, ill-conditioned {int((mism & ~tie & (cond > 1e-12)).sum())}, "
          f"unexplained {int(unexplained.sum())}); ill-conditioned share {illc.mean():.3%}")
    assert unexplained.sum() == 0, np.nonzero(unexplained)[0][:10]
    # measured on B200 (round 2) at T = 50, B = 65,536: 3 mismatches (0.005 %, all ties), ill-conditioned share 0.12 %
    # (0.02 % by states / controls / sensitivities alone; the rest are instances whose feed-forward gain k moves)
    assert mism.mean() <= max_mismatch, f"decision mismatch share {mism.mean():.3%}"
    assert illc.mean() <= max_illc, f"ill-conditioned share {illc.mean():.3%}"

    # ---- values ------------------------------------------------------------------------------------------------
    def compare(got, key, sel, floor=0.0):
        e = per_instance_rel(got.cpu().numpy()[sel], want[key][sel], floor=floor)
        tol_i = np.maximum(TOL, 10.0 * cond[sel])
        bad = e > tol_i
        print(f"[scale {mode}] {key:6s} n={e.size} median {np.median(e):.1e} p99 {np.quantile(e, 0.99):.1e} "
              f"max {e.max():.1e}  > 1e-9: {(e > TOL).mean():.3%}")
        assert not bad.any(), f"{key}: {bad.sum()} instances off, worst {e[bad].max():.3e}"

    sel = ok & same
    compare(rn.X, "Xn", sel); compare(rn.V, "Un", sel); compare(ra.X, "Xa", sel); compare(ra.V, "Ua", sel)
    compare(rn.K, "Kn", sel); compare(ra.K, "Ka", sel)
    compare(rn.k, "kn", sel, floor=k_floor); compare(ra.k, "ka", sel, floor=k_floor)
    # sensitivities and gradients come out of the same call (evaluated at the product's own optimum)
    well = sel & (cond <= 1e-11)
    compare(sa.delta_X, "dXa", well, floor=1e-6); compare(sa.delta_V, "dVa", well, floor=1e-6)
    compare(sa.delta_lambda, "dLa", well, floor=1e-6)
    scale_a = np.abs(want["dLa"]).reshape(B, -1).max(axis=1)
    e = np.abs(sa.g_theta.cpu().numpy() - want["g_aux"]).max(axis=1) / np.maximum(
        np.maximum(np.abs(want["g_aux"]).max(axis=1), 1e-6 * scale_a), 1e-300)
    print(f"[scale {mode}] g_aux  median {np.median(e[well]):.1e} max {e[well].max():.1e}")
    assert e[well].max() < 10 * TOL
    assert np.abs(sa.loss.cpu().numpy()[sel] - want["loss"][sel]).max() <= 1e-9 * np.abs(want["loss"][sel]).max()
    if full:
        compare(sn.delta_X, "dXn", well, floor=1e-6); compare(sn.delta_lambda, "dLn", well, floor=1e-6)
        scale_n = np.abs(want["dLn"]).reshape(B, -1).max(axis=1)
        # (an instance whose auxiliary controls are all saturated has delta = 0: its nominal upper gradient and g_nom vanish)
        e = np.abs(sn.g_theta.cpu().numpy() - want["g_nom"]).max(axis=1) / np.maximum(
            np.maximum(np.abs(want["g_nom"]).max(axis=1), 1e-6 * scale_n), 1e-300)
        print(f"[scale {mode}] g_nom  median {np.median(e[well]):.1e} max {e[well].max():.1e}")
        assert e[well].max() < 10 * TOL
        # the batch-summed gradient that Algorithm 2 consumes (and the ranks all-reduce)
        gs = ops.reduce_columns(sn.g_theta).cpu().numpy()
        ref = want["g_nom"][well].sum(0)
        got_ok = sn.g_theta.cpu().numpy()[well].sum(0)
        assert np.abs(got_ok - ref).max() <= 1e-8 * np.abs(want["g_nom"][well]).sum(0).max(), (got_ok, ref)
        assert np.allclose(gs, sn.g_theta.cpu().numpy().sum(0), rtol=1e-10, atol=1e-8 * np.abs(gs).max())


def test_nominal_config2_batch_matches_the_oracle(oracle):
    """BASELINE configs[1] at its own shape: 4,096 nominal-only iLQR solves with run_nominal.py semantics (wrapped heading
    error, reg = ilqr_reg, warm start v = v_max), the batch bench.py --workload nominal times, against the C oracle."""
    from differentiable_tube_mpc_b200 import default_dubins_config, ops

    cfg = default_dubins_config()
    B = bench.NOMINAL_BATCH
    prob, ocp, xb, V0 = bench.make_nominal_workload(cfg, 0, B, B)
    env = oracle.env_from_cfg(cfg)
    nom, _ = oracle.ocps_from_cfg(cfg, "paper")
    nom.angle_wrap = 1
    nom.reg = float(cfg["system"]["ilqr_reg"])
    x0 = np.concatenate([xb, np.array([oracle.b0(env, nom, v) for v in xb])[:, None]], 1)
    r = ops.ilqr_solve_batched(prob.system, ocp, cu(x0), cu(V0), cu(prob.theta_nom0), gains=True, trace=True)
    torch.cuda.synchronize()
    X, V, it, ai = r.X.cpu().numpy(), r.V.cpu().numpy(), r.iters.cpu().numpy(), r.alpha_idx.cpu().numpy()
    assert int(r.status.sum()) == 0
    # The run_nominal.py warm start (v = v_max on the whole horizon) drives about a third of random poses THROUGH an
    # obstacle: barrier states ~1e11, costs ~1e23, gains that amplify the last bit -- the first line searches of such an
    # instance are chaotic for any implementation (the reference included, tests/test_gpu_parity.py says the same of
    # the reference's own goldens).  They are solved and must stay finite, but only the TAME instances (every accepted
    # cost of the oracle below 1e12) are compared; their share is printed and bounded.
    n_tame, n_same, n_tie, worst = 0, 0, 0, 0.0
    for i in range(B):
        w = oracle.ilqr_solve(env, nom, x0[i], V0[i])
        if not np.nanmax(w["costs"][: w["iters"]]) < 1e12:
            continue
        n_tame += 1
        same = w["iters"] == it[i] and np.array_equal(w["alpha_idx"][: it[i]], ai[i][: it[i]])
        if not same:    # a roundoff tie in the oracle's own costs (e.g. the saturated problem: v = v_max all along)
            n = min(w["iters"], it[i])
            rel_m = np.min(np.abs(w["margins"][:n]) / np.maximum(np.abs(w["costs"][:n]), 1e-300)) if n else 0.0
            d = np.abs(np.diff(w["costs"][: w["iters"]]))
            near = d.size > 0 and np.min(np.abs(d - nom.tol)) < 1e-9 * abs(w["costs"][0])
            pert = oracle.ilqr_solve(env, nom, x0[i] * (1 + 1e-14), V0[i])
            moved = pert["iters"] != w["iters"] or np.max(np.abs(pert["X"] - w["X"])) > 1e-10 * np.max(np.abs(w["X"]))
            # (margins down to a few 1e-9 flip here: one of 2,7xx tame instances, candidates grazing an obstacle)
            assert rel_m < 1e-8 or near or moved, (i, w["iters"], it[i], rel_m)
            n_tie += 1
            continue
        n_same += 1
        e = max(np.max(np.abs(X[i] - w["X"])) / np.max(np.abs(w["X"])), np.max(np.abs(V[i] - w["V"])) / np.max(np.abs(w["V"])))
        if e > 1e-9:   # allowed only if the oracle's own answer moves by more than the tolerance under a 1e-14 perturbation
            pert = oracle.ilqr_solve(env, nom, x0[i] * (1 + 1e-14), V0[i])   # of x0 (rejected candidates near an obstacle
            moved = np.max(np.abs(pert["X"] - w["X"])) / np.max(np.abs(w["X"]))   # leave gains ~1e8 behind: chaotic)
            assert moved > 1e-10, (i, e, moved)
            n_tie += 1
            continue
        worst = max(worst, e)
    print(f"[scale nominal] B={B}: tame {n_tame / B:.1%}; of those decisions identical and within 1e-9 on {n_same / n_tame:.3%} "
          f"(worst {worst:.1e}), roundoff ties / ill-conditioned {n_tie}")
    # measured on B200 (round 2): tame 79.3 %; 84.4 % of the tame instances identical in every decision and within 1e-9
    # (worst 1.3e-10); the other 534 each PROVED a roundoff tie or an oracle answer that moves under a 1e-14 perturbation
    # of x0 (asserted above, instance by instance): trajectories that skirt an obstacle have jacobian entries ~1e8
    assert n_tame >= 0.7 * B and n_same >= 0.8 * n_tame and worst < 1e-9


def test_million_instance_batch_is_indexed_in_64_bits():
    """BASELINE configs[3] at its largest batch (B = 1,048,576) and T = 100: the sensitivity scratch alone is 4.1e9
    elements, past 2^31, and every per-instance offset of the API tensors is past 2^31 BYTES from a few thousand
    instances on.  Size-independent properties on the whole batch, and slices from the far end and from the middle solved
    again on their own must give bitwise the same answers (solutions, sensitivities, gradients)."""
    from differentiable_tube_mpc_b200 import default_dubins_config, ops

    if torch.cuda.mem_get_info()[0] < 130e9:
        pytest.skip("needs ~110 GB of free device memory")
    cfg = default_dubins_config()
    B, T = 1 << 20, 100
    cfg["system"]["horizon_N"] = T
    prob, w = bench.make_workload(cfg, "full", 0, B, B)
    th_n, th_a = cu(w["th_n"]), cu(w["th_a"])
    z2 = torch.zeros(B, 2, dtype=torch.float64, device=DEV)
    x0n = torch.cat([cu(w["xb"]), torch.zeros(B, 1, dtype=torch.float64, device=DEV)], 1)
    x0a = torch.cat([cu(w["x"]), torch.zeros(B, 1, dtype=torch.float64, device=DEV)], 1)
    x0n[:, 3] = ops.primitives_batched(prob.system, prob.nominal, x0n, z2, th_n)["b0"]
    x0a[:, 3] = ops.primitives_batched(prob.system, prob.auxiliary, x0a, z2, th_a)["b0"]
    V0 = torch.zeros(B, T, 2, dtype=torch.float64, device=DEV)
    from differentiable_tube_mpc_b200 import _lib
    assert _lib.load().dtmpc_sens_workspace_bytes(T, B, 8) // 8 > 2 ** 31
    ws = ops.Workspace(DEV)
    rn, ra, sa, sn = ops.tube_step_batched(prob.system, prob.nominal, prob.auxiliary, x0n, x0a, V0, V0, th_n, th_a,
                                           full=True, workspace=ws)
    torch.cuda.synchronize()
    ok = rn.status == 0
    assert float(ok.float().mean()) > 0.999
    lo, hi = prob.system.ctrl.bounds()
    for r in (rn, ra):
        assert bool((r.V >= cu(lo)).all()) and bool((r.V <= cu(hi)).all())
        assert bool((r.iters >= 2).all())
    assert bool(torch.isfinite(sa.g_theta[ok]).all()) and bool(torch.isfinite(sn.g_theta[ok]).all())
    # the far end, the middle and the first tile group on their own (fresh workspace): bitwise the same
    n = 4096
    for start in (B - n, B // 2 - n // 2, 0, B - 33):
        sl = slice(start, min(B, start + n))
        pn, pa, qa, qn = ops.tube_step_batched(prob.system, prob.nominal, prob.auxiliary, x0n[sl], x0a[sl], V0[sl], V0[sl],
                                               th_n[sl], th_a[sl], full=True, workspace=ops.Workspace(DEV))
        for got, want in ((pn.X, rn.X), (pn.V, rn.V), (pa.X, ra.X), (pa.V, ra.V), (pa.iters, ra.iters), (pn.status, rn.status),
                          (qa.delta_X, sa.delta_X), (qa.delta_lambda, sa.delta_lambda), (qa.g_theta, sa.g_theta),
                          (qn.delta_V, sn.delta_V), (qn.g_theta, sn.g_theta), (qa.loss, sa.loss)):
            assert torch.equal(got, want[sl]), start
        # the returned trajectory is the rollout of the returned controls, bit for bit
        assert torch.equal(ops.rollout_batched(prob.system, prob.auxiliary, x0a[sl], ra.V[sl], th_a[sl]), ra.X[sl])
