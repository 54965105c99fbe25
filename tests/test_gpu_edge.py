"""GPU parity on the shapes the slot/tile/segment machinery could get wrong: horizons shorter than the number of commit
segments, batches that are not whole tiles of 32, a single instance, step-size lists with one entry or with only
alpha = 0, every kind of obstacle count (none, compile-time instantiated, generic).  Checked against the CPU oracle."""

import numpy as np
import pytest
import torch

from helpers import make_instances, per_instance_rel

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)


def _setup(oracle, T, B, seed, n_obs=None, alphas=None, max_iter=None):
    from differentiable_tube_mpc_b200 import TubeProblem, default_dubins_config

    cfg = default_dubins_config()
    cfg["system"]["horizon_N"] = int(T)
    if n_obs is not None:
        obs = list(cfg["environment"]["obstacles"])
        while len(obs) < n_obs:      # extra circles far from every start state: more terms, same geometry nearby
            obs.append({"center": [40.0 + 3.0 * len(obs), -30.0], "radius": 1.0})
        cfg["environment"]["obstacles"] = obs[:n_obs]
    if alphas is not None:
        cfg["system"]["line_search_alphas"] = list(alphas)
    if max_iter is not None:
        cfg["system"]["nominal_max_iter"] = cfg["system"]["aux_max_iter"] = int(max_iter)
    prob = TubeProblem.from_config(cfg, mode="paper")
    env = oracle.env_from_cfg(cfg)
    nom, aux = oracle.ocps_from_cfg(cfg, "paper")
    ref_cfg = default_dubins_config()          # start states are drawn against the full obstacle field
    x, xb = make_instances(B, seed, ref_cfg)
    b0n = np.array([oracle.b0(env, nom, v) for v in xb])
    b0a = np.array([oracle.b0(env, aux, v) for v in x])
    x0n = np.concatenate([xb, b0n[:, None]], axis=1)
    x0a = np.concatenate([x, b0a[:, None]], axis=1)
    return prob, env, nom, aux, x0n, x0a


def _check(oracle, prob, env, nom, aux, x0n, x0a, T):
    from differentiable_tube_mpc_b200 import ops

    B = x0n.shape[0]
    V0 = np.zeros((B, T, 2))
    want = oracle.tube_solve_batch(env, nom, aux, x0n, x0a, V0, V0, mode="paper", n_threads=4)
    rn, ra = ops.tube_solve_batched(prob.system, prob.nominal, prob.auxiliary, cu(x0n), cu(x0a), cu(V0), cu(V0),
                                    cu(prob.theta_nom0), cu(prob.theta_aux0), gains=True, trace=True)
    torch.cuda.synchronize()
    assert np.array_equal(rn.status.cpu().numpy() != 0, want["status"] != 0)
    ok = want["status"] == 0
    same = ok & (want["iters_n"] == rn.iters.cpu().numpy()) & (want["iters_a"] == ra.iters.cpu().numpy())
    for i in np.nonzero(same)[0]:
        same[i] &= np.array_equal(want["aidx_n"][i][: want["iters_n"][i]], rn.alpha_idx.cpu().numpy()[i][: want["iters_n"][i]])
        same[i] &= np.array_equal(want["aidx_a"][i][: want["iters_a"][i]], ra.alpha_idx.cpu().numpy()[i][: want["iters_a"][i]])
    # every instance whose decisions differ must be a roundoff tie of the oracle's own (App. B Q12): a line-search margin
    # or a tolerance test within 1e-10 relative of flipping
    def tied(i, tag, tol):
        n = int(want[f"iters_{tag}"][i])
        c, mg = want[f"costs_{tag}"][i], want[f"marg_{tag}"][i]
        if not n:
            return False
        d = np.abs(np.diff(c[:n]))
        return bool(np.nanmin(np.abs(mg[:n]) / np.maximum(np.abs(c[:n]), 1e-300)) < 1e-10) or bool(
            d.size and np.min(np.abs(d - tol)) < 1e-9 * abs(c[0]))

    for i in np.nonzero(ok & ~same)[0]:
        nom_same = want["iters_n"][i] == rn.iters[i].item() and np.array_equal(
            want["aidx_n"][i][: want["iters_n"][i]], rn.alpha_idx.cpu().numpy()[i][: want["iters_n"][i]])
        # a tie-broken nominal solve changes the auxiliary problem's reference, so only the first differing layer is judged
        assert tied(i, "n", nom.tol) if not nom_same else tied(i, "a", aux.tol), (
            i, "decision mismatch without a tie in the oracle's own margins")
    # iterates that left the safe set (barrier costs ~1e16 and above: a full step without a smaller candidate to fall
    # back on) amplify roundoff without bound in any implementation; they are compared in the median only
    wild = (np.nan_to_num(want["costs_n"], nan=0.0).max(axis=1) > 1e12) | (np.nan_to_num(want["costs_a"], nan=0.0).max(axis=1) > 1e12)
    worst = 0.0
    for got, key in ((rn.X, "Xn"), (rn.V, "Un"), (ra.X, "Xa"), (ra.V, "Ua"), (rn.K, "Kn"), (ra.K, "Ka")):
        e = per_instance_rel(got.cpu().numpy()[same], want[key][same], floor=1e-6)
        assert e.size == 0 or np.median(e) < 1e-12, (key, np.median(e))
        worst = max(worst, e[~wild[same]].max(initial=0.0))
    print(f"[edge T={T} B={B}] decisions equal on {int(same.sum())}/{int(ok.sum())}; worst rel err where equal "
          f"{worst:.1e} ({int((wild & same).sum())} outside the safe set)")
    assert same.sum() >= 0.95 * max(1, ok.sum()), (same.sum(), ok.sum())     # measured: all, in every case of this file
    assert (wild & same).sum() <= 0.1 * B
    assert worst < 1e-9, worst
    # whatever the decisions: the returned trajectory is the rollout of the returned controls, bit for bit
    for r, ocp, th in ((rn, prob.nominal, prob.theta_nom0), (ra, prob.auxiliary, prob.theta_aux0)):
        x0 = cu(x0n if r is rn else x0a)
        X = ops.rollout_batched(prob.system, ocp, x0, r.V, cu(th))
        assert torch.equal(X, r.X)


@pytest.mark.parametrize("T,B", [(1, 33), (2, 5), (3, 64), (4, 1), (7, 95), (11, 32)])
def test_short_horizons_and_ragged_batches(oracle, T, B):
    prob, env, nom, aux, x0n, x0a = _setup(oracle, T, B, seed=100 + T)
    _check(oracle, prob, env, nom, aux, x0n, x0a, T)


@pytest.mark.parametrize("T,B", [(100, 24), (200, 12), (53, 40)])
def test_long_and_odd_horizons(oracle, T, B):
    """BASELINE config #4 horizons: segments of 20 / 40 steps, and a horizon the five segments do not divide."""
    prob, env, nom, aux, x0n, x0a = _setup(oracle, T, B, seed=200 + T)
    _check(oracle, prob, env, nom, aux, x0n, x0a, T)


@pytest.mark.parametrize("alphas", [(1.0,), (0.0,), (0.3, 0.0), (0.0, 1.0, 0.5)])
def test_degenerate_step_size_lists(oracle, alphas):
    prob, env, nom, aux, x0n, x0a = _setup(oracle, 9, 40, seed=7, alphas=alphas, max_iter=4)
    _check(oracle, prob, env, nom, aux, x0n, x0a, 9)


@pytest.mark.parametrize("n_obs", [0, 1, 3, 8, 10])
def test_obstacle_counts(oracle, n_obs):
    """1..8 circles run kernels with the count fixed at compile time, 0 and 10 the generic ones."""
    prob, env, nom, aux, x0n, x0a = _setup(oracle, 12, 48, seed=31 + n_obs, n_obs=n_obs)
    _check(oracle, prob, env, nom, aux, x0n, x0a, 12)


def test_empty_batch_and_bad_arguments():
    from differentiable_tube_mpc_b200 import TubeProblem, default_dubins_config, ops

    cfg = default_dubins_config()
    cfg["system"]["horizon_N"] = 5
    prob = TubeProblem.from_config(cfg, mode="paper")
    z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=DEV)
    rn, ra = ops.tube_solve_batched(prob.system, prob.nominal, prob.auxiliary, z(0, 4), z(0, 4), z(0, 5, 2), z(0, 5, 2),
                                    cu(prob.theta_nom0), cu(prob.theta_aux0))
    assert rn.X.shape == (0, 6, 4) and ra.V.shape == (0, 5, 2)
    with pytest.raises((ValueError, TypeError, RuntimeError)):
        ops.tube_solve_batched(prob.system, prob.nominal, prob.auxiliary, z(3, 4).cpu(), z(3, 4), z(3, 5, 2), z(3, 5, 2),
                               cu(prob.theta_nom0), cu(prob.theta_aux0))
    with pytest.raises((ValueError, TypeError, RuntimeError)):
        ops.tube_solve_batched(prob.system, prob.nominal, prob.auxiliary, z(3, 4), z(3, 4), z(3, 4, 2), z(3, 5, 2),
                               cu(prob.theta_nom0), cu(prob.theta_aux0))


def test_closed_form_weight_gradients_cross_check(oracle):
    """core/gradients.py:27-87 (defined but unused by the reference) restated in gradients.py, against the weight
    entries of the fused CUDA accumulation, for both OCPs of a full-mode step."""
    from differentiable_tube_mpc_b200 import TubeProblem, default_dubins_config, ops
    from differentiable_tube_mpc_b200.gradients import grads_auxiliary_from_deltas, grads_nominal_from_deltas

    cfg = default_dubins_config()
    cfg["system"]["horizon_N"] = 15
    prob = TubeProblem.from_config(cfg, mode="full")
    B, T = 96, 15
    x, xb = make_instances(B, 77, cfg)
    th_bar, th = prob.make_thetas(DEV, torch.float64)
    thn = th_bar.constrained().detach().expand(B, 12).contiguous()
    tha = th.constrained().detach().expand(B, 12).contiguous()
    z = torch.zeros(B, 1, dtype=torch.float64, device=DEV)
    u = torch.zeros(B, 2, dtype=torch.float64, device=DEV)
    b0n = ops.primitives_batched(prob.system, prob.nominal, torch.cat([cu(xb), z], 1), u, thn)["b0"]
    b0a = ops.primitives_batched(prob.system, prob.auxiliary, torch.cat([cu(x), z], 1), u, tha)["b0"]
    V0 = torch.zeros(B, T, 2, dtype=torch.float64, device=DEV)
    rn, ra, sa, sn = ops.tube_step_batched(prob.system, prob.nominal, prob.auxiliary, torch.cat([cu(xb), b0n.view(-1, 1)], 1),
                                           torch.cat([cu(x), b0a.view(-1, 1)], 1), V0, V0, thn, tha, full=True)
    ga = grads_auxiliary_from_deltas(X_hat=ra.X, U=ra.V, dX=sa.delta_X, dU=sa.delta_V, X_ref=rn.X, U_ref=rn.V)
    gn = grads_nominal_from_deltas(X_hat=rn.X, U=rn.V, dX=sn.delta_X, dU=sn.delta_V,
                                   target_x=cu(np.asarray(prob.system.y_target)))
    for g, s, r in ((ga, sa, ra), (gn, sn, rn)):
        k = s.g_theta
        scale = k[:, :9].abs().max(dim=1).values.clamp_min(1e-30)
        qb_T = 2.0 * r.X[:, -1, 3] * s.delta_X[:, -1, 3]
        for got, want in ((g.Q_diag, k[:, 0:3]), (g.R_diag, k[:, 3:5]), (g.Qf_diag, k[:, 5:8]),
                          ((g.q_b + qb_T).view(-1, 1), k[:, 8:9])):
            assert float(((got - want).abs().max(dim=1).values / scale).max()) < 1e-12
