"""GPU tests of the pieces around the solver that round 2 added: the hand-written fp64 elementary functions, the
in-kernel Philox disturbance, the CUDA-graph closed loop, failure locations, and shared-theta Algorithm 2 across ranks
(two processes on one GPU over gloo: NCCL refuses two ranks on one device; the N-GPU NCCL run is tests/dist_adapt.py)."""
import math
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import make_instances

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cu(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=DEV)


def _ulps(got, want):
    want = np.asarray(want, dtype=np.float64)
    return np.abs(got - want) / np.maximum(np.spacing(np.abs(want)), 5e-324)


def test_elementary_functions_are_within_two_ulp():
    """exp / log / sincos / reciprocal of dtmpc_math.cuh against mpmath (50 digits) on the argument ranges of the path."""
    import mpmath as mp
    from differentiable_tube_mpc_b200 import ops

    mp.mp.dps = 50
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-45, 45, 3000), rng.uniform(-700, 700, 500), rng.uniform(-1e-3, 1e-3, 500),
                        np.array([0.0, 1.0, -1.0, 0.5, 40.0, -40.0, 699.9, math.pi / 4, 1e4, -3e4, 2.0 ** -30])])
    out = ops.math_selftest(cu(x)).cpu().numpy()
    a = np.abs(x)
    f = lambda fn, v: np.array([float(fn(mp.mpf(float(t)))) for t in v])
    tiny = a > 690                                  # below 1e-300 the result only has to vanish against 1
    assert _ulps(out[~tiny, 0], f(lambda t: mp.exp(-t), a[~tiny])).max() <= 2.0
    assert np.all(out[tiny, 0] < 1e-299)
    assert _ulps(out[:, 1], f(mp.log, 1.0 + a)).max() <= 2.0
    frac = a - np.floor(a)
    assert _ulps(out[:, 2], f(mp.log, 1.0 + frac)).max() <= 2.0
    assert np.array_equal(out[:, 2], ops.math_selftest(cu(frac)).cpu().numpy()[:, 1])   # log_1to2 == log_ge1 on [1, 2]
    s, c = f(mp.sin, x), f(mp.cos, x)
    assert np.max(np.abs(out[:, 3] - s)) <= 2.5e-16 and np.max(np.abs(out[:, 4] - c)) <= 2.5e-16
    big = np.abs(s) > 0.1
    assert _ulps(out[big, 3], s[big]).max() <= 2.0
    nz = x != 0
    assert _ulps(out[nz, 5], 1.0 / x[nz]).max() <= 1.0
    # special values propagate the way the IEEE operations would report them
    sp = ops.math_selftest(cu(np.array([np.nan, np.inf, 0.0]))).cpu().numpy()
    assert np.isnan(sp[0]).all() and np.isnan(sp[1, 3]) and np.isnan(sp[1, 4]) and not np.isfinite(sp[2, 5])


def test_in_kernel_disturbance_is_philox_bit_for_bit():
    """plant step with the in-kernel generator == plant step fed the numpy Philox oracle's tensor (fp64 and fp32)."""
    import philox
    from differentiable_tube_mpc_b200 import TubeProblem, default_dubins_config, ops

    cfg = default_dubins_config()
    cfg["system"]["horizon_N"] = 10
    prob = TubeProblem.from_config(cfg, mode="paper")
    B, T = 1000, 10
    x, xb = make_instances(B, 3, cfg)
    low, high = cfg["system"]["disturbance"]["w_low"], cfg["system"]["disturbance"]["w_high"]
    for dtype, npdt in ((torch.float64, np.float64), (torch.float32, np.float32)):
        th_n, th_a = cu(prob.theta_nom0, dtype), cu(prob.theta_aux0, dtype)
        U = torch.rand(B, T, 2, dtype=dtype, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
        step = torch.tensor([7], dtype=torch.int64, device=DEV)
        for first in (0, 123456789012):
            st = []
            for use_rng in (True, False):
                xh = cu(np.concatenate([x, 0.1 * np.ones((B, 1))], 1), dtype)
                xhb = cu(np.concatenate([xb, 0.1 * np.ones((B, 1))], 1), dtype)
                Vn, Va = torch.zeros_like(U), torch.zeros_like(U)
                w = None if use_rng else cu(philox.disturbance(99, first, B, 7, low, high, dtype=npdt), dtype)
                rng = ops.DisturbanceSpec(low=tuple(low), high=tuple(high), seed=99, step=step, first_instance=first)
                ops.plant_step_batched(prob.system, prob.nominal, prob.auxiliary, U, U, th_n, th_a, w, xh, xhb, Vn, Va,
                                       rng=rng if use_rng else None)
                st.append((xh.clone(), xhb.clone(), Vn.clone()))
            assert torch.equal(st[0][0], st[1][0]) and torch.equal(st[0][1], st[1][1]) and torch.equal(st[0][2], st[1][2])
            assert torch.equal(st[0][2][:, :-1], U[:, 1:]) and torch.equal(st[0][2][:, -1], U[:, -1])   # warm-start shift


@pytest.mark.parametrize("adapt_nominal", [False, True])
def test_graphed_closed_loop_is_the_plain_closed_loop(adapt_nominal):
    """One captured step replayed H times == H plain-launch steps, bit for bit (states, logs, adapted parameters), and a
    tube does not depend on the batch or the shard it runs in (in-kernel disturbance keyed by the global index)."""
    from differentiable_tube_mpc_b200 import default_dubins_config
    from differentiable_tube_mpc_b200.closed_loop import ClosedLoop

    cfg = default_dubins_config()
    cfg["system"]["horizon_N"] = 20
    cfg["use_float64"] = True
    cfg["adaptation"]["adapt_nominal"] = adapt_nominal
    cfg["adaptation"]["lr_eta"] = 0.02 if not adapt_nominal else 1e-3
    H, B = 5, 192
    x0 = cu(make_instances(B, 21, cfg)[0])
    runs = {}
    for graph in (False, True):
        loop = ClosedLoop(cfg, device=DEV, batch=B, n_log=B, x0=x0, graph=graph)
        runs[graph] = (loop.run(H), loop.theta_a.clone(), loop.theta_n.clone(), loop.xh.clone())
        assert (loop._graph is not None) == graph
    for k, v in runs[False][0].items():
        assert torch.equal(v, runs[True][0][k]), k
    for a, b in zip(runs[False][1:], runs[True][1:]):
        assert torch.equal(a, b)
    assert not torch.equal(runs[True][0]["Qa"][0], runs[True][0]["Qa"][-1])          # the parameters did adapt
    assert bool(torch.isfinite(runs[True][0]["loss"]).all())
    # shard invariance with frozen parameters: the second half of the batch on its own, keyed by its global offset
    cfg["adaptation"]["lr_eta"] = 0.0
    whole = ClosedLoop(cfg, device=DEV, batch=B, n_log=B, x0=x0).run(H)
    half = ClosedLoop(cfg, device=DEV, batch=B // 2, n_log=B // 2, x0=x0[B // 2:], first_instance=B // 2).run(H)
    for k in ("x_real", "u_real", "x_bar", "b_real", "loss"):
        assert torch.equal(whole[k][:, B // 2:], half[k]), k


def test_failure_location_is_reported(oracle):
    """status + (iteration, step) of the first non-finite value, and the reference's exception type with that location in
    the message (core/ddp.py:138-159 prints `[NUMERIC-FAIL] ... non-finite ...` before raising FloatingPointError)."""
    from differentiable_tube_mpc_b200 import TubeProblem, default_dubins_config, ops
    from differentiable_tube_mpc_b200.compat import DubinsDBaSDynamics  # noqa: F401  (import check only)

    cfg = default_dubins_config()
    cfg["system"]["horizon_N"] = 12
    prob = TubeProblem.from_config(cfg, mode="paper")
    B, T = 40, 12
    x, _ = make_instances(B, 8, cfg)
    xh = np.concatenate([x, 0.05 * np.ones((B, 1))], 1)
    xh[5, 0] = np.nan                       # poisoned start state: the very first backward pass breaks at its first step
    th = np.tile(np.asarray(prob.theta_nom0), (B, 1))
    th[9, 3:5] = 0.0                        # R = 0 with reg = 0 below: Q_uu is singular where the barrier is flat
    from dataclasses import replace
    ocp = replace(prob.nominal, ilqr=replace(prob.nominal.ilqr, reg=0.0))
    r = ops.ilqr_solve_batched(prob.system, ocp, cu(xh), cu(np.zeros((B, T, 2))), cu(th))
    st, fa = r.status.cpu().numpy(), r.fail_at.cpu().numpy()
    assert st[5] != 0 and tuple(fa[5]) == (0, T - 1)
    clean = st == 0
    assert clean.sum() >= B - 2 and np.all(fa[clean] == -1)
    if st[9] != 0:
        assert fa[9, 0] >= 0 and -1 <= fa[9, 1] < T
    msg = ops.failure_message("iLQR-nominal", r)
    assert "[NUMERIC-FAIL] iLQR-nominal" in msg and "instance 5" in msg and f"k={T - 1}" in msg


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", ["paper", "full"])
def test_shared_theta_adaptation_two_ranks_equal_one_rank(tmp_path, mode):
    """BASELINE configs[4] at test size: data-parallel Algorithm 2 with SHARED theta.  Two ranks x B/2 tubes (gradient
    summed over the ranks in a fixed order, identical update on every rank) against one rank x B tubes on the same
    global batch: the per-tube trajectories are bitwise equal at step 0, the summed gradients agree to the reduction
    order (1e-12), the updated parameters to 1e-12, and the two ranks hold bitwise the same parameters."""
    out = tmp_path / "dist"
    out.mkdir()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), DTMPC_DIST_BACKEND="gloo")
    base = [sys.executable, os.path.join(ROOT, "tests", "dist_adapt.py"), "--mode", mode, "--batch", "512", "--steps", "3",
            "--horizon", "20", "--out", str(out)]
    subprocess.run(base + ["--world", "1"], check=True, env=dict(env, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0"), timeout=600)
    procs = [subprocess.Popen(base + ["--world", "2"], env=dict(env, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK="0"))
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    one = np.load(out / "w1_r0.npz")
    two = [np.load(out / f"w2_r{r}.npz") for r in range(2)]
    assert np.array_equal(two[0]["theta_a"], two[1]["theta_a"]) and np.array_equal(two[0]["theta_n"], two[1]["theta_n"])
    assert np.array_equal(two[0]["g_hist"], two[1]["g_hist"])
    # step 0 runs on identical parameters: every tube's first solve is bitwise the single-rank one
    u0 = np.concatenate([two[0]["u_real"][0], two[1]["u_real"][0]])
    assert np.array_equal(u0, one["u_real"][0])
    scale = np.abs(one["g_hist"]).max(axis=1, keepdims=True)
    assert np.max(np.abs(two[0]["g_hist"] - one["g_hist"]) / scale) < 1e-12
    for k in ("theta_a", "theta_n"):
        assert np.max(np.abs(two[0][k] - one[k])) <= 1e-12 * max(1.0, np.abs(one[k]).max()), k
    assert not np.array_equal(one["theta_a0"], one["theta_a"])    # the parameters moved


@pytest.mark.parametrize("mode", ["paper", "full"])
def test_graphed_batch_closed_loop_step_by_step_against_the_oracle(oracle, mode):
    """BASELINE configs[0] / configs[4] end to end at batch scale, paper and full mode, under CUDA-graph replay.  Before every
    replayed step the loop's device state is read back; the same step is then assembled from the CPU oracle
    (`tube_solve_batch`, a numpy restatement of the shared-theta momentum / projection update of reference
    core/tube_mpc.py:974-984 (paper) / :187-249 (full: chain rule to the raw parameters, momentum, projection), `fhat` with the closed-loop DBaS config of :609 / :614, the numpy Philox) and every link
    is compared: the two solves and the loss per tube, the batch-summed gradient, the updated parameters, and the
    complete successor state (disturbed plant state, propagated barrier states, nominal state, warm-start shifts).
    Warm-started solves of converged tubes are roundoff ties in any implementation (all candidate
    costs equal to a few ulp), so the loop is re-synchronised from the device state at every step rather than
    free-running: a tie-broken tube would otherwise move the SHARED parameters of all tubes."""
    import philox
    from helpers import per_instance_rel
    from differentiable_tube_mpc_b200 import default_dubins_config
    from differentiable_tube_mpc_b200.closed_loop import ClosedLoop

    cfg = default_dubins_config()
    T, B, H = 30, 96, 6
    cfg["system"]["horizon_N"] = T
    cfg["use_float64"] = True
    full = mode == "full"
    cfg["adaptation"]["adapt_nominal"] = full
    cfg["adaptation"]["lr_eta"] = eta = (1e-3 if full else 0.05) / B
    cfg["seed"] = 5
    mom = float(cfg["adaptation"].get("momentum", 0.9))
    x0, _ = make_instances(B, 77, cfg)
    loop = ClosedLoop(cfg, device=DEV, batch=B, n_log=B, x0=cu(x0))
    assert loop.paper == (not full) and loop.use_graph
    loop.reserve(H)
    env = oracle.env_from_cfg(cfg)
    nom, aux = oracle.ocps_from_cfg(cfg, mode)
    low, high = cfg["system"]["disturbance"]["w_low"], cfg["system"]["disturbance"]["w_high"]
    adapts = np.array([1.0] * 5 + [0.0] * 3 + [1.0, 0.0, 0.0, 0.0])
    lo = np.array([0.0] * 3 + [1e-4] * 2 + [0.0] * 7)
    hi = np.array([np.inf] * 8 + [1.0, 0.0, 0.0, 0.0])
    npy = lambda t: t.detach().cpu().numpy().copy()
    # full mode: projection boxes of the raw parameters by group (reference `_project`), chain factors, constraint maps
    p_lo = np.array([0.0] * 3 + [1e-4] * 2 + [0.0] * 3 + [0.0, 0.0, -1.0, 0.0])
    p_hi = np.array([np.inf] * 3 + [1e4] * 2 + [np.inf] * 3 + [1.0, 1.0, 1.0, 2.0])
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    sp = lambda x: np.logaddexp(0.0, x)

    def raw_update(raw, vel_, g_sum_, with_tight):
        mask = np.array([1.0] * 11 + [1.0 if with_tight else 0.0])
        f = sig(raw)
        f[10] = 1.0 - np.tanh(raw[10]) ** 2
        v = mom * vel_ + g_sum_ * f * mask
        r = raw - eta * v
        if bool(cfg["adaptation"].get("project_params", False)):
            r = np.minimum(np.maximum(r, p_lo), p_hi)
        th = sp(r)
        th[9] += 1e-6
        th[10] = np.tanh(r[10])
        return r, v, th * mask
    worst, n_same, n_wild = {}, 0, 0

    def note(k, e):
        worst[k] = max(worst.get(k, 0.0), float(e))

    for t in range(H):
        xh, xhb, Vn, Va = npy(loop.xh), npy(loop.xh_bar), npy(loop.Vn_ws), npy(loop.Va_ws)
        th_n, th_a = npy(loop.theta_n), npy(loop.theta_a)
        if full:
            raw_a, vel_a, raw_n, vel_n = npy(loop.theta.raw), npy(loop.theta.vel), npy(loop.theta_bar.raw), npy(loop.theta_bar.vel)
        else:
            vel = npy(loop.adapter.vel)
        loop.step()
        st = loop.last
        assert loop._graph is not None and int(loop.t_dev.item()) == t + 1
        Un, Ua = npy(st.nominal.V), npy(st.auxiliary.V)
        # (1) the solves and the loss, tube by tube
        nom.set_theta(th_n); aux.set_theta(th_a)
        w = oracle.tube_solve_batch(env, nom, aux, xhb, xh, Vn, Va, mode=mode, n_threads=8)
        assert not w["status"].any() and not npy(st.status).any()
        e = np.maximum(per_instance_rel(Un, w["Un"], floor=1e-3), per_instance_rel(Ua, w["Ua"], floor=1e-3))
        same = (npy(st.nominal.iters) == w["iters_n"]) & (npy(st.auxiliary.iters) == w["iters_a"]) & (e < 1e-8)
        w2 = oracle.tube_solve_batch(env, nom, aux, xhb * (1 + 1e-14), xh * (1 + 1e-14), Vn, Va, mode=mode, n_threads=8)
        moved = np.maximum(per_instance_rel(w2["Un"], w["Un"], floor=1e-3), per_instance_rel(w2["Ua"], w["Ua"], floor=1e-3))
        for i in np.nonzero(~same)[0]:
            # must be a roundoff tie in the oracle's own candidate costs, or a solve whose candidates leave the safe set
            # (the shifted warm start of a disturbed tube can cut an obstacle: costs ~1e18), where the oracle's own answer
            # moves by O(1) under a 1e-14 perturbation of the state: chaotic in any implementation
            tie, wild = False, bool(moved[i] > 1e-9)
            for tag in ("n", "a"):
                n = int(w[f"iters_{tag}"][i])
                c, mg = w[f"costs_{tag}"][i][:n], w[f"marg_{tag}"][i][:n]
                tol = nom.tol if tag == "n" else aux.tol
                tie |= bool(n and np.nanmin(np.abs(mg) / np.maximum(np.abs(c), 1e-300)) < 1e-10)
                tie |= bool(n > 1 and np.min(np.abs(np.abs(np.diff(c)) - tol)) < 1e-9 * abs(c[0]))
                wild |= bool(n and max(np.nanmax(np.abs(c)), np.nanmax(np.abs(mg))) > 1e12)
            assert tie or wild, (t, i, e[i])
            n_wild += int(wild)
        n_same += int(same.sum())
        note("U", e[same].max(initial=0.0))
        note("loss", per_instance_rel(npy(st.loss)[same, None], w["loss"].reshape(B, 1)[same]).max(initial=0.0))
        g_dev = npy(st.sens_aux.g_theta)
        note("g", (np.abs(g_dev[same] - w["g_aux"][same]).max(axis=1) / np.maximum(np.abs(w["g_aux"][same]).max(axis=1), 1e-9)).max(initial=0.0))
        # (2) batch sum and the shared-parameter update (from the device's own per-tube gradients: independent of ties)
        g_sum = npy(st.g_aux_sum)[:12]
        note("g_sum", np.abs(g_sum - g_dev.sum(axis=0)).max() / np.abs(g_dev).sum(axis=0).max())
        if full:
            gn_dev = npy(st.sens_nom.g_theta)
            ok_n = same & (moved < 1e-11)
            note("g_nom", (np.abs(gn_dev[ok_n] - w["g_nom"][ok_n]).max(axis=1) / np.maximum(
                np.abs(w["g_nom"][ok_n]).max(axis=1), 1e-6 * np.abs(w["dLn"][ok_n]).reshape(int(ok_n.sum()), -1).max(axis=1) + 1e-300)).max(initial=0.0))
            gn_sum = npy(st.g_nom_sum)
            note("g_sum", np.abs(gn_sum - gn_dev.sum(axis=0)).max() / np.abs(gn_dev).sum(axis=0).max())
            ra_, va_, pa_ = raw_update(raw_a, vel_a, g_sum, False)
            rn_, vn_, pn_ = raw_update(raw_n, vel_n, gn_sum, True)
            note("theta", max(np.abs(npy(loop.theta_a) - pa_).max() / np.abs(pa_).max(), np.abs(npy(loop.theta_n) - pn_).max() / np.abs(pn_).max()))
            note("raw", max(np.abs(npy(loop.theta.raw) - ra_).max() / np.abs(ra_).max(), np.abs(npy(loop.theta_bar.raw) - rn_).max() / np.abs(rn_).max()))
            note("vel", max(np.abs(npy(loop.theta.vel) - va_).max() / max(np.abs(va_).max(), 1e-300),
                            np.abs(npy(loop.theta_bar.vel) - vn_).max() / max(np.abs(vn_).max(), 1e-300)))
        else:
            g = g_sum * adapts
            g[0:3] += g_sum[5:8]
            vel = mom * vel + g
            p = np.minimum(np.maximum(th_a - eta * vel, lo), hi)
            p[5:8] = p[0:3]
            note("theta", np.abs(npy(loop.theta_a) - p).max() / np.abs(p).max())
            note("vel", np.abs(npy(loop.adapter.vel) - vel).max() / max(np.abs(vel).max(), 1e-300))
        # (3) successor state of EVERY tube from the device's own controls: plant + disturbance, barrier states, nominal
        # propagation (with the UPDATED parameters), warm-start shift
        pa, pn = aux.copy(), nom.copy()
        pa.set_theta(npy(loop.theta_a)); pn.set_theta(npy(loop.theta_n))
        pa.eps_dyn, pn.eps_dyn = pa.eps_jac, pn.eps_jac
        nxt = np.array([oracle.fhat(env, pa, xh[i], Ua[i, 0]) for i in range(B)])
        nxt[:, :3] += philox.disturbance(5, 0, B, t, low, high)
        note("x_next", per_instance_rel(npy(loop.xh), nxt).max())
        note("xbar_next", per_instance_rel(npy(loop.xh_bar), np.array([oracle.fhat(env, pn, xhb[i], Un[i, 0]) for i in range(B)])).max())
        assert np.array_equal(npy(loop.Vn_ws), np.concatenate([Un[:, 1:], Un[:, -1:]], axis=1))
        assert np.array_equal(npy(loop.Va_ws), np.concatenate([Ua[:, 1:], Ua[:, -1:]], axis=1))
    print(f"[closed loop vs oracle, {mode}] {H} replayed steps x {B} tubes: solves identical on {n_same}/{H * B} ({n_wild} chaotic: "
          "candidates outside the safe set, the rest roundoff ties); worst relative errors:", ", ".join(f"{k} {v:.1e}" for k, v in worst.items()))
    assert n_same >= 0.85 * H * B       # measured on B200: 509 of 576
    assert float(np.abs(npy(loop.theta_a)[:9] - np.asarray(loop.prob.theta_aux0)[:9]).max()) > 1e-7   # the parameters did move
    for k, v in worst.items():
        assert v < (1e-12 if k in ("g_sum", "theta", "raw", "vel", "x_next", "xbar_next") else 1e-8), (k, v)
