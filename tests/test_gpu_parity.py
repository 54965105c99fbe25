"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the reference-generated goldens.

Tolerance (BASELINE.json north_star): 1e-9 relative (tensor infinity norm) in fp64; discrete decisions
(iteration counts, accepted line-search step) identical whenever the decision margin is not a roundoff tie
(SURVEY App. B Q12).  Instances whose own solution moves by more than the tolerance under a 1e-14 relative
perturbation of x0 (trajectories that end inside an obstacle: jacobian entries ~1e11) are ill-conditioned for
ANY implementation, including the reference on another BLAS; they are compared at 10 x their measured
sensitivity instead and their share is bounded.
"""
import numpy as np
import pytest
import torch

from helpers import golden_step_conditioning, load_golden, make_instances, per_instance_rel, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-9
DEV = "cuda"


def cu(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=DEV)


def problem(mode="paper", T=None):
    from differentiable_tube_mpc_b200 import TubeProblem, default_dubins_config

    cfg = default_dubins_config()
    if T is not None:
        cfg["system"]["horizon_N"] = int(T)
    return cfg, TubeProblem.from_config(cfg, mode=mode)


def oracle_problem(oracle, cfg, mode):
    env = oracle.env_from_cfg(cfg)
    nom, aux = oracle.ocps_from_cfg(cfg, mode)
    return env, nom, aux


def b0_of(oracle, env, ocp, x):
    return np.array([oracle.b0(env, ocp, xi) for xi in x])


# ---------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------
def test_primitives_match_reference_goldens():
    from differentiable_tube_mpc_b200 import ops
    from differentiable_tube_mpc_b200.spec import CircleObstacle
    from dataclasses import replace

    g = load_golden("primitives")
    cfg, prob = problem("paper")
    pts = g["pts"]
    n = len(pts)
    xh = cu(np.concatenate([pts, np.zeros((n, 1))], axis=1))
    u = cu(np.zeros((n, 2)))
    theta = cu(prob.theta_nom0)
    for nm, nobs in (("M5", 5), ("M1", 1), ("M0", 0)):
        for agg in ("smoothmin", "min"):
            sysd = replace(prob.system, obstacles=prob.system.obstacles[:nobs], obs_agg=agg)
            out = ops.primitives_batched(sysd, prob.nominal, xh, u, theta)
            key = "smooth" if agg == "smoothmin" else "min"
            assert rel_err(out["h"].cpu().numpy(), g[f"h_{key}_{nm}"]) < 1e-13, (nm, agg)
            assert rel_err(out["grad_h"].cpu().numpy(), g[f"gh_{key}_{nm}"], floor=1.0) < 1e-13, (nm, agg)
    sysd = replace(prob.system, obstacles=prob.system.obstacles[:1], obs_agg="single")
    out = ops.primitives_batched(sysd, prob.nominal, xh, u, theta)
    assert rel_err(out["h"].cpu().numpy(), g["h_single"]) < 1e-14
    assert rel_err(out["grad_h"].cpu().numpy(), g["gh_single"]) < 1e-14

    # dbas_step + augmented jacobian: gamma in {0, +-0.8}, alpha in {0, 0.5}, both aggregations, both branches
    rows = g["step_jac"]
    for agg_id, agg in ((0.0, "smoothmin"), (1.0, "min")):
        sel = rows[rows[:, 0] == agg_id]
        sysd = replace(prob.system, obs_agg=agg)
        th = np.tile(np.asarray(prob.theta_nom0), (len(sel), 1))
        th[:, 9], th[:, 10] = sel[:, 1], sel[:, 2]
        out = ops.primitives_batched(sysd, prob.nominal, cu(sel[:, 3:7]), cu(sel[:, 7:9]), cu(th))
        assert rel_err(out["x_next"].cpu().numpy(), sel[:, 9:13]) < 1e-12
        A, Bm = out["A"].cpu().numpy().reshape(-1, 16), out["B"].cpu().numpy().reshape(-1, 8)
        assert np.max(per_instance_rel(A, sel[:, 13:29])) < 1e-12
        assert np.max(per_instance_rel(Bm, sel[:, 29:37])) < 1e-12

    # control box
    uu = g["ctrl_u"]
    out = ops.primitives_batched(prob.system, prob.nominal, cu(np.zeros((len(uu), 4))), cu(uu), theta)
    assert np.array_equal(out["u_clamped"].cpu().numpy(), g["ctrl_clamp"])
    assert np.array_equal(out["active"].cpu().numpy(), g["ctrl_active"])


def test_barrier_branches_via_b0(oracle):
    """B_alpha on both sides of a = max(alpha, eps) and at zeta == a: with no obstacles h = 1, so
    b0 = B(1 - tight)."""
    from differentiable_tube_mpc_b200 import ops
    from dataclasses import replace

    g = load_golden("primitives")
    cfg, prob = problem("paper")
    sysd = replace(prob.system, obstacles=())
    for eps in (1e-4, 1e-6):
        sel = g["B_eps"] == eps
        z, al = g["B_z"][sel], g["B_alpha"][sel]
        tight = 1.0 - z
        zz = 1.0 - tight  # what the kernel sees
        th = np.tile(np.asarray(prob.theta_nom0), (len(z), 1))
        th[:, 9], th[:, 11] = al, tight
        ocp = replace(prob.nominal, eps_dyn=eps, eps_jac=eps)
        out = ops.primitives_batched(sysd, ocp, cu(np.zeros((len(z), 4))), cu(np.zeros((len(z), 2))), cu(th))
        want = np.array([oracle.B(a, b, eps) for a, b in zip(zz, al)])
        assert rel_err(out["b0"].cpu().numpy(), want) < 1e-14
        exact = zz == z
        assert rel_err(out["b0"].cpu().numpy()[exact], g["B_val"][sel][exact]) < 1e-14


# ---------------------------------------------------------------------------
# solver parity against the oracle on seeded synthetic batches
# ---------------------------------------------------------------------------
def _oracle_sensitivity(oracle, env, nom, aux, x0n, x0a, Vn0, Va0, mode):
    """How far the oracle's own answer moves under a 1e-14 relative perturbation of the initial states."""
    base = oracle.tube_solve_batch(env, nom, aux, x0n, x0a, Vn0, Va0, mode=mode, n_threads=8)
    pert = oracle.tube_solve_batch(env, nom, aux, x0n * (1 + 1e-14), x0a * (1 + 1e-14), Vn0, Va0, mode=mode,
                                   n_threads=8)
    s = np.zeros(x0n.shape[0])
    for k in ("Xn", "Un", "Xa", "Ua"):
        s = np.maximum(s, per_instance_rel(pert[k], base[k]))
    return base, s


def _compare(got, want, sens, what, floor=0.0, max_loose_share=0.1):
    e = per_instance_rel(got, want, floor=floor)
    tol_i = np.maximum(TOL, 10.0 * sens)
    bad = e > tol_i
    assert not bad.any(), f"{what}: {bad.sum()} instances off, worst {e[bad].max():.3e} (tol {tol_i[bad].min():.1e})"
    loose = (tol_i > TOL).mean()
    assert loose <= max_loose_share, f"{what}: {loose:.0%} of the instances needed a conditioning allowance"
    print(f"[parity] {what}: n={e.size} median {np.median(e):.1e} p99 {np.quantile(e, 0.99):.1e} max {e.max():.1e} "
          f"(conditioning allowance used by {loose:.1%})")
    return e


@pytest.mark.parametrize("mode,B,seed", [("paper", 192, 11), ("full", 128, 12)])
def test_tube_solve_sens_grads_vs_oracle(oracle, mode, B, seed):
    from differentiable_tube_mpc_b200 import ops

    cfg, prob = problem(mode)
    env, nom, aux = oracle_problem(oracle, cfg, mode)
    T = prob.nominal.ilqr.horizon
    x, xb = make_instances(B, seed, cfg)
    rng = np.random.default_rng(seed + 1)
    # per-instance parameters: weights x U(0.5, 2)
    th_n = np.tile(nom.theta(), (B, 1))
    th_a = np.tile(aux.theta(), (B, 1))
    th_n[:, :9] *= rng.uniform(0.5, 2.0, size=(B, 9))
    th_a[:, :9] *= rng.uniform(0.5, 2.0, size=(B, 9))
    th_a[:, 8] = np.minimum(th_a[:, 8], 1.0)
    b0n = np.empty(B)
    b0a = np.empty(B)
    for i in range(B):
        nom.set_theta(th_n[i]); aux.set_theta(th_a[i])
        b0n[i], b0a[i] = oracle.b0(env, nom, xb[i]), oracle.b0(env, aux, x[i])
    x0n = np.concatenate([xb, b0n[:, None]], axis=1)
    x0a = np.concatenate([x, b0a[:, None]], axis=1)
    Vn0 = np.zeros((B, T, 2))
    Va0 = np.zeros((B, T, 2))

    want = oracle.tube_solve_batch(env, nom, aux, x0n, x0a, Vn0, Va0, theta_n=th_n, theta_a=th_a, mode=mode, n_threads=8)
    pert = oracle.tube_solve_batch(env, nom, aux, x0n * (1 + 1e-14), x0a * (1 + 1e-14), Vn0, Va0, theta_n=th_n,
                                   theta_a=th_a, mode=mode, n_threads=8)
    sens = np.zeros(B)
    for k in ("Xn", "Un", "Xa", "Ua", "dXa", "dLa"):
        sens = np.maximum(sens, per_instance_rel(pert[k], want[k], floor=1e-6))

    rn, ra = ops.tube_solve_batched(prob.system, prob.nominal, prob.auxiliary, cu(x0n), cu(x0a), cu(Vn0), cu(Va0),
                                    cu(th_n), cu(th_a), gains=True, trace=True)
    torch.cuda.synchronize()
    assert np.array_equal(rn.status.cpu().numpy() != 0, want["status"] != 0)
    ok = want["status"] == 0
    assert ok.mean() > 0.95

    # decisions: identical iteration counts / accepted steps unless the oracle's own margin is a roundoff tie
    for tag, r in (("n", rn), ("a", ra)):
        it_w, it_g = want[f"iters_{tag}"], r.iters.cpu().numpy()
        ai_w, ai_g = want[f"aidx_{tag}"], r.alpha_idx.cpu().numpy()
        marg, costs = want[f"marg_{tag}"], want[f"costs_{tag}"]
        n_tie = 0
        for i in np.nonzero(ok)[0]:
            same = it_w[i] == it_g[i] and np.array_equal(ai_w[i][: it_w[i]], ai_g[i][: it_g[i]])
            if same:
                continue
            n = min(it_w[i], it_g[i])
            rel_m = np.nanmin(np.abs(marg[i][:n]) / np.maximum(np.abs(costs[i][:n]), 1e-300)) if n else 0.0
            d = np.abs(np.diff(costs[i][: it_w[i]]))
            near_tol = d.size > 0 and np.min(np.abs(d - (nom.tol if tag == "n" else aux.tol))) < 1e-9 * np.abs(costs[i][0])
            assert rel_m < 1e-10 or near_tol or sens[i] > 1e-12, (
                f"decision mismatch on a well-separated instance {i} ({tag}): iters {it_w[i]} vs {it_g[i]}, "
                f"alpha {ai_w[i][:it_w[i]]} vs {ai_g[i][:it_g[i]]}, margin {rel_m:.2e}")
            n_tie += 1
        assert n_tie <= 0.1 * B, f"{n_tie} decision ties in {B} instances"

    sel = ok
    _compare(rn.X.cpu().numpy()[sel], want["Xn"][sel], sens[sel], "X_nom")
    _compare(rn.V.cpu().numpy()[sel], want["Un"][sel], sens[sel], "U_nom")
    _compare(ra.X.cpu().numpy()[sel], want["Xa"][sel], sens[sel], "X_aux")
    _compare(ra.V.cpu().numpy()[sel], want["Ua"][sel], sens[sel], "U_aux")
    same_it = sel & (want["iters_n"] == rn.iters.cpu().numpy()) & (want["iters_a"] == ra.iters.cpu().numpy())
    _compare(rn.K.cpu().numpy()[same_it], want["Kn"][same_it], sens[same_it], "K_nom")
    _compare(ra.K.cpu().numpy()[same_it], want["Ka"][same_it], sens[same_it], "K_aux")
    _compare(rn.k.cpu().numpy()[same_it], want["kn"][same_it], sens[same_it], "k_nom", floor=1e-3)
    _compare(ra.k.cpu().numpy()[same_it], want["ka"][same_it], sens[same_it], "k_aux", floor=1e-3)

    # sensitivity + gradients evaluated at the ORACLE's optimum (isolates a5/a6 from solver conditioning)
    Xa, Ua, Xn, Un = cu(want["Xa"]), cu(want["Ua"]), cu(want["Xn"]), cu(want["Un"])
    sa = ops.ddp_sensitivity_batched(prob.system, prob.auxiliary, Xa, Ua, cu(th_a), X_bar=Xn, X_ref=Xn, U_ref=Un,
                                     want_grads=True, want_ref_grads=True)
    z = np.zeros(B)
    scale = np.abs(want["dLa"]).reshape(B, -1).max(axis=1)
    for got, key in ((sa.delta_X, "dXa"), (sa.delta_V, "dVa"), (sa.delta_lambda, "dLa")):
        e = per_instance_rel(got.cpu().numpy()[sel], want[key][sel], floor=1e-9)
        assert e.max() < TOL, (key, e.max())
    assert rel_err(sa.loss.cpu().numpy()[sel], want["loss"][sel]) < 1e-12
    e = np.abs(sa.g_theta.cpu().numpy()[sel] - want["g_aux"][sel]).max(axis=1) / np.maximum(
        np.abs(want["g_aux"][sel]).max(axis=1), 1e-6 * scale[sel])
    assert e.max() < TOL, ("g_aux", e.max())
    assert np.max(per_instance_rel(sa.g_xref.cpu().numpy()[sel], want["gXref"][sel], floor=1e-9)) < TOL
    assert np.max(per_instance_rel(sa.g_uref.cpu().numpy()[sel], want["gUref"][sel], floor=1e-9)) < TOL
    if mode == "full":
        gx = torch.cat([cu(want["gXref"]), torch.zeros(B, T + 1, 1, dtype=torch.float64, device=DEV)], dim=2)
        sn = ops.ddp_sensitivity_batched(prob.system, prob.nominal, Xn, Un, cu(th_n), upper_grad_x=gx,
                                         upper_grad_u=cu(want["gUref"]), want_grads=True)
        for got, key in ((sn.delta_X, "dXn"), (sn.delta_V, "dVn"), (sn.delta_lambda, "dLn")):
            e = per_instance_rel(got.cpu().numpy()[sel], want[key][sel], floor=1e-9)
            assert e.max() < TOL, (key, e.max())
        sc = np.abs(want["dLn"]).reshape(B, -1).max(axis=1)
        e = np.abs(sn.g_theta.cpu().numpy()[sel] - want["g_nom"][sel]).max(axis=1) / np.maximum(
            np.abs(want["g_nom"][sel]).max(axis=1), 1e-6 * sc[sel])
        assert e.max() < TOL, ("g_nom", e.max())


# ---------------------------------------------------------------------------
# directly against what the reference computed (tests/golden, made by oracle/make_golden.py)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["closedloop_paper", "random_paper", "paper_T20", "closedloop_full", "random_full",
                                  "full_gamma", "closedloop_paper_rand", "closedloop_full_rand"])
def test_against_reference_goldens(oracle, name):
    from differentiable_tube_mpc_b200 import ops

    g = load_golden(name)
    mode = str(g["mode"])
    n, H = g["nom_X"].shape[:2]
    T = int(g["T"][0])
    cfg, prob = problem(mode, T)
    flat = lambda k: g[k].reshape((n * H,) + g[k].shape[2:])
    B = n * H
    rn, ra = ops.tube_solve_batched(prob.system, prob.nominal, prob.auxiliary, cu(flat("nom_x0")), cu(flat("aux_x0")),
                                    cu(flat("nom_V_init")), cu(flat("aux_V_init")), cu(flat("nom_theta")),
                                    cu(flat("aux_theta")), gains=True, trace=True)
    assert int(rn.status.abs().sum()) == 0
    # decisions agree except on roundoff ties; trajectories within 1e-9 wherever decisions agree
    good = np.ones(B, dtype=bool)
    tied = np.zeros(B, dtype=bool)
    wild = np.zeros(B, dtype=bool)
    loose = set()
    for tag, r in (("nom", rn), ("aux", ra)):
        it_w, it_g = flat(f"{tag}_iters"), r.iters.cpu().numpy()
        ai_w, ai_g = flat(f"{tag}_alpha_idx"), r.alpha_idx.cpu().numpy()
        cands = flat(f"{tag}_cands")
        for i in range(B):
            if it_w[i] != it_g[i] or not np.array_equal(ai_w[i][: it_w[i]], ai_g[i][: it_g[i]]):
                good[i] = False
                # explained only by a roundoff tie between the reference's own candidate costs (Q12: the
                # saturated KAT-0 nominal problem has all 7 candidates equal to 0-3 ulp) or by candidates
                # that left the safe set (costs ~1e16: chaotic for any implementation)
                c = cands[i][: min(it_w[i], max(it_g[i], 1))]
                srt = np.sort(c, axis=1)
                tied[i] |= bool(np.min((srt[:, 1] - srt[:, 0]) / np.abs(srt[:, 0])) < 1e-9)
                wild[i] |= bool(np.nanmax(c) > 1e12)
    assert (good | tied | wild).all(), f"unexplained decision mismatch at {np.nonzero(~(good | tied | wild))[0]}"
    # measured share of the reference's decisions reproduced exactly (printed below): all of random_paper (56),
    # random_full (24), full_gamma (16), closedloop_full_rand (12); 23/24 of paper_T20 and 22/24 of closedloop_paper_rand
    # (solves among candidates outside the safe set); the KAT-0 closed loops are roundoff-tie chains (0/4, 2/3)
    floor_ = {"closedloop_paper": 0.0, "closedloop_full": 0.5, "closedloop_paper_rand": 0.9}.get(name, 0.95)
    assert good.mean() >= floor_, f"only {good.mean():.0%} of the reference decisions reproduced"
    for key, got in (("nom_X", rn.X), ("nom_V", rn.V), ("aux_X", ra.X), ("aux_V", ra.V), ("nom_K", rn.K),
                     ("aux_K", ra.K)):
        e = per_instance_rel(got.cpu().numpy(), flat(key))
        print(f"[golden {name}] {key}: median {np.median(e):.1e} max(decisions equal) {e[good].max(initial=0.0):.1e} "
              f"max(all) {e.max():.1e}; decisions equal on {good.sum()}/{B}")
        # the fixtures contain a few ill-conditioned instances (reference vs oracle differ by up to 4e-10 there); a
        # solve further off than that must be one the oracle itself cannot hold still under a 1e-14 perturbation
        for i in np.nonzero(good & (e >= 5e-9))[0]:
            if int(i) in loose:     # classified by an earlier tensor of the same step (the auxiliary problem tracks
                continue            # the nominal solution, so it inherits the nominal solve's conditioning)
            _, moved = golden_step_conditioning(oracle, g, cfg, i // H, i % H, key[:3])
            print(f"[golden {name}] {key} solve {i}: off by {e[i]:.1e}, the oracle moves {moved:.1e} under 1e-14")
            assert e[i] < 100.0 * moved, (key, i, e[i], moved)
            loose.add(int(i))
        assert len(loose) <= max(1, B // 20), loose
        if good.any():
            assert np.median(e[good]) < 1e-12, (key, np.median(e[good]))
        if key.endswith(("_X", "_V")):
            # tied candidates differ by <= 1e-14 in V (Q12), so even tie-broken runs agree closely
            sel = tied & ~good & ~wild
            assert e[sel].max(initial=0.0) < 1e-6, (key, e[sel].max())

    # sensitivities and gradients at the reference's own optimum
    Xa, Ua, Xn, Un = cu(flat("aux_X")), cu(flat("aux_V")), cu(flat("nom_X")), cu(flat("nom_V"))
    sa = ops.ddp_sensitivity_batched(prob.system, prob.auxiliary, Xa, Ua, cu(flat("aux_theta")), X_bar=Xn, X_ref=Xn,
                                     U_ref=Un, want_grads=True, want_ref_grads=True)
    scale = np.abs(flat("sa_dL")).reshape(B, -1).max(axis=1)
    for got, key in ((sa.delta_X, "sa_dX"), (sa.delta_V, "sa_dV"), (sa.delta_lambda, "sa_dL")):
        assert np.max(per_instance_rel(got.cpu().numpy(), flat(key), floor=1e-9)) < TOL, key
    gth = sa.g_theta.cpu().numpy()
    if mode == "paper":
        gQ, gR, gqb = gth[:, 0:3] + gth[:, 5:8], gth[:, 3:5], gth[:, 8]
        got = np.concatenate([gQ, gR, gqb[:, None]], axis=1)
        want = np.concatenate([flat("gQ"), flat("gR"), flat("gqb")[:, None]], axis=1)
        e = np.abs(got - want).max(axis=1) / np.maximum(np.abs(want).max(axis=1), 1e-6 * scale)
        assert e.max() < TOL, e.max()
        assert rel_err(sa.loss.cpu().numpy(), flat("loss")) < 1e-12
    else:
        th = flat("aux_theta")
        chain = np.concatenate([1 - np.exp(-th[:, :9]), 1 - np.exp(-(th[:, 9:10] - 1e-6)), 1 - th[:, 10:11] ** 2], axis=1)
        got, want = gth[:, :11] * chain, flat("g_theta_raw")
        e = np.abs(got - want).max(axis=1) / np.maximum(np.abs(want).max(axis=1), 1e-6 * scale)
        assert e.max() < TOL, e.max()
        assert np.max(per_instance_rel(sa.g_xref.cpu().numpy(), flat("g_xref"), floor=1e-9)) < TOL
        assert np.max(per_instance_rel(sa.g_uref.cpu().numpy(), flat("g_uref"), floor=1e-9)) < TOL
        sn = ops.ddp_sensitivity_batched(prob.system, prob.nominal, Xn, Un, cu(flat("nom_theta")),
                                         upper_grad_x=cu(flat("sn_gx")), upper_grad_u=cu(flat("sn_gu")), want_grads=True)
        gs = np.maximum(np.abs(flat("sn_gx")).reshape(B, -1).max(axis=1), 1e-3)
        for got_t, key in ((sn.delta_X, "sn_dX"), (sn.delta_V, "sn_dV"), (sn.delta_lambda, "sn_dL")):
            e = np.abs(got_t.cpu().numpy() - flat(key)).reshape(B, -1).max(axis=1) / np.maximum(
                np.abs(flat(key)).reshape(B, -1).max(axis=1), 1e-9 * gs)
            assert e.max() < TOL, (key, e.max())
        thn = flat("nom_theta")
        chain = np.concatenate([1 - np.exp(-thn[:, :9]), 1 - np.exp(-(thn[:, 9:10] - 1e-6)), 1 - thn[:, 10:11] ** 2,
                                1 - np.exp(-thn[:, 11:12])], axis=1)
        got, want = sn.g_theta.cpu().numpy() * chain, flat("g_theta_bar_raw")
        sc = np.abs(flat("sn_dL")).reshape(B, -1).max(axis=1)
        e = np.abs(got - want).max(axis=1) / np.maximum(np.abs(want).max(axis=1), 1e-6 * sc)
        assert e.max() < TOL, e.max()


# ---------------------------------------------------------------------------
# size-independent properties at the BASELINE batch sizes
# ---------------------------------------------------------------------------
def test_large_batch_properties():
    from differentiable_tube_mpc_b200 import ops

    cfg, prob = problem("paper")
    B, T = 65536, prob.nominal.ilqr.horizon
    x, xb = make_instances(B, 2024, cfg)
    th_n, th_a = cu(prob.theta_nom0), cu(prob.theta_aux0)
    z4, z2 = torch.zeros(B, 4, dtype=torch.float64, device=DEV), torch.zeros(B, 2, dtype=torch.float64, device=DEV)
    xbh = torch.cat([cu(xb), z4[:, :1]], dim=1)
    xh = torch.cat([cu(x), z4[:, :1]], dim=1)
    xbh[:, 3] = ops.primitives_batched(prob.system, prob.nominal, xbh, z2, th_n)["b0"]
    xh[:, 3] = ops.primitives_batched(prob.system, prob.auxiliary, xh, z2, th_a)["b0"]
    V0 = torch.zeros(B, T, 2, dtype=torch.float64, device=DEV)
    rn, ra = ops.tube_solve_batched(prob.system, prob.nominal, prob.auxiliary, xbh, xh, V0, V0, th_n, th_a, trace=True)
    assert int((rn.status != 0).sum()) == 0
    # (1) the returned trajectory IS the rollout of the returned controls, bit for bit
    assert torch.equal(ops.rollout_batched(prob.system, prob.nominal, xbh, rn.V, th_n), rn.X)
    assert torch.equal(ops.rollout_batched(prob.system, prob.auxiliary, xh, ra.V, th_a), ra.X)
    # (2) controls inside the box
    lo, hi = prob.system.ctrl.bounds()
    for r in (rn, ra):
        assert bool((r.V >= cu(lo)).all()) and bool((r.V <= cu(hi)).all())
    # (3) alpha = 0 is in the list, so accepted costs never increase (core/ddp.py:293 + SURVEY A.7)
    for r in (rn, ra):
        c = r.costs
        d = c[:, 1:] - c[:, :-1]
        assert bool((d[~torch.isnan(d)] <= 0).all())
        assert bool((r.iters >= 2).all()) and bool((r.iters <= r.costs.shape[1]).all())
    # (4) shard invariance: solving the batch as 8 slices gives bitwise the same result
    parts = []
    for s in range(8):
        sl = slice(s * B // 8, (s + 1) * B // 8)
        pn, pa = ops.tube_solve_batched(prob.system, prob.nominal, prob.auxiliary, xbh[sl], xh[sl], V0[sl], V0[sl], th_n,
                                        th_a)
        parts.append((pn.X.clone(), pa.X.clone(), pa.V.clone(), pa.iters.clone()))
    assert torch.equal(torch.cat([p[0] for p in parts]), rn.X)
    assert torch.equal(torch.cat([p[1] for p in parts]), ra.X)
    assert torch.equal(torch.cat([p[2] for p in parts]), ra.V)
    assert torch.equal(torch.cat([p[3] for p in parts]), ra.iters)
    # (5) idempotence: warm-started from its own optimum the solver stops after the minimum two iterations
    r2 = ops.ilqr_solve_batched(prob.system, prob.nominal, xbh, rn.V, th_n, trace=True)
    conv = rn.iters < prob.nominal.ilqr.max_iter
    assert float((r2.iters[conv] == 2).float().mean()) > 0.97
    # (6) deterministic batch sum: same bits on every call, equal to the float64 sum within roundoff
    sa = ops.ddp_sensitivity_batched(prob.system, prob.auxiliary, ra.X, ra.V, th_a, X_bar=rn.X, X_ref=rn.X, U_ref=rn.V)
    s1, s2 = ops.reduce_columns(sa.g_theta), ops.reduce_columns(sa.g_theta)
    assert torch.equal(s1, s2)
    ref = sa.g_theta.sum(dim=0)
    assert float(((s1 - ref).abs() / (sa.g_theta.abs().sum(dim=0) + 1e-300)).max()) < 1e-12
    assert bool(torch.isfinite(sa.g_theta).all()) and bool(torch.isfinite(sa.loss).all())


def test_edge_cases():
    from differentiable_tube_mpc_b200 import ops

    cfg, prob = problem("paper", T=5)
    th = cu(prob.theta_nom0)
    # empty batch
    r = ops.ilqr_solve_batched(prob.system, prob.nominal, torch.zeros(0, 4, dtype=torch.float64, device=DEV),
                               torch.zeros(0, 5, 2, dtype=torch.float64, device=DEV), th)
    assert r.X.shape == (0, 6, 4)
    # ragged sizes around the block size, T = 1
    cfg1, prob1 = problem("paper", T=1)
    for B in (1, 31, 129, 1000):
        x0 = torch.zeros(B, 4, dtype=torch.float64, device=DEV)
        x0[:, 2] = 0.7853981633974483
        x0[:, 3] = 0.05272775814973351
        r = ops.ilqr_solve_batched(prob1.system, prob1.nominal, x0, torch.zeros(B, 1, 2, dtype=torch.float64, device=DEV),
                                   th)
        assert int(r.status.sum()) == 0 and bool(torch.isfinite(r.X).all())
        assert bool((r.X == r.X[:1]).all())
    # non-finite input is reported per instance instead of raising (reference: FloatingPointError)
    x0 = torch.zeros(4, 4, dtype=torch.float64, device=DEV)
    x0[:, 3] = 0.05
    x0[2, 0] = float("nan")
    r = ops.ilqr_solve_batched(prob.system, prob.nominal, x0, torch.zeros(4, 5, 2, dtype=torch.float64, device=DEV), th)
    assert r.status.cpu().tolist() == [0, 0, 1, 0]
    # wrong device / dtype / shape are rejected before any launch
    with pytest.raises(ValueError):
        ops.ilqr_solve_batched(prob.system, prob.nominal, torch.zeros(2, 4, dtype=torch.float64),
                               torch.zeros(2, 5, 2, dtype=torch.float64), th)
    with pytest.raises(TypeError):
        ops.ilqr_solve_batched(prob.system, prob.nominal, torch.zeros(2, 4, dtype=torch.float16, device=DEV),
                               torch.zeros(2, 5, 2, dtype=torch.float16, device=DEV), th)
    with pytest.raises(ValueError):
        ops.ilqr_solve_batched(prob.system, prob.nominal, torch.zeros(2, 4, dtype=torch.float64, device=DEV),
                               torch.zeros(2, 7, 2, dtype=torch.float64, device=DEV), th)


def test_run_nominal_angle_wrapped_variant_against_reference():
    """BASELINE config #2 problem (run_nominal.py semantics: wrapped heading error, reg = ilqr_reg) vs the reference."""
    from dataclasses import replace
    from differentiable_tube_mpc_b200 import ops

    g = load_golden("nominal_wrap")
    cfg, prob = problem("paper", T=30)
    ocp = replace(prob.nominal, angle_wrap=True, ilqr=replace(prob.nominal.ilqr, reg=float(cfg["system"]["ilqr_reg"])))
    n = g["X"].shape[0]
    flat = lambda k: g[k].reshape((2 * n,) + g[k].shape[2:])
    r = ops.ilqr_solve_batched(prob.system, ocp, cu(flat("x0")), cu(flat("V_init")), cu(prob.theta_nom0), gains=True,
                               trace=True)
    assert int(r.status.sum()) == 0
    it_w, it_g = flat("iters"), r.iters.cpu().numpy()
    ai_w, ai_g = flat("alpha_idx"), r.alpha_idx.cpu().numpy()
    good = np.array([it_w[i] == it_g[i] and np.array_equal(ai_w[i][: it_w[i]], ai_g[i][: it_g[i]]) for i in range(2 * n)])
    assert good.mean() > 0.7
    tame = np.nanmax(flat("best"), axis=1) < 1e12   # a start inside an obstacle (cost ~1e21) has chaotic gains
    for key, got in (("X", r.X), ("V", r.V), ("K", r.K)):
        e = per_instance_rel(got.cpu().numpy(), flat(key))
        sel = good & tame if key == "K" else good
        assert e[sel].max() < 1e-9, (key, e[sel].max())
