"""GPU tests of the reference-facing API: the function-level shims (ilqr_solve / ddp_sensitivity / ift_gradient with the
reference's signatures), the closed-loop drop-in `run_closed_loop_experiment`, the autograd.Function, and fp32 mode --
all against what the reference computed (tests/golden)."""
import math
import os

import numpy as np
import pytest
import torch

from helpers import golden_step_conditioning, load_golden, make_instances, per_instance_rel, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=DEV)


def _family(cfg, mode, theta_n_raw=None):
    """The declarative closures of one reference closed-loop step (what core/tube_mpc.py:268-399 builds)."""
    from differentiable_tube_mpc_b200 import (BoxClampControl, CircleObstacle, DBaSConfig, DubinsConfig, ILQRConfig)

    s = cfg["system"]
    dub = DubinsConfig(dt=s["dt"], v_max=s["control_bounds"]["v_max"], omega_max=s["control_bounds"]["omega_max"],
                       x_target=tuple(s["target"]))
    obs = [CircleObstacle(center=tuple(o["center"]), radius=o["radius"]) for o in cfg["environment"]["obstacles"]]
    ctrl = BoxClampControl(u_min=torch.tensor([s["control_bounds"]["v_min"], -dub.omega_max], dtype=torch.float64),
                           u_max=torch.tensor([dub.v_max, dub.omega_max], dtype=torch.float64))
    return dub, obs, ctrl


def test_reference_signature_shims_single_and_batched():
    from differentiable_tube_mpc_b200 import DBaSConfig, ILQRConfig, default_dubins_config
    from differentiable_tube_mpc_b200.compat import (DubinsDBaSDynamics, IFTInputs, TrackingCost, ddp_sensitivity,
                                                     ift_gradient, ilqr_solve, rollout)

    g = load_golden("random_paper")
    cfg = default_dubins_config()
    dub, obs, ctrl = _family(cfg, "paper")
    eps = cfg["dbas"]["eps"]
    dyn = DubinsDBaSDynamics(dub, obs, 20.0, "smoothmin", DBaSConfig(alpha=0.0, gamma=0.0, eps=eps))
    icfg = ILQRConfig(horizon=50, max_iter=10, tol=1e-3, line_search_alphas=tuple(cfg["system"]["line_search_alphas"]))
    i = 3
    th = g["nom_theta"][i, 0]
    cost = TrackingCost(cu(th[0:3]), cu(th[3:5]), cu(th[8]), cu(th[5:8]), target=cfg["system"]["target"])
    # unbatched call, CPU tensors in -> CPU tensors out, exactly like the reference call site (tube_mpc.py:844-857)
    x0, V0 = torch.tensor(g["nom_x0"][i, 0]), torch.tensor(g["nom_V_init"][i, 0])
    X, V = ilqr_solve(x0=x0, V_init=V0, cfg=icfg, f=dyn, f_jac=dyn.jacobian, ctrl=ctrl, stage_cost=cost.stage_cost,
                      terminal_cost=cost.terminal_cost, stage_derivs=cost.stage_derivs,
                      terminal_derivs=cost.terminal_derivs, debug_name="iLQR-nominal")
    assert X.shape == (51, 4) and V.shape == (50, 2) and X.device.type == "cpu"
    assert rel_err(X.numpy(), g["nom_X"][i, 0]) < 1e-9 and rel_err(V.numpy(), g["nom_V"][i, 0]) < 1e-9
    assert torch.equal(rollout(x0, V, f=dyn), X)
    # the callable objects evaluate single points like the reference's closures do
    xh, u = torch.tensor(g["nom_X"][i, 0][7]), torch.tensor(g["nom_V"][i, 0][7])
    assert rel_err(dyn(xh, u).numpy(), g["nom_X"][i, 0][8]) < 1e-13
    A, Bm = dyn.jacobian(xh, u)
    assert A.shape == (4, 4) and Bm.shape == (4, 2) and float(A[3, 3]) == 0.0
    c = sum(float(cost.stage_cost(torch.tensor(g["nom_X"][i, 0][k]), torch.tensor(g["nom_V"][i, 0][k]), k)) for k in range(50))
    c += float(cost.terminal_cost(torch.tensor(g["nom_X"][i, 0][50])))
    assert abs(c - g["nom_best"][i, 0][int(g["nom_iters"][i, 0]) - 1]) < 1e-9 * abs(c)

    # auxiliary solve + sensitivity + gradient, batched over 8 fixture instances that share the aux weights
    sel = slice(0, 8)
    tha = g["aux_theta"][0, 0]
    Xn, Un = cu(g["nom_X"][sel, 0]), cu(g["nom_V"][sel, 0])
    caux = TrackingCost(cu(tha[0:3]), cu(tha[3:5]), cu(tha[8]), cu(tha[5:8]), X_ref=Xn[..., :3], U_ref=Un)
    acfg = ILQRConfig(horizon=50, max_iter=20, tol=1e-3, line_search_alphas=tuple(cfg["system"]["line_search_alphas"]))
    Xa, Ua = ilqr_solve(x0=cu(g["aux_x0"][sel, 0]), V_init=cu(g["aux_V_init"][sel, 0]), cfg=acfg, f=dyn,
                        f_jac=dyn.jacobian, ctrl=ctrl, stage_cost=caux.stage_cost, terminal_cost=caux.terminal_cost,
                        stage_derivs=caux.stage_derivs, terminal_derivs=caux.terminal_derivs)
    assert np.max(per_instance_rel(Xa.cpu().numpy(), g["aux_X"][sel, 0])) < 5e-9
    Xa, Ua = cu(g["aux_X"][sel, 0]), cu(g["aux_V"][sel, 0])
    x_nom = Xn[..., :3]
    sens = ddp_sensitivity(X=Xa, V=Ua, f=dyn, f_jac=dyn.jacobian, ctrl=ctrl, stage_hess=caux.stage_hess,
                           terminal_hess=caux.terminal_hess,
                           upper_grad_x=lambda xk, k: torch.cat([2.0 * (xk[..., :3] - x_nom[..., k, :]), 2.0 * xk[..., 3:]], -1),
                           upper_grad_u=lambda vk, k: torch.zeros_like(vk),
                           upper_grad_xN=lambda xN: torch.cat([2.0 * (xN[..., :3] - x_nom[..., 50, :]), 2.0 * xN[..., 3:]], -1))
    for got, key in ((sens.delta_X, "sa_dX"), (sens.delta_V, "sa_dV"), (sens.delta_lambda, "sa_dL")):
        assert np.max(per_instance_rel(got.cpu().numpy(), g[key][sel, 0], floor=1e-9)) < 1e-9, key
    with pytest.raises(TypeError, match="closures"):
        ilqr_solve(x0=x0, V_init=V0, cfg=icfg, f=lambda x, u: x, stage_cost=cost.stage_cost,
                   terminal_cost=cost.terminal_cost, stage_derivs=cost.stage_derivs, terminal_derivs=cost.terminal_derivs)
    bad = x0.clone()
    bad[0] = float("nan")
    with pytest.raises(FloatingPointError, match="iLQR-nominal"):
        ilqr_solve(x0=bad, V_init=V0, cfg=icfg, f=dyn, ctrl=ctrl, stage_cost=cost.stage_cost, debug_name="iLQR-nominal",
                   terminal_cost=cost.terminal_cost, stage_derivs=cost.stage_derivs, terminal_derivs=cost.terminal_derivs)


def test_ift_gradient_shim_matches_reference_autograd():
    """ift_gradient(theta_tensors=[raw tensors..., X_ref, U_ref]) against the gradients the reference's autograd
    produced in full mode (core/tube_mpc.py:489-502, 586-599)."""
    from differentiable_tube_mpc_b200 import DBaSConfig, TubeProblem, default_dubins_config
    from differentiable_tube_mpc_b200.compat import DubinsDBaSDynamics, IFTInputs, TrackingCost, ift_gradient

    g = load_golden("closedloop_full")
    cfg = default_dubins_config()
    dub, obs, ctrl = _family(cfg, "full")
    prob = TubeProblem.from_config(cfg, mode="full")
    th_n, th_a = prob.make_thetas(DEV, torch.float64)   # t = 0 parameters of the run
    for t_ in th_a.tensors() + th_n.tensors():
        t_.requires_grad_(True)
    eps = cfg["dbas"]["eps"]
    t = 0
    Xn, Un = cu(g["nom_X"][0, t]), cu(g["nom_V"][0, t])
    X_ref = Xn[:, :3].clone().requires_grad_(True)
    U_ref = Un.clone().requires_grad_(True)
    dyn_a = DubinsDBaSDynamics(dub, obs, 20.0, "smoothmin", DBaSConfig(alpha=0.5, gamma=0.0, eps=eps), theta=th_a)
    cost_a = TrackingCost(X_ref=X_ref, U_ref=U_ref, theta=th_a)
    grads = ift_gradient(inputs=IFTInputs(X=cu(g["aux_X"][0, t]), V=cu(g["aux_V"][0, t]), delta_X=cu(g["sa_dX"][0, t]),
                                          delta_V=cu(g["sa_dV"][0, t]), delta_lambda=cu(g["sa_dL"][0, t])),
                         theta_tensors=th_a.tensors() + [X_ref, U_ref], xi_fn=lambda: None, f_fn=dyn_a,
                         stage_cost_fn=cost_a.stage_cost, terminal_cost_fn=cost_a.terminal_cost)
    assert len(grads) == 8 and [tuple(x.shape) for x in grads[:6]] == [(3,), (2,), (3,), (), (), ()]
    got = torch.cat([x.reshape(-1) for x in grads[:6]]).cpu().numpy()
    assert rel_err(got, g["g_theta_raw"][0, t]) < 1e-10
    assert rel_err(grads[6].cpu().numpy(), g["g_xref"][0, t], floor=1e-12) < 1e-10
    assert rel_err(grads[7].cpu().numpy(), g["g_uref"][0, t], floor=1e-12) < 1e-10
    dyn_n = DubinsDBaSDynamics(dub, obs, 20.0, "smoothmin", DBaSConfig(alpha=0.5, gamma=0.0, eps=eps), theta=th_n)
    cost_n = TrackingCost(target=cfg["system"]["target"], theta=th_n)
    gn = ift_gradient(inputs=IFTInputs(X=Xn, V=Un, delta_X=cu(g["sn_dX"][0, t]), delta_V=cu(g["sn_dV"][0, t]),
                                       delta_lambda=cu(g["sn_dL"][0, t])), theta_tensors=th_n.tensors(),
                      xi_fn=lambda: None, f_fn=dyn_n, stage_cost_fn=cost_n.stage_cost, terminal_cost_fn=cost_n.terminal_cost)
    got = torch.cat([x.reshape(-1) for x in gn]).cpu().numpy()
    scale = np.abs(g["sn_dL"][0, t]).max()
    assert rel_err(got, g["g_theta_bar_raw"][0, t], floor=1e-12 * scale) < 1e-9
    th_a.Q_raw.requires_grad_(False)
    assert ift_gradient(inputs=IFTInputs(X=cu(g["aux_X"][0, t]), V=cu(g["aux_V"][0, t]), delta_X=cu(g["sa_dX"][0, t]),
                                         delta_V=cu(g["sa_dV"][0, t]), delta_lambda=cu(g["sa_dL"][0, t])),
                        theta_tensors=th_a.tensors(), f_fn=dyn_a, stage_cost_fn=cost_a.stage_cost,
                        terminal_cost_fn=cost_a.terminal_cost)[0] is None


@pytest.mark.parametrize("name,adapt_nominal", [("closedloop_paper", False), ("closedloop_full", True)])
def test_closed_loop_drop_in_matches_the_reference_run(tmp_path, name, adapt_nominal):
    """run_closed_loop_experiment with the reference's own disturbance samples injected: states, controls, barrier
    state, loss, adapted weights and the summary must follow the reference run step by step."""
    from differentiable_tube_mpc_b200 import default_dubins_config
    from differentiable_tube_mpc_b200.tube_mpc import run_closed_loop_experiment

    g = load_golden(name)
    H = g["nom_x0"].shape[1]
    cfg = default_dubins_config()
    cfg["system"]["task_horizon_H"] = H
    cfg["adaptation"]["adapt_nominal"] = adapt_nominal
    w = cu(g["w"][0])
    out = run_closed_loop_experiment(cfg, device=torch.device(DEV), run_dir=str(tmp_path),
                                     disturbance_fn=lambda t: w[t].view(1, 3))
    ld = lambda f: np.load(os.path.join(str(tmp_path), f))
    # KAT-0 is a chain of roundoff ties (Q12): trajectories agree closely but not to 1e-9 after tie-broken steps
    tol = 1e-6
    assert rel_err(ld("x_real.npy"), g["aux_x0"][0, :, :3]) < tol
    assert rel_err(ld("x_bar.npy"), g["nom_x0"][0, :, :3]) < tol
    assert rel_err(ld("b_real.npy"), g["aux_x0"][0, :, 3]) < tol
    assert rel_err(ld("u_real.npy"), g["aux_V"][0, :, 0, :]) < 1e-5
    assert rel_err(ld("u_bar.npy"), g["nom_V"][0, :, 0, :]) < 1e-5
    if name == "closedloop_paper":
        assert rel_err(ld("loss.npy"), g["loss"][0]) < 1e-6
        assert rel_err(ld("Qa_history.npy")[:-1], g["aux_theta"][0, 1:, 0:3]) < 1e-8
        assert rel_err(ld("Ra_history.npy")[:-1], g["aux_theta"][0, 1:, 3:5]) < 1e-8
        assert rel_err(ld("qba_history.npy")[:-1], g["aux_theta"][0, 1:, 8]) < 1e-8
        assert out["summary"]["note"].startswith("Dubins run aligned to paper")
    else:
        assert rel_err(ld("Qa_history.npy")[:-1], g["aux_theta"][0, 1:, 0:3]) < 1e-8
    assert abs(out["summary"]["final_loss"] - float(g["summary_final_loss"][0])) < 1e-6 * abs(float(g["summary_final_loss"][0]))
    assert abs(out["summary"]["final_barrier_state"] - float(g["summary_final_b"][0])) < 1e-6
    assert set(out["summary"]) >= {"system", "H", "N", "final_state", "final_barrier_state", "final_loss"}
    assert ld("x_real.npy").shape == (H, 3) and ld("u_real.npy").shape == (H, 2) and ld("loss.npy").shape == (H,)


@pytest.mark.parametrize("name", ["closedloop_paper_rand", "closedloop_full_rand"])
def test_closed_loop_from_random_starts_follows_the_reference(oracle, name):
    """Closed loops the reference ran from random, perturbed starts (no roundoff-tie chain, unlike KAT-0): with the
    reference's own disturbance samples injected, states, controls, barrier state, loss and the adapted weights follow
    the reference step by step.  One tube per run, as in the reference (its update uses that tube's gradient)."""
    from differentiable_tube_mpc_b200 import default_dubins_config
    from differentiable_tube_mpc_b200.closed_loop import ClosedLoop

    g = load_golden(name)
    n, H = g["nom_x0"].shape[:2]
    worst, n_full = {}, 0
    for i in range(n):
        # leading steps of this run that are well-conditioned (judged by the CPU oracle on the reference's recorded
        # inputs): after an ill-conditioned solve the run is a different, equally valid, closed loop
        Hc, amp = 0, 0.0
        for t in range(H):
            c = [golden_step_conditioning(oracle, g, default_dubins_config(), i, t, tag) for tag in ("nom", "aux")]
            if not all(same and moved < 1e-8 for same, moved in c):
                break
            Hc += 1
            amp = max(amp, max(moved for _, moved in c))
        tol = max(1e-9, 10.0 * amp)      # 1e-9 unless the oracle itself moves more than 1e-10 under a 1e-14 perturbation
        n_full += Hc == H
        cfg = default_dubins_config()
        cfg["system"]["task_horizon_H"] = H
        cfg["adaptation"]["adapt_nominal"] = name == "closedloop_full_rand"
        w = cu(g["w"][i])
        loop = ClosedLoop(cfg, device=torch.device(DEV), batch=1, disturbance_fn=lambda t: w[t].view(1, 3),
                          x0=cu(g["x0"][i]).view(1, 3), x_bar0=cu(g["xbar0"][i]).view(1, 3))
        logs = {k: v.detach().cpu().numpy() for k, v in loop.run(H).items()}
        # the reference's decisions at every step must have been reproduced for a step-by-step 1e-9 comparison to be fair
        pairs = [("x_real", g["aux_x0"][i, :, :3]), ("x_bar", g["nom_x0"][i, :, :3]), ("b_real", g["aux_x0"][i, :, 3]),
                 ("u_real", g["aux_V"][i, :, 0, :]), ("u_bar", g["nom_V"][i, :, 0, :])]
        if name == "closedloop_paper_rand":
            pairs.append(("loss", g["loss"][i]))
        pairs += [("Qa", g["aux_theta"][i, 1:, 0:3]), ("Ra", g["aux_theta"][i, 1:, 3:5]), ("qba", g["aux_theta"][i, 1:, 8])]
        for k, want in pairs:
            got = logs[k][:, 0] if k in ("x_real", "x_bar", "b_real", "u_real", "u_bar", "loss") else logs[k][:-1]
            got = np.asarray(got, dtype=np.float64).reshape(want.shape)
            m = min(H, Hc + 1) if k in ("x_real", "x_bar", "b_real") else Hc   # the state entering step t needs steps < t
            e = rel_err(got[:m], want[:m])
            assert e < tol, (i, k, e, tol)
            worst[k] = max(worst.get(k, 0.0), e)
        print(f"[closed loop {name}] run {i}: {Hc} of {H} steps well-conditioned, tolerance {tol:.0e}")
    print(f"[closed loop {name}] worst relative error over the well-conditioned steps of {n} runs: " +
          ", ".join(f"{k} {v:.1e}" for k, v in worst.items()))
    assert n_full >= n - 1, "the fixture is supposed to be non-degenerate"
    assert max(worst.values()) < 1e-8


def test_closed_loop_batch_extension_and_errors(tmp_path):
    from differentiable_tube_mpc_b200 import default_dubins_config
    from differentiable_tube_mpc_b200.tube_mpc import run_closed_loop_experiment

    cfg = default_dubins_config()
    cfg["system"]["task_horizon_H"] = 2
    cfg["system"]["horizon_N"] = 10
    cfg["batch"] = 64
    torch.manual_seed(0)
    out = run_closed_loop_experiment(cfg, device=torch.device(DEV), run_dir=str(tmp_path))
    assert math.isfinite(out["summary"]["final_loss"])
    cfg["system"]["name"] = "cartpole"
    with pytest.raises(NotImplementedError):
        run_closed_loop_experiment(cfg, device=torch.device(DEV), run_dir=str(tmp_path))
    cfg["system"]["name"] = "dubins"
    with pytest.raises(ValueError, match="CUDA"):
        run_closed_loop_experiment(cfg, device=torch.device("cpu"), run_dir=str(tmp_path))


def test_autograd_function_backward_is_the_sensitivity_sweep():
    from differentiable_tube_mpc_b200 import TubeProblem, default_dubins_config, ops
    from differentiable_tube_mpc_b200.autograd import tube_mpc_solve

    g = load_golden("random_full")
    cfg = default_dubins_config()
    prob = TubeProblem.from_config(cfg, mode="full")
    n = 12
    th_n = cu(g["nom_theta"][:n, 0]).requires_grad_(True)
    th_a = cu(g["aux_theta"][:n, 0]).requires_grad_(True)
    Xa, Ua, Xn, Un = tube_mpc_solve(prob, cu(g["nom_x0"][:n, 0]), cu(g["aux_x0"][:n, 0]), cu(g["nom_V_init"][:n, 0]),
                                    cu(g["aux_V_init"][:n, 0]), th_n, th_a)
    # the reference's upper loss with the nominal trajectory treated as data (its closed loop, App. B Q10)
    loss = ((Xa[..., :3] - Xn[..., :3].detach()) ** 2).sum() + (Xa[..., 3] ** 2).sum()
    loss.backward()
    tha = g["aux_theta"][:n, 0]
    chain = np.concatenate([1 - np.exp(-tha[:, :9]), 1 - np.exp(-(tha[:, 9:10] - 1e-6)), 1 - tha[:, 10:11] ** 2], axis=1)
    got = th_a.grad.cpu().numpy()[:, :11] * chain
    want = g["g_theta_raw"][:n, 0]
    e = np.abs(got - want).max(axis=1) / np.abs(want).max(axis=1)
    assert np.median(e) < 1e-9 and e.max() < 1e-6, e
    thn = g["nom_theta"][:n, 0]
    chain = np.concatenate([1 - np.exp(-thn[:, :9]), 1 - np.exp(-(thn[:, 9:10] - 1e-6)), 1 - thn[:, 10:11] ** 2,
                            1 - np.exp(-thn[:, 11:12])], axis=1)
    got = th_n.grad.cpu().numpy() * chain
    want = g["g_theta_bar_raw"][:n, 0]
    e = np.abs(got - want).max(axis=1) / np.maximum(np.abs(want).max(axis=1), 1e-9)
    assert np.median(e) < 1e-8 and e.max() < 1e-5, e
    # shared parameters: the gradient is the deterministic batch sum
    th_s = cu(g["aux_theta"][0, 0]).requires_grad_(True)
    Xa2, _, Xn2, _ = tube_mpc_solve(prob, cu(g["nom_x0"][:n, 0]), cu(g["aux_x0"][:n, 0]), cu(g["nom_V_init"][:n, 0]),
                                    cu(g["aux_V_init"][:n, 0]), cu(g["nom_theta"][0, 0]), th_s)
    ((Xa2[..., 3] ** 2).sum()).backward()
    assert th_s.grad.shape == (12,) and bool(torch.isfinite(th_s.grad).all())


def test_fp32_mode_tracks_fp64():
    """Optional fp32 mode (north_star: 1e-4): same kernels instantiated for float.  Compared with the fp64 path on
    instances whose decisions agree (App. B Q16); gains and trajectories within 1e-4 relative in the median."""
    from differentiable_tube_mpc_b200 import TubeProblem, default_dubins_config, ops

    cfg = default_dubins_config()
    prob = TubeProblem.from_config(cfg, mode="paper")
    B, T = 512, 50
    x, xb = make_instances(B, 77, cfg, h_min=1.0)
    th_n, th_a = np.asarray(prob.theta_nom0), np.asarray(prob.theta_aux0)
    res = {}
    for dt in (torch.float64, torch.float32):
        z2 = torch.zeros(B, 2, dtype=dt, device=DEV)
        xbh = torch.cat([cu(xb, dt), torch.zeros(B, 1, dtype=dt, device=DEV)], 1)
        xbh[:, 3] = ops.primitives_batched(prob.system, prob.nominal, xbh, z2, cu(th_n, dt))["b0"]
        V0 = torch.zeros(B, T, 2, dtype=dt, device=DEV)
        res[dt] = ops.ilqr_solve_batched(prob.system, prob.nominal, xbh, V0, cu(th_n, dt), trace=True)
        assert res[dt].X.dtype == dt and int(res[dt].status.sum()) == 0
    r64, r32 = res[torch.float64], res[torch.float32]
    same = (r64.iters == r32.iters) & (r64.alpha_idx == r32.alpha_idx).all(dim=1)
    assert float(same.float().mean()) > 0.5
    e = per_instance_rel(r32.X.double().cpu().numpy(), r64.X.cpu().numpy())[same.cpu().numpy()]
    assert np.median(e) < 1e-4, np.median(e)
    assert np.quantile(e, 0.9) < 5e-3, np.quantile(e, 0.9)


@pytest.mark.parametrize("name", ["random_paper_f32", "random_full_f32"])
def test_fp32_mode_against_the_reference_float32_run(name):
    """fp32 mode vs the reference executed with use_float64: false (core/tube_mpc.py:51, 675): 1e-4 relative on
    instances whose discrete decisions agree (north_star; SURVEY App. B Q16 -- fp32 decisions are roundoff-sensitive,
    the reference's own fp32 run differs from its fp64 run by up to 8e-5 in X/U and 1e-3 in the sensitivities)."""
    from differentiable_tube_mpc_b200 import TubeProblem, default_dubins_config, ops

    g = load_golden(name)
    mode = str(g["mode"])
    cfg = default_dubins_config()
    prob = TubeProblem.from_config(cfg, mode=mode)
    f32 = torch.float32
    n = g["nom_X"].shape[0]
    c32 = lambda k: cu(g[k][:, 0], f32)
    rn, ra = ops.tube_solve_batched(prob.system, prob.nominal, prob.auxiliary, c32("nom_x0"), c32("aux_x0"),
                                    c32("nom_V_init"), c32("aux_V_init"), c32("nom_theta"), c32("aux_theta"), trace=True)
    assert rn.X.dtype == f32 and int(rn.status.sum()) == 0
    good = np.ones(n, dtype=bool)
    for tag, r in (("nom", rn), ("aux", ra)):
        it_w, it_g = g[f"{tag}_iters"][:, 0], r.iters.cpu().numpy()
        ai_w, ai_g = g[f"{tag}_alpha_idx"][:, 0], r.alpha_idx.cpu().numpy()
        for i in range(n):
            if it_w[i] != it_g[i] or not np.array_equal(ai_w[i][: it_w[i]], ai_g[i][: it_g[i]]):
                good[i] = False
    # full mode uses tol = 1e-6 on costs of ~1e5, far below fp32 resolution: its tolerance break is pure roundoff
    assert good.mean() >= (0.4 if mode == "paper" else 0.2), f"fp32 decisions reproduced on only {good.mean():.0%}"
    for key, got in (("nom_X", rn.X), ("nom_V", rn.V), ("aux_X", ra.X), ("aux_V", ra.V)):
        e = per_instance_rel(got.double().cpu().numpy(), g[key][:, 0])
        print(f"[fp32 {name}] {key}: decisions equal {good.sum()}/{n}, median {np.median(e[good]):.1e} max {e[good].max():.1e}; "
              f"all instances median {np.median(e):.1e}")
        assert np.median(e[good]) < 1e-4, (key, np.median(e[good]))
        assert e[good].max() < 5e-3, (key, e[good].max())
    # ---- the honest pin (round 2) ---------------------------------------------------------------------------------
    # fp32 decisions are roundoff ties almost everywhere near convergence (candidate costs equal to 1-8 ulp_fp32, the
    # tolerance break decided by whether two fp32 costs are bitwise equal), so "same decisions" cannot define the
    # comparable set.  The yardstick is the reference itself: delta_i = how far the reference's OWN fp32 run is from its
    # fp64 run on instance i (goldens *_f32_ref64: the same instances solved by the reference in fp64).  Where the
    # reference's two precisions agree to a quarter of the tolerance the instance is well-posed in fp32, and the CUDA
    # fp32 mode must be within 1e-4 (MAX, not median) of both reference runs.  The other instances are UNPINNABLE: the
    # reference's own two answers already differ by up to 1.6e-2 there, and a third fp32 implementation lands
    # somewhere else again (measured: up to 2e-2 from the fp32 reference); they are reported, and only their median is
    # bounded.
    g64 = load_golden(name + "_ref64")
    assert np.allclose(g64["x0"], g["x0"])
    keys = (("nom_X", rn.X), ("nom_V", rn.V), ("aux_X", ra.X), ("aux_V", ra.V))
    delta = np.max([per_instance_rel(g[k][:, 0], g64[k][:, 0], floor=1e-3) for k, _ in keys], axis=0)
    pin = delta <= 2.5e-5
    hist = np.histogram(np.log10(np.maximum(delta, 1e-9)), bins=[-9, -7, -6, -5, -4.6, -4, -3, -2, 0])[0]
    print(f"[fp32 {name}] reference fp32-vs-fp64 gap per instance: log10 histogram (-9,-7,-6,-5,-4.6,-4,-3,-2,0) = {hist.tolist()}; "
          f"pinned {pin.sum()}/{n}")
    assert pin.sum() >= (0.3 if mode == "paper" else 0.6) * n   # measured: 11/32 (paper), see the histogram above
    for key, got in keys:
        e32 = per_instance_rel(got.double().cpu().numpy(), g[key][:, 0], floor=1e-3)
        e64 = per_instance_rel(got.double().cpu().numpy(), g64[key][:, 0], floor=1e-3)
        print(f"[fp32 {name}] {key}: pinned ({pin.sum()}) max vs ref32 {e32[pin].max():.1e}, vs ref64 {e64[pin].max():.1e}; "
              f"unpinnable ({(~pin).sum()}) median {np.median(e32[~pin]):.1e} max {e32[~pin].max():.1e} "
              f"(reference fp32-vs-fp64 there: median {np.median(delta[~pin]):.1e} max {delta[~pin].max():.1e})")
        # a pinned instance may exceed 1e-4 only where a roundoff-tied decision (accepted step / tolerance break) went the
        # other way than in the reference's fp32 run -- then the solve stops an iteration earlier or later (measured: one
        # instance of 11 in paper mode, 3e-4 in X_aux) -- never with identical decisions, and on at most 10 % of them
        off = pin & ((e32 >= 1e-4) | (e64 >= 1e-4))
        assert not np.any(off & good), (key, np.nonzero(off & good)[0], e32[off & good])
        assert off.sum() <= max(1, 0.1 * pin.sum()), (key, off.sum(), pin.sum())
        assert e32[pin & ~off].max() < 1e-4 and e32[pin].max() < 1e-3
        assert np.median(e32[~pin]) < 1e-3, (key, np.median(e32[~pin]))
    # sensitivity + gradients at the reference's own fp32 optimum
    Xa, Ua, Xn, Un = c32("aux_X"), c32("aux_V"), c32("nom_X"), c32("nom_V")
    sa = ops.ddp_sensitivity_batched(prob.system, prob.auxiliary, Xa, Ua, c32("aux_theta"), X_bar=Xn, X_ref=Xn, U_ref=Un,
                                     want_grads=True)
    for got, key in ((sa.delta_X, "sa_dX"), (sa.delta_V, "sa_dV"), (sa.delta_lambda, "sa_dL")):
        e = per_instance_rel(got.double().cpu().numpy(), g[key][:, 0], floor=1e-3)
        print(f"[fp32 {name}] {key}: median {np.median(e):.1e} max {e.max():.1e}")
        assert np.median(e) < 1e-3, (key, np.median(e))
    assert rel_err(sa.loss.double().cpu().numpy(), g["loss"][:, 0] if "loss" in g.files else
                   ((g["aux_X"][:, 0, :, :3] - g["nom_X"][:, 0, :, :3]) ** 2).sum((1, 2)) + (g["aux_X"][:, 0, :, 3] ** 2).sum(1)) < 1e-5


def test_device_closed_loop_batch():
    """closed_loop.ClosedLoop: B tubes for H steps on the device (plant step, DBaS propagation, warm-start shift,
    Algorithm-2 update) with device-side logs.  With the learning rate at 0 the tubes are independent, so a tube's
    trajectory must not depend on the batch it runs in (bitwise)."""
    from differentiable_tube_mpc_b200 import default_dubins_config
    from differentiable_tube_mpc_b200.closed_loop import ClosedLoop

    cfg = default_dubins_config()
    cfg["system"]["horizon_N"] = 20
    cfg["adaptation"]["lr_eta"] = 0.0
    H = 4
    w = torch.rand(H, 256, 3, dtype=torch.float64, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3)) * 0.1 - 0.05
    x0 = cu(make_instances(256, 9, cfg)[0])
    runs = {}
    for B in (256, 64):
        loop = ClosedLoop(cfg, device=DEV, batch=B, n_log=64, x0=x0[:B], disturbance_fn=lambda t, B=B: w[t, :B].contiguous())
        runs[B] = loop.run(H)
        assert runs[B]["x_real"].shape == (H, 64, 3) and runs[B]["loss"].shape == (H, 64) and runs[B]["Qa"].shape == (H, 3)
    for k in ("x_real", "u_real", "x_bar", "b_real", "loss"):
        assert torch.equal(runs[256][k], runs[64][k]), k
    # default device disturbance + adaptation on: runs, stays finite, parameters move, reproducible from the seed
    cfg["adaptation"]["lr_eta"] = 0.05
    a = ClosedLoop(cfg, device=DEV, batch=128, n_log=2).run(3)
    b = ClosedLoop(cfg, device=DEV, batch=128, n_log=2).run(3)
    assert torch.equal(a["x_real"], b["x_real"]) and bool(torch.isfinite(a["loss"]).all())
    assert not torch.equal(a["Qa"][0], a["Qa"][-1])
    # a non-finite start state is reported when the run is checked, not silently carried
    bad = x0[:8].clone()
    bad[3, 0] = float("nan")
    with pytest.raises(FloatingPointError):
        ClosedLoop(cfg, device=DEV, batch=8, x0=bad).run(2)


def test_default_workspace_is_per_stream():
    """calls that are not stream-ordered against each other must not share scratch."""
    from differentiable_tube_mpc_b200 import ops

    a = ops.default_workspace(DEV)
    assert ops.default_workspace(DEV) is a
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        b = ops.default_workspace(DEV)
        assert ops.default_workspace(DEV) is b
    assert b is not a and ops.default_workspace(DEV) is a
