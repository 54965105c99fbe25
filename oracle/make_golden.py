#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING the reference (read-only at /root/reference).

TEST INFRASTRUCTURE ONLY; runs in the build container (the GPU box has no
/root/reference -- the committed .npz fixtures are what travels).

The reference ships no golden vectors, so parity is pinned on what
its own code computes here (torch CPU, fp64).  Nothing from the reference is
copied into the repo: its modules are imported through a symlink named
`diff_tube_mpc_strict_pt` (the package name its imports require), and
`core/tube_mpc.py` / `core/ddp.py` are re-executed IN MEMORY with three
observation hooks spliced in:

  * the hard-coded start state (core/tube_mpc.py:95-97, 772-773) and the zero
    warm starts (:170-171, 778-779) become overridable from the cfg dict
    (keys `_x0`, `_xbar0`, `_Vn0`, `_Va0`),
  * `ilqr_solve` (core/ddp.py:102-307) records per-iteration candidate costs
    and the gains K, k of every backward pass (the reference does not return
    them, :210-211),
  * ilqr_solve / ddp_sensitivity / ift_gradient / sample_disturbance calls
    made by the closed loop are recorded with their results, and the
    paper-mode closed-form gradient (core/tube_mpc.py:974-976) is recorded.

None of the hooks changes a floating-point operation.

usage:  python oracle/make_golden.py [--workers 8] [--only NAME]
"""
from __future__ import annotations

import argparse
import copy
import importlib
import math
import os
import sys
import types
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
PYREF = os.path.join(HERE, "_ref", "pyref")
PKG = "diff_tube_mpc_strict_pt"
GOLDEN = os.path.join(REPO, "tests", "golden")
RUNS = os.path.join(HERE, "_ref", "runs")

sys.dont_write_bytecode = True  # the reference tree is read-only


def _setup_path():
    os.makedirs(PYREF, exist_ok=True)
    link = os.path.join(PYREF, PKG)
    if not os.path.islink(link):
        os.symlink("/root/reference", link)
    if PYREF not in sys.path:
        sys.path.insert(0, PYREF)


def _replace_once(src: str, old: str, new: str, count: int = 1) -> str:
    n = src.count(old)
    assert n == count, f"expected {count} occurrence(s) of {old!r}, found {n}"
    return src.replace(old, new)


_CACHE = {}


def load_instrumented():
    """Returns (tube_mpc_module, REC dict, yaml cfg)."""
    if "mod" in _CACHE:
        return _CACHE["mod"], _CACHE["rec"], copy.deepcopy(_CACHE["cfg"])
    _setup_path()
    import torch
    import yaml

    torch.set_num_threads(1)
    ddp = importlib.import_module(f"{PKG}.core.ddp")
    REC = {"ilqr": [], "sens": [], "ift": [], "w": [], "paper_g": [], "trace": []}

    # --- ilqr_solve with a trace of candidate costs and gains -------------
    src = open(ddp.__file__, encoding="utf-8").read()
    src = _replace_once(src, "        best_cost: Optional[Tensor] = None\n",
                        "        best_cost: Optional[Tensor] = None\n        _cand_costs = []\n")
    src = _replace_once(src, "            if best_cost is None or cand_cost < best_cost:\n",
                        "            _cand_costs.append(float(cand_cost))\n"
                        "            if best_cost is None or cand_cost < best_cost:\n")
    src = _replace_once(src, "        X, V = best_X, best_U\n",
                        "        X, V = best_X, best_U\n"
                        "        _TRACE.append(dict(cands=list(_cand_costs), K=torch.stack(K_seq).clone(),"
                        " k=torch.stack(k_seq).clone(), best=float(best_cost)))\n")
    ddp_i = types.ModuleType(f"{PKG}.core._ddp_instrumented")
    ddp_i.__package__ = f"{PKG}.core"
    ddp_i.__dict__["_TRACE"] = REC["trace"]
    sys.modules[ddp_i.__name__] = ddp_i  # dataclasses look their module up by name
    exec(compile(src, "<ddp instrumented>", "exec"), ddp_i.__dict__)

    # --- tube_mpc with overridable start state / warm starts --------------
    tm = importlib.import_module(f"{PKG}.core.tube_mpc")
    src = open(tm.__file__, encoding="utf-8").read()
    src = _replace_once(
        src, "    x = torch.tensor([0.0, 0.0, float(np.pi / 4)], device=device, dtype=dtype)\n",
        "    x = torch.tensor(cfg.get('_x0', [0.0, 0.0, float(np.pi / 4)]), device=device, dtype=dtype)\n", 2)
    src = _replace_once(
        src, "    x_bar = x.clone()\n",
        "    x_bar = x.clone() if '_xbar0' not in cfg else torch.tensor(cfg['_xbar0'], device=device, dtype=dtype)\n", 2)
    src = _replace_once(
        src, "    V_nom_ws = torch.zeros(N, 2, device=device, dtype=dtype)\n",
        "    V_nom_ws = torch.zeros(N, 2, device=device, dtype=dtype) if '_Vn0' not in cfg else"
        " torch.tensor(cfg['_Vn0'], device=device, dtype=dtype)\n", 2)
    src = _replace_once(
        src, "    V_aux_ws = torch.zeros(N, 2, device=device, dtype=dtype)\n",
        "    V_aux_ws = torch.zeros(N, 2, device=device, dtype=dtype) if '_Va0' not in cfg else"
        " torch.tensor(cfg['_Va0'], device=device, dtype=dtype)\n", 2)
    src = _replace_once(
        src, "        gqb = (2.0 * db[:-1] * ddb[:-1]).sum() + 2.0 * db[-1] * ddb[-1]\n",
        "        gqb = (2.0 * db[:-1] * ddb[:-1]).sum() + 2.0 * db[-1] * ddb[-1]\n"
        "        _REC['paper_g'].append((gQ.clone(), gR.clone(), gqb.clone(), L.clone()))\n")
    tm_i = types.ModuleType(f"{PKG}.core._tube_mpc_instrumented")
    tm_i.__package__ = f"{PKG}.core"
    tm_i.__dict__["_REC"] = REC
    sys.modules[tm_i.__name__] = tm_i
    exec(compile(src, "<tube_mpc instrumented>", "exec"), tm_i.__dict__)

    # --- recorders around the solver entry points the closed loop calls ---
    def _theta_of(kw):
        """Constrained parameter values the solve ran with: weights from the (constant) cost Hessians the
        closures return (core/cost_derivs.py:73-75, 127-129, 141-146), DBaS values from the f_hat closure cells."""
        x0, N = kw["x0"].detach(), kw["cfg"].horizon
        _, _, lxx, luu, _ = kw["stage_derivs"](x0, kw["V_init"][0].detach(), 0)
        _, pxx = kw["terminal_derivs"](x0)
        cells = dict(zip(kw["f"].__code__.co_freevars, [c.cell_contents for c in (kw["f"].__closure__ or ())]))
        if "db_cfg" in cells:  # paper mode: one fixed DBaSConfig
            alpha, gamma, tight = float(cells["db_cfg"].alpha), float(cells["db_cfg"].gamma), 0.0
        else:
            alpha = float(cells.get("alpha_n", cells.get("alpha_a")))
            gamma = float(cells.get("gamma_n", cells.get("gamma_a")))
            tight = float(cells["tight_n"]) if "tight_n" in cells else 0.0
        d = torch.diagonal
        return torch.cat([d(lxx)[:3] / 2.0, d(luu) / 2.0, d(pxx)[:3] / 2.0, d(lxx)[3:] / 2.0,
                          torch.tensor([alpha, gamma, tight], dtype=lxx.dtype)]).detach().clone()

    def rec_ilqr(**kw):
        n0 = len(REC["trace"])
        X, V = ddp_i.ilqr_solve(**kw)
        REC["ilqr"].append(dict(x0=kw["x0"].detach().clone(), V_init=kw["V_init"].detach().clone(),
                                X=X.detach().clone(), V=V.detach().clone(), trace=REC["trace"][n0:],
                                max_iter=kw["cfg"].max_iter, theta=_theta_of(kw),
                                cfgv=torch.tensor([kw["cfg"].tol, kw["cfg"].reg] + list(kw["cfg"].line_search_alphas),
                                                  dtype=torch.float64)))
        return X, V

    def rec_sens(**kw):
        N = kw["V"].shape[0]
        r = ddp_i.ddp_sensitivity(**kw)
        gx = torch.stack([kw["upper_grad_x"](kw["X"][k], k) for k in range(N)] + [kw["upper_grad_xN"](kw["X"][N])])
        gu = torch.stack([kw["upper_grad_u"](kw["V"][k], k) for k in range(N)])
        REC["sens"].append(dict(dX=r.delta_X.clone(), dV=r.delta_V.clone(), dL=r.delta_lambda.clone(),
                                gx=gx.detach().clone(), gu=gu.detach().clone()))
        return r

    orig_ift = tm_i.ift_gradient

    def rec_ift(**kw):
        g = orig_ift(**kw)
        REC["ift"].append([None if t is None else t.detach().clone() for t in g])
        return g

    orig_w = tm_i.sample_disturbance

    def rec_w(x, *, cfg):
        w = orig_w(x, cfg=cfg)
        REC["w"].append(w.detach().clone())
        return w

    tm_i.ilqr_solve = rec_ilqr
    tm_i.ddp_sensitivity = rec_sens
    tm_i.ift_gradient = rec_ift
    tm_i.sample_disturbance = rec_w

    cfg = yaml.safe_load(open(os.path.join("/root/reference", "configs", "dubins.yaml"), encoding="utf-8"))
    cfg["plot"] = False
    cfg["device"] = "cpu"
    cfg["out_dir"] = RUNS
    _CACHE.update(mod=tm_i, rec=REC, cfg=cfg)
    return tm_i, REC, copy.deepcopy(cfg)


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def _pack_ilqr(r, n_alpha):
    mi = r["max_iter"]
    it = len(r["trace"])
    cands = np.full((mi, n_alpha), np.nan)
    aidx = np.full(mi, -1, dtype=np.int32)
    best = np.full(mi, np.nan)
    for i, t in enumerate(r["trace"]):
        cands[i, : len(t["cands"])] = t["cands"]
        # first strict minimum in list order == reference winner (core/ddp.py:293)
        aidx[i] = int(np.argmin(np.asarray(t["cands"])))
        best[i] = t["best"]
    return dict(x0=_np(r["x0"]), V_init=_np(r["V_init"]), X=_np(r["X"]), V=_np(r["V"]), iters=np.int32(it),
                theta=_np(r["theta"]), cfgv=_np(r["cfgv"]), max_iter=np.int32(mi),
                cands=cands, alpha_idx=aidx, best=best, K=_np(r["trace"][-1]["K"]), k=_np(r["trace"][-1]["k"]))


_THETA_SHAPES = (3, 2, 3, 1, 1, 1, 1)  # Q, R, Qf, qb, alpha, gamma[, tight]: core/params.py:36-37


def _cat_grads(gs):
    return np.concatenate([np.zeros(n) if g is None else _np(g).reshape(-1) for g, n in zip(gs, _THETA_SHAPES)])


def run_steps(cfg, H, seed=0):
    """Run H closed-loop steps of the (instrumented) reference; returns a list of per-step dicts."""
    import torch

    tm_i, REC, _ = load_instrumented()
    for v in REC.values():
        del v[:]
    cfg = copy.deepcopy(cfg)
    cfg["system"]["task_horizon_H"] = H
    torch.manual_seed(seed)
    np.random.seed(seed)
    run_dir = os.path.join(RUNS, f"golden_{os.getpid()}")
    paper = bool(cfg.get("paper_dubins_mode", True)) and not bool(cfg["adaptation"].get("adapt_nominal", True))
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        out = tm_i.run_closed_loop_experiment(cfg, device=torch.device("cpu"), run_dir=run_dir)
    n_alpha = len(cfg["system"]["line_search_alphas"]) if paper else 4
    steps = []
    for t in range(H):
        nom, aux = REC["ilqr"][2 * t], REC["ilqr"][2 * t + 1]
        d = {f"nom_{k}": v for k, v in _pack_ilqr(nom, n_alpha).items()}
        d.update({f"aux_{k}": v for k, v in _pack_ilqr(aux, n_alpha).items()})
        d["w"] = _np(REC["w"][t])
        if paper:
            s = REC["sens"][t]
            gQ, gR, gqb, L = REC["paper_g"][t]
            d.update(sa_dX=_np(s["dX"]), sa_dV=_np(s["dV"]), sa_dL=_np(s["dL"]), sa_gx=_np(s["gx"]), sa_gu=_np(s["gu"]),
                     gQ=_np(gQ), gR=_np(gR), gqb=_np(gqb), loss=_np(L))
        else:
            sa, sn = REC["sens"][2 * t], REC["sens"][2 * t + 1]
            ga, gn = REC["ift"][2 * t], REC["ift"][2 * t + 1]
            d.update(sa_dX=_np(sa["dX"]), sa_dV=_np(sa["dV"]), sa_dL=_np(sa["dL"]), sa_gx=_np(sa["gx"]),
                     sa_gu=_np(sa["gu"]),
                     sn_dX=_np(sn["dX"]), sn_dV=_np(sn["dV"]), sn_dL=_np(sn["dL"]), sn_gx=_np(sn["gx"]),
                     sn_gu=_np(sn["gu"]),
                     g_theta_raw=_cat_grads(ga[:6]), g_xref=_np(ga[6]), g_uref=_np(ga[7]),
                     g_theta_bar_raw=_cat_grads(gn[:7]))
        steps.append(d)
    steps[-1]["summary_final_loss"] = np.float64(out["summary"]["final_loss"])
    steps[-1]["summary_final_b"] = np.float64(out["summary"]["final_barrier_state"])
    return steps


# ---------------------------------------------------------------------------
# instance generators (seeded; SURVEY 8d synthetic inputs)
# ---------------------------------------------------------------------------
def _h_smoothmin(p, cfg):
    obs = cfg["environment"]["obstacles"]
    beta = cfg["environment"]["obstacle_smoothmin_beta"]
    hs = np.array([(p[0] - o["center"][0]) ** 2 + (p[1] - o["center"][1]) ** 2 - o["radius"] ** 2 for o in obs])
    z = -beta * hs
    return -(z.max() + math.log(np.exp(z - z.max()).sum())) / beta


def random_instances(n, seed, cfg, vary_weights=False, h_min=0.25):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        xb = np.array([rng.uniform(0, 10), rng.uniform(0, 10), rng.uniform(-math.pi, math.pi)])
        if _h_smoothmin(xb, cfg) <= h_min:
            continue
        w = rng.uniform(-0.05, 0.05, size=3)
        x = xb + w
        if _h_smoothmin(x, cfg) <= h_min:
            continue
        inst = dict(_x0=x.tolist(), _xbar0=xb.tolist())
        if vary_weights:
            inst["wn"] = rng.uniform(0.5, 2.0, size=9).tolist()
            inst["wa"] = rng.uniform(0.5, 2.0, size=6).tolist()
        out.append(inst)
    return out


def _apply_instance(cfg, inst, mode):
    cfg = copy.deepcopy(cfg)
    cfg["_x0"], cfg["_xbar0"] = inst["_x0"], inst["_xbar0"]
    for k in ("_Vn0", "_Va0"):
        if k in inst:
            cfg[k] = inst[k]
    if "wn" in inst:
        wn, wa = inst["wn"], inst["wa"]
        cn, ca = cfg["cost_nominal"], cfg["cost_auxiliary"]
        cn["Q"] = [a * b for a, b in zip(cn["Q"], wn[0:3])]
        cn["R"] = [a * b for a, b in zip(cn["R"], wn[3:5])]
        cn["Qf"] = [a * b for a, b in zip(cn["Qf"], wn[5:8])]
        cn["q_b"] = cn["q_b"] * wn[8]
        ca["Q"] = [a * b for a, b in zip(ca["Q"], wa[0:3])]
        ca["R"] = [a * b for a, b in zip(ca["R"], wa[3:5])]
        ca["q_b"] = min(1.0, ca["q_b"] * wa[5])
    if "T" in inst:
        cfg["system"]["horizon_N"] = int(inst["T"])
    if "dbas" in inst:
        cfg["dbas"].update(inst["dbas"])
    if inst.get("fp32"):
        cfg["use_float64"] = False   # the reference's float32 mode (core/tube_mpc.py:51, 675)
    if mode == "full":
        cfg["adaptation"]["adapt_nominal"] = True
    return cfg


def _job(args):
    name, inst, mode, H, seed = args
    _, _, cfg = load_instrumented()
    cfg = _apply_instance(cfg, inst, mode)
    steps = run_steps(cfg, H, seed)
    meta = dict(x0=np.array(inst["_x0"]), xbar0=np.array(inst["_xbar0"]),
                wn=np.array(inst.get("wn", [1.0] * 9)), wa=np.array(inst.get("wa", [1.0] * 6)),
                T=np.int32(cfg["system"]["horizon_N"]),
                dbas_alpha=np.float64(cfg["dbas"]["alpha"]), dbas_gamma=np.float64(cfg["dbas"]["gamma"]),
                dbas_tight=np.float64(cfg["dbas"].get("nominal_tightening", 0.0)))
    return name, meta, steps


def _stack(results, H):
    """results: list of (meta, steps) -> dict of arrays with leading dims [n_inst, H, ...]."""
    out = {}
    for k in results[0][0]:
        out[k] = np.stack([m[k] for m, _ in results])
    keys = [k for k in results[0][1][0] if not k.startswith("summary_")]
    for k in keys:
        out[k] = np.stack([np.stack([st[t][k] for t in range(H)]) for _, st in results])
    for k in ("summary_final_loss", "summary_final_b"):
        out[k] = np.stack([st[-1][k] for _, st in results])
    return out


def primitives_golden():
    """KATs of the L1 primitives straight from the reference functions."""
    _setup_path()
    import torch

    from diff_tube_mpc_strict_pt.core.barrier import DBaSConfig, dbas_step, relaxed_inverse_barrier_B_alpha
    from diff_tube_mpc_strict_pt.core.control import BoxClampControl
    from diff_tube_mpc_strict_pt.core.systems.dubins import DubinsConfig, dubins_step
    from diff_tube_mpc_strict_pt.core.systems.dubins_aug_jac import _dB_relaxed_inv_dz, dubins_augmented_jacobian
    from diff_tube_mpc_strict_pt.core.systems import dubins_obstacles as ob

    _, _, cfg = load_instrumented()
    obs = [ob.CircleObstacle(center=tuple(o["center"]), radius=float(o["radius"])) for o in
           cfg["environment"]["obstacles"]]
    rng = np.random.default_rng(7)
    f64 = torch.float64
    pts = np.concatenate([rng.uniform(-1, 11, size=(40, 3)),
                          np.array([[4.0, 2.0, 0.1], [4.0, 3.0, 0.2], [3.3, 2.0, -1.0], [5.0, 5.0, 0.0],
                                    [6.0, 7.0 + 1e-9, 2.0], [2.0, 4.0 - 1.0, 3.0], [5.0, 3.0, 0.3]])])
    out = dict(pts=pts)
    for nm, lst in (("M5", obs), ("M1", obs[:1]), ("M0", [])):
        out[f"h_smooth_{nm}"] = np.array([float(ob.h_multi_circle_obstacles(torch.tensor(p, dtype=f64), obstacles=lst, beta=20.0)) for p in pts])
        out[f"h_min_{nm}"] = np.array([float(ob.h_min_circle_obstacles(torch.tensor(p, dtype=f64), obstacles=lst)) for p in pts])
        out[f"gh_smooth_{nm}"] = np.stack([_np(ob.grad_h_multi_circle_obstacles(torch.tensor(p, dtype=f64), obstacles=lst, beta=20.0)) for p in pts])
        out[f"gh_min_{nm}"] = np.stack([_np(ob.grad_h_min_circle_obstacles(torch.tensor(p, dtype=f64), obstacles=lst)) for p in pts])
    out["h_single"] = np.array([float(ob.h_circle_obstacle(torch.tensor(p, dtype=f64), obs=obs[0])) for p in pts])
    out["gh_single"] = np.stack([_np(ob.grad_h_circle_obstacle(torch.tensor(p, dtype=f64), obs=obs[0])) for p in pts])
    # barrier on both sides of a = max(alpha, eps) and exactly at zeta = a
    zs, als, eps_l, Bv, dBv = [], [], [], [], []
    for alpha in (0.0, 1e-4, 0.1, 0.6931481805599453, 1.0):
        for eps in (1e-4, 1e-6):
            a = max(alpha, eps)
            for z in (a, a * (1 + 1e-12), a * (1 - 1e-12), 0.5 * a, -0.3, 0.0, 2.0 * a, 18.9, 1e-7, -5.0):
                zt = torch.tensor(z, dtype=f64)
                zs.append(z); als.append(alpha); eps_l.append(eps)
                Bv.append(float(relaxed_inverse_barrier_B_alpha(zt, alpha=torch.tensor(alpha, dtype=f64), eps=eps)))
                dBv.append(float(_dB_relaxed_inv_dz(zt, alpha=torch.tensor(alpha, dtype=f64), eps=eps)))
    out.update(B_z=np.array(zs), B_alpha=np.array(als), B_eps=np.array(eps_l), B_val=np.array(Bv), dB_val=np.array(dBv))
    # dbas_step + augmented jacobian for gamma in {0, +-0.8}, alpha in {0, 0.5}, both aggregations
    dub = DubinsConfig(dt=0.01, v_max=10.0, omega_max=math.pi)
    f = lambda x, u: dubins_step(x, u, cfg=dub)
    rows = []
    xs = np.concatenate([pts[:12], pts[-7:]])
    us = rng.uniform([-10, -math.pi], [10, math.pi], size=(len(xs), 2))
    bs = rng.uniform(0.0, 3.0, size=len(xs))
    for agg in ("smoothmin", "min"):
        hfun = (lambda x: ob.h_multi_circle_obstacles(x, obstacles=obs, beta=20.0)) if agg == "smoothmin" else (
            lambda x: ob.h_min_circle_obstacles(x, obstacles=obs))
        for alpha in (0.0, 0.5):
            for gamma in (0.0, 0.8, -0.8):
                dbc = DBaSConfig(barrier_type="inverse", alpha=torch.tensor(alpha, dtype=f64),
                                 gamma=torch.tensor(gamma, dtype=f64), eps=1e-4)
                for x, u, b in zip(xs, us, bs):
                    xt, ut, bt = torch.tensor(x, dtype=f64), torch.tensor(u, dtype=f64), torch.tensor(b, dtype=f64)
                    xn, bn = dbas_step(x_k=xt, u_k=ut, b_k=bt, f=f, h=hfun, cfg=dbc)
                    A, Bm = dubins_augmented_jacobian(torch.cat([xt, bt.view(1)]), ut, cfg=dub, obs=obs, db_cfg=dbc,
                                                      obs_beta=20.0, obs_agg=agg)
                    rows.append(np.concatenate([[0.0 if agg == "smoothmin" else 1.0, alpha, gamma], x, [b], u,
                                                _np(xn), [float(bn)], _np(A).ravel(), _np(Bm).ravel()]))
    out["step_jac"] = np.stack(rows)
    ctrl = BoxClampControl(u_min=torch.tensor([-10.0, -math.pi], dtype=f64), u_max=torch.tensor([10.0, math.pi], dtype=f64))
    uu = np.array([[10.0, 0.0], [10.0 - 1e-9, math.pi], [10.0 - 1e-7, -math.pi + 1e-9], [-10.0, 3.2], [11.0, -4.0],
                   [9.99999999, 3.14159265], [0.0, 0.0]])
    out["ctrl_u"] = uu
    out["ctrl_clamp"] = np.stack([_np(ctrl.clamp(torch.tensor(u, dtype=f64))) for u in uu])
    out["ctrl_active"] = np.stack([ctrl.active_mask(torch.tensor(u, dtype=f64)).numpy() for u in uu])
    return out


def nominal_wrap_golden(workers):
    """run_nominal.py `run_nominal_receding` (angle-wrapped heading error, warm start v = v_max, reg = ilqr_reg;
    run_nominal.py:204-391) from seeded start states, 2 receding-horizon steps each.  The start state
    (run_nominal.py:266) is made overridable in memory and ilqr_solve is the instrumented one."""
    _, _, cfg = load_instrumented()
    rng = np.random.default_rng(909)
    insts = []
    for inst in random_instances(14, 4711, cfg, h_min=0.5):
        x0 = list(inst["_xbar0"])
        x0[2] = float(rng.uniform(-2.0 * math.pi, 2.0 * math.pi))   # heading errors beyond +-pi exercise the wrap
        insts.append(x0)
    with ProcessPoolExecutor(workers) as ex:
        res = list(ex.map(_nominal_job, insts))
    out = {k: np.stack([r[k] for r in res]) for k in res[0]}
    out["mode"] = np.array("nominal_wrap")
    return out


def _nominal_job(x0):
    import contextlib
    import io
    import torch

    tm_i, REC, cfg = load_instrumented()
    for v in REC.values():
        del v[:]
    ddp = importlib.import_module(f"{PKG}.core.ddp")
    ddp_i = sys.modules[f"{PKG}.core._ddp_instrumented"]
    rn = importlib.import_module(f"{PKG}.run_nominal")
    src = open(rn.__file__, encoding="utf-8").read()
    src = _replace_once(src, "    x = torch.tensor([0.0, 0.0, float(np.pi / 4)], device=device, dtype=dtype)\n",
                        "    x = torch.tensor(cfg['_x0'], device=device, dtype=dtype)\n")
    mod = types.ModuleType(f"{PKG}._run_nominal_instrumented")
    mod.__package__ = PKG
    sys.modules[mod.__name__] = mod
    exec(compile(src, "<run_nominal instrumented>", "exec"), mod.__dict__)
    calls = []

    def rec_ilqr(**kw):
        n0 = len(REC["trace"])
        X, V = ddp_i.ilqr_solve(**kw)
        calls.append(dict(x0=kw["x0"].detach().clone(), V_init=kw["V_init"].detach().clone(), X=X.detach().clone(),
                          V=V.detach().clone(), trace=REC["trace"][n0:], max_iter=kw["cfg"].max_iter,
                          theta=torch.zeros(12, dtype=torch.float64),
                          cfgv=torch.tensor([kw["cfg"].tol, kw["cfg"].reg] + list(kw["cfg"].line_search_alphas),
                                            dtype=torch.float64)))
        return X, V

    orig = ddp.ilqr_solve
    ddp.ilqr_solve = rec_ilqr
    try:
        cfg = copy.deepcopy(cfg)
        cfg["_x0"] = x0
        cfg["system"]["task_horizon_H"] = 2
        cfg["system"]["horizon_N"] = 30
        with contextlib.redirect_stdout(io.StringIO()):
            mod.run_nominal_receding(cfg, device=torch.device("cpu"), run_dir=os.path.join(RUNS, f"nominal_{os.getpid()}"))
    finally:
        ddp.ilqr_solve = orig
    n_alpha = len(cfg["system"]["line_search_alphas"])
    packed = [_pack_ilqr(c, n_alpha) for c in calls[:2]]
    while len(packed) < 2:  # the run stopped early (success / collision): repeat the last solve to keep shapes
        packed.append(packed[-1])
    return {k: np.stack([p[k] for p in packed]) for k in packed[0]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--only", type=str, default="")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    os.makedirs(RUNS, exist_ok=True)
    _, _, cfg = load_instrumented()

    kat = dict(_x0=[0.0, 0.0, math.pi / 4], _xbar0=[0.0, 0.0, math.pi / 4])
    sets = {
        # name: (instances, mode, H)
        "closedloop_paper": ([kat], "paper", 4),
        "closedloop_full": ([kat], "full", 3),
        # closed loops from random, perturbed starts (KAT-0 above is a chain of roundoff ties; these are not)
        "closedloop_paper_rand": (random_instances(4, 909, cfg, h_min=0.5), "paper", 6),
        "closedloop_full_rand": (random_instances(3, 910, cfg, h_min=0.5), "full", 4),
        "random_paper": (random_instances(40, 2024, cfg) + random_instances(16, 77, cfg, vary_weights=True), "paper", 1),
        "random_full": (random_instances(16, 4242, cfg) + random_instances(8, 99, cfg, vary_weights=True), "full", 1),
        "paper_T20": ([dict(i, T=20) for i in random_instances(12, 31337, cfg, vary_weights=True, h_min=0.05)], "paper", 2),
        "random_paper_f32": ([dict(i, fp32=True) for i in random_instances(32, 2024, cfg, h_min=1.0)], "paper", 1),
        "random_full_f32": ([dict(i, fp32=True) for i in random_instances(16, 4242, cfg, h_min=1.0)], "full", 1),
        # the SAME instances as the two fp32 sets, solved by the reference in fp64: the reference's own fp32-vs-fp64 gap per
        # instance is the yardstick the fp32 parity test measures the CUDA fp32 mode with
        "random_paper_f32_ref64": (random_instances(32, 2024, cfg, h_min=1.0), "paper", 1),
        "random_full_f32_ref64": (random_instances(16, 4242, cfg, h_min=1.0), "full", 1),
        "full_gamma": ([dict(i, T=20, dbas=dict(alpha=-1.0, gamma=g, nominal_tightening=-2.0))
                        for i, g in zip(random_instances(8, 555, cfg, h_min=0.02), (0.5, -0.5, 0.9, -0.9, 0.3, -0.3, 1.5, -1.5))],
                       "full", 2),
    }
    if not args.only or args.only == "primitives":
        np.savez_compressed(os.path.join(GOLDEN, "primitives.npz"), **primitives_golden())
        print("wrote primitives.npz", flush=True)
    if not args.only or args.only == "nominal_wrap":
        np.savez_compressed(os.path.join(GOLDEN, "nominal_wrap.npz"), **nominal_wrap_golden(args.workers))
        print("wrote nominal_wrap.npz", flush=True)
    for name, (insts, mode, H) in sets.items():
        if args.only and args.only != name:
            continue
        jobs = [(name, inst, mode, H, 1000 + j) for j, inst in enumerate(insts)]
        with ProcessPoolExecutor(args.workers) as ex:
            res = list(ex.map(_job, jobs))
        data = _stack([(m, s) for _, m, s in res], H)
        data["mode"] = np.array(mode)
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **data)
        print(f"wrote {name}.npz  ({len(insts)} instances x {H} steps)", flush=True)


if __name__ == "__main__":
    main()
