"""ctypes front-end of the CPU parity checker (oracle/dtmpc_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from the product
package (differentiable_tube_mpc_b200 does not import this module and fails
loudly when its CUDA library is missing).

The config -> problem mapping below restates what the reference extracts from
its YAML dict: core/tube_mpc.py:674-768 (paper mode) and :51-165 (full mode).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "libdtmpc_oracle.so")
MAX_OBS = 16
MAX_ALPHAS = 8

AGG = {"smoothmin": 0, "min": 1, "single": 2}


class OrcEnv(C.Structure):
    _fields_ = [
        ("dt", C.c_double),
        ("u_min", C.c_double * 2),
        ("u_max", C.c_double * 2),
        ("active_tol", C.c_double),
        ("n_obs", C.c_int),
        ("obs_agg", C.c_int),
        ("obs_cx", C.c_double * MAX_OBS),
        ("obs_cy", C.c_double * MAX_OBS),
        ("obs_r", C.c_double * MAX_OBS),
        ("obs_beta", C.c_double),
    ]


class OrcOcp(C.Structure):
    _fields_ = [
        ("T", C.c_int),
        ("max_iter", C.c_int),
        ("n_alpha", C.c_int),
        ("angle_wrap", C.c_int),
        ("tol", C.c_double),
        ("reg", C.c_double),
        ("alphas", C.c_double * MAX_ALPHAS),
        ("eps_dyn", C.c_double),
        ("eps_jac", C.c_double),
        ("tight_in_jac", C.c_int),
        ("sens_reg_zero", C.c_int),
        ("Q", C.c_double * 3),
        ("R", C.c_double * 2),
        ("Qf", C.c_double * 3),
        ("qb", C.c_double),
        ("alpha", C.c_double),
        ("gamma", C.c_double),
        ("tight", C.c_double),
        ("target", C.c_double * 3),
    ]

    def set_theta(self, th) -> None:
        th = [float(v) for v in th]
        self.Q[:] = th[0:3]
        self.R[:] = th[3:5]
        self.Qf[:] = th[5:8]
        self.qb, self.alpha, self.gamma, self.tight = th[8], th[9], th[10], th[11]

    def theta(self) -> np.ndarray:
        return np.array(list(self.Q) + list(self.R) + list(self.Qf) + [self.qb, self.alpha, self.gamma, self.tight])

    def copy(self) -> "OrcOcp":
        o = OrcOcp()
        C.memmove(C.byref(o), C.byref(self), C.sizeof(OrcOcp))
        return o


class OrcTubeOut(C.Structure):
    _D = C.POINTER(C.c_double)
    _I = C.POINTER(C.c_int)
    _fields_ = [(n, C.POINTER(C.c_double)) for n in (
        "Xn", "Un", "Xa", "Ua", "Kn", "kn", "Ka", "ka", "dXa", "dVa", "dLa", "dXn", "dVn", "dLn",
        "g_aux", "g_nom", "gXref", "gUref", "loss")] + [(n, C.POINTER(C.c_int)) for n in (
        "iters_n", "iters_a", "status", "aidx_n", "aidx_a")] + [(n, C.POINTER(C.c_double)) for n in (
        "costs_n", "costs_a", "marg_n", "marg_a")]


def build(force: bool = False) -> str:
    """Compile the checker with plain gcc (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(os.path.join(_HERE, f))
                                          for f in ("dtmpc_oracle.c", "dtmpc_oracle.h"))):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_h.restype = C.c_double
        L.orc_h.argtypes = [C.POINTER(OrcEnv), dp]
        L.orc_grad_h.argtypes = [C.POINTER(OrcEnv), dp, dp]
        L.orc_B.restype = C.c_double
        L.orc_B.argtypes = [C.c_double] * 3
        L.orc_dB.restype = C.c_double
        L.orc_dB.argtypes = [C.c_double] * 3
        L.orc_fhat.argtypes = [C.POINTER(OrcEnv), C.POINTER(OrcOcp), dp, dp, dp]
        L.orc_b0.restype = C.c_double
        L.orc_b0.argtypes = [C.POINTER(OrcEnv), C.POINTER(OrcOcp), dp]
        L.orc_aug_jac.argtypes = [C.POINTER(OrcEnv), C.POINTER(OrcOcp), dp, dp, dp, dp]
        L.orc_clamp.argtypes = [C.POINTER(OrcEnv), dp, dp]
        L.orc_active.argtypes = [C.POINTER(OrcEnv), dp, ip]
        L.orc_stage_cost.restype = C.c_double
        L.orc_stage_cost.argtypes = [C.POINTER(OrcOcp), dp, dp, dp, dp]
        L.orc_terminal_cost.restype = C.c_double
        L.orc_terminal_cost.argtypes = [C.POINTER(OrcOcp), dp, dp]
        L.orc_ilqr_solve.restype = C.c_int
        L.orc_ilqr_solve.argtypes = [C.POINTER(OrcEnv), C.POINTER(OrcOcp), dp, dp, dp, dp, dp, dp, dp, dp, ip, ip,
                                     dp, dp]
        L.orc_sensitivity.restype = C.c_int
        L.orc_sensitivity.argtypes = [C.POINTER(OrcEnv), C.POINTER(OrcOcp), dp, dp, dp, dp, dp, dp, dp, ip]
        L.orc_ift_accumulate.argtypes = [C.POINTER(OrcEnv), C.POINTER(OrcOcp), dp, dp, dp, dp, dp, dp, dp, dp, dp, dp]
        L.orc_tube_solve_batch.restype = C.c_int
        L.orc_tube_solve_batch.argtypes = [C.POINTER(OrcEnv), C.POINTER(OrcOcp), C.POINTER(OrcOcp), C.c_long,
                                           dp, dp, dp, dp, dp, C.c_long, dp, C.c_long, C.c_int, C.c_int,
                                           C.POINTER(OrcTubeOut)]
        _lib = L
    return _lib


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


# ---------------------------------------------------------------------------
# config -> problem description
# ---------------------------------------------------------------------------
def softplus(x):
    x = np.asarray(x, dtype=np.float64)
    # torch.nn.functional.softplus (beta=1, threshold=20): core/params.py:9-11
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


def sigmoid(x):
    x = np.asarray(x, dtype=np.float64)
    return 1.0 / (1.0 + np.exp(-x))


def env_from_cfg(cfg: Dict[str, Any]) -> OrcEnv:
    """core/tube_mpc.py:680-721 (identical in :56-88)."""
    s = cfg["system"]
    e = OrcEnv()
    e.dt = float(s["dt"])
    v_max = float(s["control_bounds"]["v_max"])
    om = float(s["control_bounds"]["omega_max"])
    v_min = float(s["control_bounds"].get("v_min", -v_max))
    e.u_min[:] = [v_min, -om]
    e.u_max[:] = [v_max, om]
    e.active_tol = 1e-8
    env = cfg.get("environment", {})
    e.obs_beta = float(env.get("obstacle_smoothmin_beta", 20.0))
    agg = str(env.get("obstacle_aggregation", "min"))
    if "obstacles" in env:
        obs = env["obstacles"]
        e.n_obs = len(obs)
        for i, o in enumerate(obs):
            e.obs_cx[i], e.obs_cy[i] = float(o["center"][0]), float(o["center"][1])
            e.obs_r[i] = float(o["radius"])
        e.obs_agg = AGG["smoothmin"] if agg == "smoothmin" else AGG["min"]
    else:
        o = env.get("obstacle", {"center": [5.0, 5.0], "radius": 1.5})
        e.n_obs = 1
        e.obs_cx[0], e.obs_cy[0], e.obs_r[0] = float(o["center"][0]), float(o["center"][1]), float(o["radius"])
        e.obs_agg = AGG["single"]
    return e


def ocps_from_cfg(cfg: Dict[str, Any], mode: str = "paper"):
    """Returns (nominal OrcOcp, auxiliary OrcOcp) with the constrained parameter
    values the reference would solve with at t=0.

    paper: core/tube_mpc.py:706-768 (alpha=gamma=0, YAML eps, reg default 1e-6, tol 1e-3, YAML alphas,
           aux terminal weight = Qa, :885/:891).
    full:  core/tube_mpc.py:105-165 (YAML values are RAW: softplus/tanh maps, reg=ilqr_reg, default tol and
           alphas, solve-time eps = DBaSConfig default 1e-6, jac-time eps = YAML; App. B Q2-Q5).
    """
    s = cfg["system"]
    T = int(s["horizon_N"])
    nom, aux = OrcOcp(), OrcOcp()
    for o in (nom, aux):
        o.T = T
        o.angle_wrap = 0
        o.tight_in_jac = 0
        o.target[:] = [float(v) for v in s["target"]]
    eps_yaml = float(cfg["dbas"].get("eps", 1e-6))
    if mode == "paper":
        alphas = [float(a) for a in s.get("line_search_alphas", [1.0])]
        for o, key in ((nom, "nominal_max_iter"), (aux, "aux_max_iter")):
            o.max_iter = int(s.get(key, 10))
            o.tol, o.reg = 1e-3, 1e-6
            o.n_alpha = len(alphas)
            o.alphas[: len(alphas)] = alphas
            o.eps_dyn = o.eps_jac = eps_yaml
        cn = cfg["cost_nominal"]
        nom.set_theta(list(cn["Q"]) + list(cn["R"]) + list(cn["Qf"]) + [float(cn["q_b"]), 0.0, 0.0, 0.0])
        ca = cfg.get("cost_auxiliary", {})
        Qa = list(ca.get("Q", [1.0, 1.0, 1.0]))
        Ra = list(ca.get("R", [1.0, 1.0]))
        qba = float(ca.get("q_b", 1.0))
        aux.set_theta(Qa + Ra + Qa + [qba, 0.0, 0.0, 0.0])
    elif mode == "full":
        reg = float(s.get("ilqr_reg", 1e-6))
        for o, key in ((nom, "nominal_max_iter"), (aux, "aux_max_iter")):
            o.max_iter = int(s.get(key, 10))
            o.tol, o.reg = 1e-6, reg
            o.n_alpha = 4
            o.alphas[:4] = [1.0, 0.5, 0.25, 0.1]
            o.eps_dyn, o.eps_jac = 1e-6, eps_yaml
        raw_n, raw_a = raw_theta_from_cfg(cfg)
        nom.set_theta(constrain_theta(raw_n))
        aux.set_theta(constrain_theta(raw_a))
    else:
        raise ValueError(mode)
    return nom, aux


def raw_theta_from_cfg(cfg):
    """core/tube_mpc.py:105-132: raw parameter vectors (12 entries, aux tight unused = -inf -> 0)."""
    cn = cfg["cost_nominal"]
    a0, g0 = float(cfg["dbas"]["alpha"]), float(cfg["dbas"]["gamma"])
    t0 = float(cfg["dbas"].get("nominal_tightening", 0.0))
    raw_n = np.array(list(cn["Q"]) + list(cn["R"]) + list(cn["Qf"]) + [float(cn["q_b"]), a0, g0, t0], dtype=np.float64)
    qba = float(cfg.get("cost_auxiliary", {}).get("q_b", float(cn["q_b"])))
    raw_a = np.array(list(cn["Q"]) + list(cn["R"]) + list(cn["Qf"]) + [qba, a0, g0, -np.inf], dtype=np.float64)
    return raw_n, raw_a


def constrain_theta(raw: np.ndarray) -> np.ndarray:
    """core/params.py:29-35, 53-58."""
    raw = np.asarray(raw, dtype=np.float64)
    out = np.empty(raw.shape, dtype=np.float64)
    out[..., 0:9] = softplus(raw[..., 0:9])
    out[..., 9] = softplus(raw[..., 9]) + 1e-6
    out[..., 10] = np.tanh(raw[..., 10])
    with np.errstate(over="ignore"):
        out[..., 11] = np.where(np.isneginf(raw[..., 11]), 0.0, softplus(raw[..., 11]))
    return out


def chain_theta(raw: np.ndarray, g_constrained: np.ndarray) -> np.ndarray:
    """d/d raw from d/d constrained (softplus' = sigmoid, tanh' = 1 - tanh^2)."""
    raw = np.asarray(raw, dtype=np.float64)
    f = np.empty(raw.shape, dtype=np.float64)
    f[..., 0:10] = sigmoid(raw[..., 0:10])
    f[..., 10] = 1.0 - np.tanh(raw[..., 10]) ** 2
    f[..., 11] = np.where(np.isneginf(raw[..., 11]), 0.0, sigmoid(raw[..., 11]))
    return f * g_constrained


# ---------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------
def h(env: OrcEnv, x) -> float:
    return lib().orc_h(C.byref(env), _dp(_f64(x)))


def grad_h(env: OrcEnv, x) -> np.ndarray:
    g = np.zeros(3)
    lib().orc_grad_h(C.byref(env), _dp(_f64(x)), _dp(g))
    return g


def B(z: float, alpha: float, eps: float) -> float:
    return lib().orc_B(float(z), float(alpha), float(eps))


def dB(z: float, alpha: float, eps: float) -> float:
    return lib().orc_dB(float(z), float(alpha), float(eps))


def fhat(env: OrcEnv, ocp: OrcOcp, xh, u) -> np.ndarray:
    out = np.zeros(4)
    lib().orc_fhat(C.byref(env), C.byref(ocp), _dp(_f64(xh)), _dp(_f64(u)), _dp(out))
    return out


def b0(env: OrcEnv, ocp: OrcOcp, x) -> float:
    return lib().orc_b0(C.byref(env), C.byref(ocp), _dp(_f64(x)))


def aug_jac(env: OrcEnv, ocp: OrcOcp, xh, u):
    A, Bm = np.zeros((4, 4)), np.zeros((4, 2))
    lib().orc_aug_jac(C.byref(env), C.byref(ocp), _dp(_f64(xh)), _dp(_f64(u)), _dp(A), _dp(Bm))
    return A, Bm


def clamp(env: OrcEnv, u) -> np.ndarray:
    out = np.zeros(2)
    lib().orc_clamp(C.byref(env), _dp(_f64(u)), _dp(out))
    return out


def active(env: OrcEnv, u) -> np.ndarray:
    out = np.zeros(2, dtype=np.int32)
    lib().orc_active(C.byref(env), _dp(_f64(u)), _ip(out))
    return out.astype(bool)


def stage_cost(ocp: OrcOcp, xh, u, xref, uref) -> float:
    return lib().orc_stage_cost(C.byref(ocp), _dp(_f64(xh)), _dp(_f64(u)), _dp(_f64(xref)), _dp(_f64(uref)))


def terminal_cost(ocp: OrcOcp, xh, xref) -> float:
    return lib().orc_terminal_cost(C.byref(ocp), _dp(_f64(xh)), _dp(_f64(xref)))


def ilqr_solve(env: OrcEnv, ocp: OrcOcp, x0, V_init, Xref=None, Uref=None) -> Dict[str, Any]:
    T, MI = ocp.T, ocp.max_iter
    x0, V_init = _f64(x0), _f64(V_init)
    Xref = None if Xref is None else _f64(Xref)
    Uref = None if Uref is None else _f64(Uref)
    X, V = np.zeros((T + 1, 4)), np.zeros((T, 2))
    K, k = np.zeros((T, 2, 4)), np.zeros((T, 2))
    iters = np.zeros(1, dtype=np.int32)
    aidx = np.zeros(MI, dtype=np.int32)
    costs, marg = np.zeros(MI), np.zeros(MI)
    st = lib().orc_ilqr_solve(C.byref(env), C.byref(ocp), _dp(x0), _dp(V_init), _dp(Xref), _dp(Uref), _dp(X), _dp(V),
                              _dp(K), _dp(k), _ip(iters), _ip(aidx), _dp(costs), _dp(marg))
    return dict(X=X, V=V, K=K, k=k, iters=int(iters[0]), alpha_idx=aidx, costs=costs, margins=marg, status=st)


def sensitivity(env: OrcEnv, ocp: OrcOcp, X, V, gx, gu) -> Dict[str, Any]:
    T = ocp.T
    dX, dV, dL = np.zeros((T + 1, 4)), np.zeros((T, 2)), np.zeros((T + 1, 4))
    act = np.zeros((T, 2), dtype=np.int32)
    st = lib().orc_sensitivity(C.byref(env), C.byref(ocp), _dp(_f64(X)), _dp(_f64(V)), _dp(_f64(gx)), _dp(_f64(gu)),
                               _dp(dX), _dp(dV), _dp(dL), _ip(act))
    return dict(dX=dX, dV=dV, dL=dL, active=act.astype(bool), status=st)


def ift_accumulate(env: OrcEnv, ocp: OrcOcp, X, V, Xref, Uref, dX, dV, dL) -> Dict[str, Any]:
    T = ocp.T
    g12 = np.zeros(12)
    gXr, gUr = np.zeros((T + 1, 3)), np.zeros((T, 2))
    Xref = None if Xref is None else _f64(Xref)
    Uref = None if Uref is None else _f64(Uref)
    lib().orc_ift_accumulate(C.byref(env), C.byref(ocp), _dp(_f64(X)), _dp(_f64(V)), _dp(Xref), _dp(Uref),
                             _dp(_f64(dX)), _dp(_f64(dV)), _dp(_f64(dL)), _dp(g12), _dp(gXr), _dp(gUr))
    return dict(g=g12, gXref=gXr, gUref=gUr)


def _tube_slice(env, nom, aux, x0n, x0a, Vn0, Va0, th_n, th_a, mode):
    Bn = x0n.shape[0]
    T, MIn, MIa = nom.T, nom.max_iter, aux.max_iter
    z = np.zeros
    out = dict(
        Xn=z((Bn, T + 1, 4)), Un=z((Bn, T, 2)), Xa=z((Bn, T + 1, 4)), Ua=z((Bn, T, 2)),
        Kn=z((Bn, T, 2, 4)), kn=z((Bn, T, 2)), Ka=z((Bn, T, 2, 4)), ka=z((Bn, T, 2)),
        dXa=z((Bn, T + 1, 4)), dVa=z((Bn, T, 2)), dLa=z((Bn, T + 1, 4)),
        dXn=z((Bn, T + 1, 4)), dVn=z((Bn, T, 2)), dLn=z((Bn, T + 1, 4)),
        g_aux=z((Bn, 12)), g_nom=z((Bn, 12)), gXref=z((Bn, T + 1, 3)), gUref=z((Bn, T, 2)), loss=z(Bn),
        iters_n=z(Bn, dtype=np.int32), iters_a=z(Bn, dtype=np.int32), status=z(Bn, dtype=np.int32),
        aidx_n=z((Bn, MIn), dtype=np.int32), aidx_a=z((Bn, MIa), dtype=np.int32),
        costs_n=z((Bn, MIn)), costs_a=z((Bn, MIa)), marg_n=z((Bn, MIn)), marg_a=z((Bn, MIa)),
    )
    o = OrcTubeOut()
    for name, _ in OrcTubeOut._fields_:
        a = out[name]
        setattr(o, name, _ip(a) if a.dtype == np.int32 else _dp(a))
    sn = 0 if th_n.ndim == 1 else 12
    sa = 0 if th_a.ndim == 1 else 12
    lib().orc_tube_solve_batch(C.byref(env), C.byref(nom), C.byref(aux), Bn, _dp(x0n), _dp(x0a), _dp(Vn0), _dp(Va0),
                               _dp(th_n), sn, _dp(th_a), sa, 0 if mode == "paper" else 1, 1, C.byref(o))
    return out


def tube_solve_batch(env: OrcEnv, nom: OrcOcp, aux: OrcOcp, x0n, x0a, Vn0, Va0, theta_n=None, theta_a=None,
                     mode: str = "paper", n_threads: int = 1) -> Dict[str, np.ndarray]:
    """Nominal iLQR -> auxiliary iLQR -> aux sensitivity -> gradient [-> nominal sensitivity -> gradient]
    for B instances; theta_* [12] (shared) or [B,12] constrained values (default: those inside nom/aux)."""
    x0n, x0a, Vn0, Va0 = _f64(x0n), _f64(x0a), _f64(Vn0), _f64(Va0)
    Bn = x0n.shape[0]
    th_n = _f64(nom.theta() if theta_n is None else theta_n)
    th_a = _f64(aux.theta() if theta_a is None else theta_a)
    n_threads = max(1, min(int(n_threads), Bn))
    bounds = np.linspace(0, Bn, n_threads + 1).astype(int)

    def run(j):
        lo, hi = bounds[j], bounds[j + 1]
        tn = th_n if th_n.ndim == 1 else np.ascontiguousarray(th_n[lo:hi])
        ta = th_a if th_a.ndim == 1 else np.ascontiguousarray(th_a[lo:hi])
        return _tube_slice(env, nom, aux, np.ascontiguousarray(x0n[lo:hi]), np.ascontiguousarray(x0a[lo:hi]),
                           np.ascontiguousarray(Vn0[lo:hi]), np.ascontiguousarray(Va0[lo:hi]), tn, ta, mode)

    if n_threads == 1:
        parts = [run(0)]
    else:
        with ThreadPoolExecutor(n_threads) as ex:
            parts = list(ex.map(run, range(n_threads)))
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}


def paper_grads(g12: np.ndarray):
    """core/tube_mpc.py:965-976: gQ includes the k=T term because the aux terminal weight is Qa (Q6)."""
    g12 = np.asarray(g12)
    return g12[..., 0:3] + g12[..., 5:8], g12[..., 3:5], g12[..., 8]
