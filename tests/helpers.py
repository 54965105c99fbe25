"""Shared test utilities: seeded synthetic instances, golden loading, conditioning-aware comparison."""
from __future__ import annotations

import math
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, f"{name}.npz"))


from differentiable_tube_mpc_b200.workload import h_smoothmin_np, make_instances  # noqa: E402,F401


def rel_err(a, b, floor=0.0):
    """max|a-b| / max(max|b|, floor): tensor infinity-norm relative error."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), floor, 1e-300))


def per_instance_rel(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64).reshape(a.shape[0], -1)
    b = np.asarray(b, dtype=np.float64).reshape(b.shape[0], -1)
    num = np.max(np.abs(a - b), axis=1)
    den = np.maximum(np.max(np.abs(b), axis=1), max(floor, 1e-300))
    return num / den


def golden_step_conditioning(oracle, g, cfg, i, t, tag):
    """Is step (i, t) of a golden set well-conditioned?  Judged by the CPU oracle alone, on the reference's recorded
    inputs of that solve: (the oracle reproduces the reference's decisions, how far the oracle's own answer moves when
    the initial state is perturbed by 1e-14 relative).  A solve whose iterates pass through the steep part of the
    barrier amplifies roundoff by ten orders of magnitude in any implementation; it cannot be compared at 1e-9."""
    mode = str(g["mode"])
    env = oracle.env_from_cfg(cfg)
    nom, aux = oracle.ocps_from_cfg(cfg, mode)
    ocp = nom if tag == "nom" else aux
    ocp.set_theta(g[f"{tag}_theta"][i, t])
    Xref = g["nom_X"][i, t][:, :3] if tag == "aux" else None
    Uref = g["nom_V"][i, t] if tag == "aux" else None
    x0, V0 = g[f"{tag}_x0"][i, t], g[f"{tag}_V_init"][i, t]
    a = oracle.ilqr_solve(env, ocp, x0, V0, Xref, Uref)
    b = oracle.ilqr_solve(env, ocp, x0 * (1 + 1e-14), V0, Xref, Uref)
    it = int(g[f"{tag}_iters"][i, t])
    same = a["iters"] == it and np.array_equal(a["alpha_idx"][:it], g[f"{tag}_alpha_idx"][i, t][:it])
    moved = max(rel_err(b[k], a[k]) for k in ("X", "V", "K"))
    return bool(same), float(moved)
