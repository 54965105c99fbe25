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
