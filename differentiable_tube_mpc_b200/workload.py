"""Seeded synthetic Dubins batches: input generation for tests and bench.py, host side, numpy.

x_bar0 ~ U([0,10]^2 x [-pi,pi]) with h(x_bar0) > h_min, x0 = x_bar0 + w, w ~ U(-w_max, w_max)^3 (the disturbance box of
configs/dubins.yaml:29-31).  Candidates are drawn in fixed chunks, so instance i is a pure function of (seed, i):
any prefix or slice of a larger batch is identical to generating that many instances directly (shard invariance).
"""
from __future__ import annotations

import math

import numpy as np

_CHUNK = 8192


def h_smoothmin_np(p, cfg):
    """Smooth-min safety function, vectorised over the rows of p (instance screening only)."""
    env = cfg["environment"]
    beta = float(env["obstacle_smoothmin_beta"])
    hs = np.stack([(p[:, 0] - o["center"][0]) ** 2 + (p[:, 1] - o["center"][1]) ** 2 - o["radius"] ** 2
                   for o in env["obstacles"]], axis=0)
    z = -beta * hs
    zm = z.max(axis=0)
    return -(zm + np.log(np.exp(z - zm).sum(axis=0))) / beta


def make_instances(B, seed, cfg, h_min=0.25, w_max=0.05):
    """Returns (x0 [B,3], x_bar0 [B,3])."""
    rng = np.random.default_rng(seed)
    xs, xbs, n = [], [], 0
    while n < B:
        c = np.stack([rng.uniform(0, 10, _CHUNK), rng.uniform(0, 10, _CHUNK), rng.uniform(-math.pi, math.pi, _CHUNK)], 1)
        w = rng.uniform(-w_max, w_max, size=(_CHUNK, 3))
        ok = (h_smoothmin_np(c, cfg) > h_min) & (h_smoothmin_np(c + w, cfg) > h_min)
        xbs.append(c[ok])
        xs.append((c + w)[ok])
        n += int(ok.sum())
    return np.ascontiguousarray(np.concatenate(xs)[:B]), np.ascontiguousarray(np.concatenate(xbs)[:B])
