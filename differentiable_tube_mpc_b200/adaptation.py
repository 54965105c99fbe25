"""Algorithm-2 parameter updates (momentum + projection) on packed device vectors.

Both adapters keep their whole state in three 12-vectors in the C-ABI theta layout
([0:3] Q, [3:5] R, [5:8] Qf, [8] q_b, [9] alpha, [10] gamma, [11] tight): the parameters, the momentum buffer and the
constrained values the kernels read.  `update(g_sum)` takes the batch-summed (and, across ranks, all-reduced) gradient
with respect to the CONSTRAINED parameters, as `sens_kernel` produces it, and works in place with a fixed, small
sequence of elementwise operations: no host synchronisation, no reallocation, so a closed-loop step can be captured
into a CUDA graph and every rank that sees the same summed gradient ends up with bitwise the same parameters.
The same code runs on CPU tensors (world_size-2 gloo tests).

paper mode (`PaperAdapter`), reference core/tube_mpc.py:965-984: only the auxiliary Q [3], R [2], q_b adapt; the terminal
    weight IS Q (App. B Q6), so its gradient is folded into Q's; v <- mu v + g, p <- p - eta v, then Q >= 0,
    R >= 1e-4, 0 <= q_b <= 1.
full mode (`RawAdapter`), reference core/tube_mpc.py:187-249: the parameters are the RAW tensors of core/params.py;
    g_raw = g_constrained * d constrained / d raw (softplus' = sigmoid, tanh' = 1 - tanh^2); optional 2-norm clip per
    parameter group; momentum; p <- p - eta step; optional projection of the raw values by group name.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
from torch import Tensor

NTHETA = 12
# parameter groups of the packed layout, in the order of NominalTheta.tensors() (core/params.py:36-37, 58-59)
GROUP_NAMES = ("Q_raw", "R_raw", "Qf_raw", "qb_raw", "alpha_raw", "gamma_raw", "tight_raw")
GROUP_SLICES = (slice(0, 3), slice(3, 5), slice(5, 8), slice(8, 9), slice(9, 10), slice(10, 11), slice(11, 12))
GROUP_OF = (0, 0, 0, 1, 1, 2, 2, 2, 3, 4, 5, 6)
NAMES_THETA = list(GROUP_NAMES[:6])
NAMES_THETA_BAR = list(GROUP_NAMES)

_INF = math.inf
# projection box per group (reference `_project`, core/tube_mpc.py:187-224)
_PROJ = {"Q_raw": (0.0, _INF), "Qf_raw": (0.0, _INF), "R_raw": (1e-4, 1e4), "qb_raw": (0.0, 1.0),
         "gamma_raw": (-1.0, 1.0), "alpha_raw": (0.0, 1.0), "tight_raw": (0.0, 2.0)}


def _vec(values, like: Tensor) -> Tensor:
    return torch.tensor(list(values), dtype=like.dtype, device=like.device)


_consts = {}


def _layout_consts(like: Tensor, with_tight: bool):
    """Device constants of the packed layout, made once per (device, dtype): selector of the tanh entry (gamma), the
    1e-6 offset of alpha, and the mask that removes `tight` from a parameter set without it.  Everything below is
    tensor-tensor arithmetic: no host scalar is copied to the device inside an update (CUDA-graph capture forbids it)."""
    key = (like.device, like.dtype, bool(with_tight))
    if key not in _consts:
        _consts[key] = (_vec([0.0] * 10 + [1.0, 0.0], like).bool(), _vec([0.0] * 9 + [1e-6, 0.0, 0.0], like),
                        _vec([1.0] * 11 + [1.0 if with_tight else 0.0], like))
    return _consts[key]


def constrain(raw: Tensor, with_tight: bool) -> Tensor:
    """raw [12] -> constrained [12] (core/params.py:29-35, 52-57): softplus, softplus + 1e-6 (alpha), tanh (gamma)."""
    is_tanh, offset, mask = _layout_consts(raw, with_tight)
    return torch.where(is_tanh, torch.tanh(raw), torch.nn.functional.softplus(raw) + offset) * mask


def chain_factor(raw: Tensor, with_tight: bool) -> Tensor:
    """d constrained / d raw, elementwise [12]."""
    is_tanh, _, mask = _layout_consts(raw, with_tight)
    return torch.where(is_tanh, 1.0 - torch.tanh(raw) ** 2, torch.sigmoid(raw)) * mask


class RawAdapter:
    """Full-mode update of one parameter set (theta or theta_bar) held as a packed raw vector."""

    def __init__(self, raw: Tensor, *, with_tight: bool, lr: float, momentum: float = 0.0, clip_norm: float = 0.0,
                 project_params: bool = False):
        assert raw.shape == (NTHETA,)
        self.raw = raw.clone()
        self.with_tight = bool(with_tight)
        if not self.with_tight:
            self.raw[11] = 0.0
        self.vel = torch.zeros_like(self.raw)
        self.lr, self.momentum, self.clip_norm = float(lr), float(momentum), float(clip_norm)
        self.project = bool(project_params)
        self.theta = constrain(self.raw, self.with_tight)             # what the kernels read; refreshed in place
        self._lo = _vec([_PROJ[GROUP_NAMES[g]][0] for g in GROUP_OF], raw)
        self._hi = _vec([_PROJ[GROUP_NAMES[g]][1] for g in GROUP_OF], raw)
        self._group = torch.tensor(GROUP_OF, dtype=torch.long, device=raw.device)
        self._mask = _vec([1.0] * 11 + [1.0 if with_tight else 0.0], raw)

    @classmethod
    def from_tensors(cls, tensors: Sequence[Tensor], **kw) -> "RawAdapter":
        """From the reference's per-group raw tensors (Q_raw [3], R_raw [2], Qf_raw [3], qb, alpha, gamma[, tight])."""
        flat = [t.reshape(-1) for t in tensors]
        with_tight = len(flat) == 7
        if not with_tight:
            flat.append(torch.zeros(1, dtype=flat[0].dtype, device=flat[0].device))
        return cls(torch.cat(flat), with_tight=with_tight, **kw)

    def tensors(self) -> List[Tensor]:
        """Views of the packed raw vector shaped like `NominalTheta.tensors()` / `AuxiliaryTheta.tensors()`."""
        n = 7 if self.with_tight else 6
        return [self.raw[s] if s.stop - s.start > 1 else self.raw[s.start] for s in GROUP_SLICES[:n]]

    def raw_gradient(self, g_constrained: Tensor) -> Tensor:
        return g_constrained * chain_factor(self.raw, self.with_tight)

    def update(self, g_constrained_sum: Tensor) -> None:
        with torch.no_grad():
            g = self.raw_gradient(g_constrained_sum) * self._mask
            if self.clip_norm > 0.0:
                # 2-norm per parameter group, compared on the device: groups above the threshold are scaled onto it
                n = torch.zeros(7, dtype=g.dtype, device=g.device).index_add_(0, self._group, g * g).sqrt_()
                scale = torch.where(n > self.clip_norm, self.clip_norm / (n + 1e-12), torch.ones_like(n))
                g = g * scale[self._group]
            if self.momentum > 0.0:
                self.vel.mul_(self.momentum).add_(g)
                g = self.vel
            self.raw.sub_(self.lr * g)
            if self.project:
                torch.maximum(self.raw, self._lo, out=self.raw)
                torch.minimum(self.raw, self._hi, out=self.raw)
            self.theta.copy_(constrain(self.raw, self.with_tight))


class PaperAdapter:
    """Paper-mode update of the auxiliary weights; `theta` is the [12] vector the kernels read (updated in place)."""

    def __init__(self, Qa: Tensor, Ra: Tensor, qba: Tensor, eta: float, mom: float):
        z = torch.zeros(3, dtype=Qa.dtype, device=Qa.device)
        self.theta = torch.cat([Qa.reshape(3), Ra.reshape(2), Qa.reshape(3), qba.reshape(1), z]).clone()
        self.vel = torch.zeros_like(self.theta)
        self.eta, self.mom = float(eta), float(mom)
        self._lo = _vec([0.0] * 3 + [1e-4] * 2 + [0.0] * 3 + [0.0, 0.0, 0.0, 0.0], Qa)
        self._hi = _vec([_INF] * 8 + [1.0, 0.0, 0.0, 0.0], Qa)
        self._adapts = _vec([1.0] * 5 + [0.0] * 3 + [1.0, 0.0, 0.0, 0.0], Qa)

    @property
    def Qa(self) -> Tensor:
        return self.theta[0:3]

    @property
    def Ra(self) -> Tensor:
        return self.theta[3:5]

    @property
    def qba(self) -> Tensor:
        return self.theta[8]

    @staticmethod
    def grads_from_g12(g12: Tensor):
        """(gQ, gR, gqb) of core/tube_mpc.py:974-976 from d/d(Q, R, Qf, qb, ...): gQ includes the k = T term."""
        return g12[..., 0:3] + g12[..., 5:8], g12[..., 3:5], g12[..., 8]

    def update(self, g12_sum: Tensor) -> None:
        with torch.no_grad():
            g = g12_sum[:NTHETA] * self._adapts
            g[0:3] += g12_sum[5:8]
            self.vel.mul_(self.mom).add_(g)
            p = self.theta - self.eta * self.vel
            torch.maximum(p, self._lo, out=p)
            torch.minimum(p, self._hi, out=p)
            p[5:8] = p[0:3]                       # the terminal weight is Q
            self.theta.copy_(p)


# -- reference-shaped helpers kept for callers that hold per-group tensors ------------------------------------------
def split_raw_grads(g_raw: Tensor, with_tight: bool) -> List[Tensor]:
    """[12] raw-gradient vector -> list shaped like theta.tensors() (core/params.py:36-37, 58-59)."""
    n = 7 if with_tight else 6
    return [g_raw[s] if s.stop - s.start > 1 else g_raw[s.start] for s in GROUP_SLICES[:n]]


def project_(p: Tensor, name: str) -> None:
    lo, hi = _PROJ[name]
    p.clamp_(min=lo, max=None if hi == _INF else hi)


def apply_update(params: Sequence[Tensor], grads: Sequence[Optional[Tensor]], vel: List[Tensor], names: Sequence[str], *,
                 lr: float, momentum: float = 0.0, clip_norm: float = 0.0, project_params: bool = False) -> None:
    """The per-group form of `RawAdapter.update` for a list of raw tensors (same arguments as the reference's
    `_apply_update`): packs the groups, runs the packed update, scatters parameters and momentum back in place."""
    live = [i for i, g in enumerate(grads) if g is not None]
    if not live:
        return
    with torch.no_grad():
        gid = [GROUP_NAMES.index(names[i]) for i in live]
        sizes = [params[i].numel() for i in live]
        p = torch.cat([params[i].reshape(-1) for i in live])
        g = torch.cat([grads[i].reshape(-1).to(p.dtype) for i in live])
        v = torch.cat([vel[i].reshape(-1) for i in live])
        seg = torch.repeat_interleave(torch.arange(len(live), device=p.device), torch.tensor(sizes, device=p.device))
        if clip_norm and clip_norm > 0:
            n = torch.zeros(len(live), dtype=p.dtype, device=p.device).index_add_(0, seg, g * g).sqrt_()
            g = g * torch.where(n > clip_norm, clip_norm / (n + 1e-12), torch.ones_like(n))[seg]
        if momentum and momentum > 0:
            v = v * momentum + g
            g = v
        p = p - lr * g
        if project_params:
            lo = _vec([_PROJ[GROUP_NAMES[k]][0] for k, s in zip(gid, sizes) for _ in range(s)], p)
            hi = _vec([_PROJ[GROUP_NAMES[k]][1] for k, s in zip(gid, sizes) for _ in range(s)], p)
            p = torch.minimum(torch.maximum(p, lo), hi)
        off = 0
        for i, s in zip(live, sizes):
            params[i].copy_(p[off:off + s].view_as(params[i]))
            if momentum and momentum > 0:
                vel[i].copy_(v[off:off + s].view_as(vel[i]))
            off += s
