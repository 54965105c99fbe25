"""Algorithm-2 parameter updates (momentum + projection), on device.

paper mode: reference core/tube_mpc.py:978-984   v <- mu v + g ; Qa <- max(Qa - eta v, 0) ; Ra <- max(Ra - eta v, 1e-4) ;
                                                 qba <- clip(qba - eta v, 0, 1)
full mode:  reference core/tube_mpc.py:187-249   optional norm clip, vel <- mu vel + g, p <- p - eta vel, projection of
                                                 the RAW tensors by name (Q, Qf >= 0; R in [1e-4, 1e4]; qb in [0,1];
                                                 gamma in [-1,1]; alpha in [0,1]; tight in [0,2])
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
from torch import Tensor


@dataclass
class PaperAdapter:
    """Adapts the auxiliary weights (Q[3], R[2], q_b) as the reference's paper mode does."""
    Qa: Tensor
    Ra: Tensor
    qba: Tensor
    eta: float
    mom: float
    vQ: Tensor = None
    vR: Tensor = None
    vqb: Tensor = None

    def __post_init__(self):
        self.vQ = torch.zeros_like(self.Qa)
        self.vR = torch.zeros_like(self.Ra)
        self.vqb = torch.zeros_like(self.qba)

    def theta(self) -> Tensor:
        """[12] constrained vector in the C-ABI layout: the aux terminal weight is Qa (App. B Q6), alpha=gamma=tight=0."""
        z = torch.zeros(3, dtype=self.Qa.dtype, device=self.Qa.device)
        return torch.cat([self.Qa, self.Ra, self.Qa, self.qba.view(1), z])

    @staticmethod
    def grads_from_g12(g12: Tensor):
        """(gQ, gR, gqb) of core/tube_mpc.py:974-976 from d/d(Q, R, Qf, qb, ...): gQ includes the k = T term."""
        return g12[..., 0:3] + g12[..., 5:8], g12[..., 3:5], g12[..., 8]

    def update(self, g12_sum: Tensor) -> None:
        gQ, gR, gqb = self.grads_from_g12(g12_sum)
        self.vQ = self.mom * self.vQ + gQ
        self.vR = self.mom * self.vR + gR
        self.vqb = self.mom * self.vqb + gqb
        self.Qa = (self.Qa - self.eta * self.vQ).clamp(min=0.0)
        self.Ra = (self.Ra - self.eta * self.vR).clamp(min=1e-4)
        self.qba = (self.qba - self.eta * self.vqb).clamp(min=0.0, max=1.0)


_PROJ = {"Q_raw": (0.0, None), "Qf_raw": (0.0, None), "R_raw": (1e-4, 1e4), "qb_raw": (0.0, 1.0),
         "gamma_raw": (-1.0, 1.0), "alpha_raw": (0.0, 1.0), "tight_raw": (0.0, 2.0)}

NAMES_THETA = ["Q_raw", "R_raw", "Qf_raw", "qb_raw", "alpha_raw", "gamma_raw"]
NAMES_THETA_BAR = NAMES_THETA + ["tight_raw"]


def project_(p: Tensor, name: str) -> None:
    lo, hi = _PROJ[name]
    p.clamp_(min=lo, max=hi)


def split_raw_grads(g_raw: Tensor, with_tight: bool) -> List[Tensor]:
    """[12] raw-gradient vector -> list shaped like theta.tensors() (core/params.py:36-37, 58-59)."""
    out = [g_raw[0:3], g_raw[3:5], g_raw[5:8], g_raw[8], g_raw[9], g_raw[10]]
    if with_tight:
        out.append(g_raw[11])
    return out


def apply_update(params: Sequence[Tensor], grads: Sequence[Optional[Tensor]], vel: List[Tensor], names: Sequence[str], *,
                 lr: float, momentum: float = 0.0, clip_norm: float = 0.0, project_params: bool = False) -> None:
    """In-place update of raw parameter tensors, reference core/tube_mpc.py:232-249."""
    with torch.no_grad():
        for i, (p, g) in enumerate(zip(params, grads)):
            if g is None:
                continue
            g_eff = g
            if clip_norm and clip_norm > 0:
                n = torch.linalg.vector_norm(g_eff).item()
                if n > clip_norm:
                    g_eff = g_eff * (clip_norm / (n + 1e-12))
            if momentum and momentum > 0:
                vel[i].mul_(momentum).add_(g_eff)
                step = vel[i]
            else:
                step = g_eff
            p -= lr * step
            if project_params:
                project_(p, names[i])
