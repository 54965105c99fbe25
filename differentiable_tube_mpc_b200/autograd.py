"""The two-layer tube-MPC solve as a `torch.autograd.Function`.

forward : nominal iLQR solve, then the auxiliary iLQR solve tracking it (reference core/tube_mpc.py:305-399 / 844-909)
backward: implicit differentiation at the optimum.  The incoming output gradients play the role of the upper-level
          gradients `upper_grad_x / upper_grad_u / upper_grad_xN` of `ddp_sensitivity` (reference core/ddp.py:363,
          384-385); the parameter gradients are the closed-form IFT accumulation (core/ift.py:35-92), and the
          auxiliary problem's gradients w.r.t. its reference (dL/dx_ref, dL/du_ref) are chained into the nominal
          problem exactly as the reference does (core/tube_mpc.py:520-559).

    X_aux, U_aux, X_nom, U_nom = tube_mpc_solve(problem, x0_nom, x0_aux, Vn_init, Va_init, theta_nom, theta_aux)
    loss = ((X_aux[..., :3] - X_nom[..., :3]) ** 2).sum() + (X_aux[..., 3] ** 2).sum()
    loss.backward()        # theta_nom.grad, theta_aux.grad: d loss / d (constrained theta), [B,12] or [12]

Unlike the reference's closed loop (App. B Q10), a direct dependence of the loss on X_nom / U_nom is honoured: its
gradient is added to the nominal problem's upper-level gradient.  Initial states and warm starts get no gradient
(the reference detaches them too: xi_fn returns a detached tensor).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops
from .spec import TubeProblem


class TubeMPCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, problem: TubeProblem, x0_nom: Tensor, x0_aux: Tensor, Vn_init: Tensor, Va_init: Tensor,
                theta_nom: Tensor, theta_aux: Tensor, target: Optional[Tensor]):
        p = problem
        rn, ra = ops.tube_solve_batched(p.system, p.nominal, p.auxiliary, x0_nom.detach(), x0_aux.detach(),
                                        Vn_init.detach(), Va_init.detach(), theta_nom.detach(), theta_aux.detach(),
                                        target=target)
        if bool((rn.status != 0).any()):
            raise FloatingPointError(ops.failure_message("iLQR-nominal", rn) or
                                     ops.failure_message("iLQR-auxiliary", ra, shift=4) or "tube MPC solve: non-finite detected")
        ctx.problem = p
        ctx.target = target
        ctx.shared = (theta_nom.dim() == 1, theta_aux.dim() == 1)
        ctx.save_for_backward(rn.X, rn.V, ra.X, ra.V, theta_nom.detach(), theta_aux.detach())
        ctx.mark_non_differentiable(rn.iters, ra.iters)
        return ra.X, ra.V, rn.X, rn.V, rn.iters, ra.iters

    @staticmethod
    def backward(ctx, gXa, gUa, gXn, gUn, _gi_n, _gi_a):
        p = ctx.problem
        Xn, Un, Xa, Ua, th_n, th_a = ctx.saved_tensors
        B, T = Xn.shape[0], Un.shape[1]
        z = lambda ref: torch.zeros_like(ref)
        gXa = z(Xa) if gXa is None else gXa.contiguous()
        gUa = z(Ua) if gUa is None else gUa.contiguous()
        # auxiliary problem: upper gradients = incoming gradients of (X_aux, U_aux)
        sa = ops.ddp_sensitivity_batched(p.system, p.auxiliary, Xa, Ua, th_a, upper_grad_x=gXa, upper_grad_u=gUa,
                                         X_ref=Xn, U_ref=Un, want_grads=True, want_ref_grads=True)
        # nominal problem: chained reference gradients [dL/dx_ref, 0], dL/du_ref  + any direct gradient
        gx = torch.cat([sa.g_xref, torch.zeros(B, T + 1, 1, dtype=Xn.dtype, device=Xn.device)], dim=2)
        gu = sa.g_uref
        if gXn is not None:
            gx = gx + gXn
        if gUn is not None:
            gu = gu + gUn
        sn = ops.ddp_sensitivity_batched(p.system, p.nominal, Xn, Un, th_n, upper_grad_x=gx.contiguous(),
                                         upper_grad_u=gu.contiguous(), target=ctx.target, want_grads=True)
        g_n = ops.reduce_columns(sn.g_theta) if ctx.shared[0] else sn.g_theta
        g_a = ops.reduce_columns(sa.g_theta) if ctx.shared[1] else sa.g_theta
        return None, None, None, None, None, g_n, g_a, None


def tube_mpc_solve(problem: TubeProblem, x0_nom: Tensor, x0_aux: Tensor, Vn_init: Tensor, Va_init: Tensor,
                   theta_nom: Tensor, theta_aux: Tensor, target: Optional[Tensor] = None
                   ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Differentiable batched tube-MPC solve: returns (X_aux [B,T+1,4], U_aux [B,T,2], X_nom, U_nom)."""
    Xa, Ua, Xn, Un, _, _ = TubeMPCFunction.apply(problem, x0_nom, x0_aux, Vn_init, Va_init, theta_nom, theta_aux, target)
    return Xa, Ua, Xn, Un
