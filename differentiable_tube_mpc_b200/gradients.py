"""Closed-form weight gradients from the sensitivities: mirror of the reference's core/gradients.py:27-87.

The reference defines these accumulation forms but never calls them (its runs use the autograd-based `ift_gradient` or
the paper-mode sums of core/tube_mpc.py:965-976); they are the independent statement of what the fused CUDA
accumulation (`dtmpc_sens.cuh`) must produce for the weight entries, and are kept for that cross-check
(tests/test_gpu_edge.py).  Inputs may carry a leading batch dimension; tensors stay on their device.

Relation to the kernel's `g_theta` (order Q[3], R[2], Qf[3], q_b, alpha, gamma, tight):
    Q_diag = g_theta[0:3],  R_diag = g_theta[3:5],  Qf_diag = g_theta[5:8],
    q_b    = g_theta[8] - 2 b_T db_T      (these forms stop the q_b sum at k = T-1, core/gradients.py:50, 81;
                                           the cost the solvers use charges q_b b^2 at k = T too, core/tube_mpc.py:829-832)
"""
from __future__ import annotations

from dataclasses import dataclass

from torch import Tensor


@dataclass
class NominalGrads:          # core/gradients.py:11-16
    Q_diag: Tensor
    R_diag: Tensor
    Qf_diag: Tensor
    q_b: Tensor


@dataclass
class AuxiliaryGrads:        # core/gradients.py:19-24
    Q_diag: Tensor
    R_diag: Tensor
    Qf_diag: Tensor
    q_b: Tensor


def _forms(dx: Tensor, du: Tensor, b: Tensor, dX: Tensor, dU: Tensor):
    d_x, d_b = dX[..., :-1], dX[..., -1]
    grad_Q = (2.0 * dx[..., :-1, :] * d_x[..., :-1, :]).sum(dim=-2)
    grad_R = (2.0 * du * dU).sum(dim=-2)
    grad_qb = (2.0 * b[..., :-1] * d_b[..., :-1]).sum(dim=-1)
    grad_Qf = 2.0 * dx[..., -1, :] * d_x[..., -1, :]
    return grad_Q, grad_R, grad_Qf, grad_qb


def grads_nominal_from_deltas(*, X_hat: Tensor, U: Tensor, dX: Tensor, dU: Tensor, target_x: Tensor) -> NominalGrads:
    """core/gradients.py:27-59: stage sums over k < N, separate terminal term; X_hat [..., N+1, 4], U [..., N, 2]."""
    x, b = X_hat[..., :-1], X_hat[..., -1]
    gQ, gR, gQf, gqb = _forms(x - target_x, U, b, dX, dU)
    return NominalGrads(Q_diag=gQ, R_diag=gR, Qf_diag=gQf, q_b=gqb)


def grads_auxiliary_from_deltas(*, X_hat: Tensor, U: Tensor, dX: Tensor, dU: Tensor, X_ref: Tensor,
                                U_ref: Tensor) -> AuxiliaryGrads:
    """core/gradients.py:62-87: the same forms for the tracking cost (x - x_ref, u - u_ref)."""
    x, b = X_hat[..., :-1], X_hat[..., -1]
    gQ, gR, gQf, gqb = _forms(x - X_ref[..., :3], U - U_ref, b, dX, dU)
    return AuxiliaryGrads(Q_diag=gQ, R_diag=gR, Qf_diag=gQf, q_b=gqb)
