"""Reference-signature entry points: `ilqr_solve`, `ddp_sensitivity`, `ift_gradient`, `rollout`.

Same keyword signatures, return types and exceptions as the reference (core/ddp.py:89-117, 310-329; core/ift.py:10-43).
The reference passes Python closures; here the "closures" are declarative objects of the one problem family the
reference wires up (Dubins + DBaS barrier state, diagonal quadratic tracking costs):

    dyn  = DubinsDBaSDynamics(dub_cfg, obstacles, obs_beta, obs_agg, db_cfg)        # f = dyn, f_jac = dyn.jacobian
    cost = TrackingCost(Q, R, qb, Qf, target=...) or TrackingCost(..., X_ref=, U_ref=)
    X, V = ilqr_solve(x0=x0, V_init=V0, cfg=ILQRConfig(...), f=dyn, f_jac=dyn.jacobian, ctrl=ctrl,
                      stage_cost=cost.stage_cost, terminal_cost=cost.terminal_cost,
                      stage_derivs=cost.stage_derivs, terminal_derivs=cost.terminal_derivs)

The shims read the data behind those callables and run the batched CUDA kernels; an arbitrary Python closure raises
TypeError (there is no generic-closure or CPU path).  Every tensor may carry a leading batch dimension.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Sequence, Union

import torch
from torch import Tensor

from . import ops
from .spec import (AuxiliaryTheta, BoxClampControl, CircleObstacle, DBaSConfig, DubinsConfig, ILQRConfig, NominalTheta,
                   OcpSpec, SystemSpec)

Theta = Union[NominalTheta, AuxiliaryTheta]


def _f(v) -> float:
    return float(v.detach().item()) if isinstance(v, Tensor) else float(v)


def _cuda_device(*tensors) -> torch.device:
    for t in tensors:
        if isinstance(t, Tensor) and t.device.type == "cuda":
            return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("differentiable_tube_mpc_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


class DubinsDBaSDynamics:
    """x_hat+ = f_hat(x_hat, u): Dubins step (core/systems/dubins.py:24-43) with the DBaS barrier state
    (core/barrier.py:75-108) over circular obstacles (core/systems/dubins_obstacles.py).  `tight` shifts the safety
    function, h_bar = h - tight (core/tube_mpc.py:152-153); `theta` optionally ties alpha / gamma / tight to a raw
    parameter container so that `ift_gradient` can return gradients for them."""

    def __init__(self, dub_cfg: DubinsConfig, obstacles: Sequence[CircleObstacle], obs_beta: float = 20.0,
                 obs_agg: str = "smoothmin", db_cfg: Optional[DBaSConfig] = None, tight: Any = 0.0,
                 theta: Optional[Theta] = None):
        self.dub_cfg = dub_cfg
        self.obstacles = tuple(obstacles)
        self.obs_beta, self.obs_agg = float(obs_beta), obs_agg
        self.db_cfg = db_cfg or DBaSConfig()
        self.tight = tight
        self.theta = theta

    def dbas_values(self):
        if self.theta is not None:
            t = self.theta
            tight = _f(t.tight()) if isinstance(t, NominalTheta) else 0.0
            return _f(t.alpha()), _f(t.gamma()), tight
        return _f(self.db_cfg.alpha), _f(self.db_cfg.gamma), _f(self.tight)

    def system(self, ctrl: Optional[BoxClampControl]) -> SystemSpec:
        if ctrl is None:
            inf = float("inf")
            ctrl = BoxClampControl(u_min=[-inf, -inf], u_max=[inf, inf])
        return SystemSpec(name="dubins", nx=3, nu=2, dt=self.dub_cfg.dt, horizon_N=0, task_horizon_H=0,
                          dubins=self.dub_cfg, ctrl=ctrl, obstacles=self.obstacles, obs_beta=self.obs_beta,
                          obs_agg=self.obs_agg, y_target=tuple(self.dub_cfg.x_target))

    def _point(self, x_hat: Tensor, u: Tensor):
        dev = _cuda_device(x_hat, u)
        xb = x_hat.reshape(-1, 4).to(dev)
        ub = u.reshape(-1, 2).to(dev)
        a, g, s = self.dbas_values()
        theta = torch.tensor([0.0] * 9 + [a, g, s], dtype=xb.dtype, device=dev)
        ocp = OcpSpec(ILQRConfig(horizon=1), eps_dyn=self.db_cfg.eps, eps_jac=self.db_cfg.eps)
        return ops.primitives_batched(self.system(None), ocp, xb, ub, theta)

    def __call__(self, x_hat: Tensor, u: Tensor) -> Tensor:
        return self._point(x_hat, u)["x_next"].reshape(x_hat.shape).to(x_hat.device)

    def jacobian(self, x_hat: Tensor, u: Tensor):
        """(A, B) of core/systems/dubins_aug_jac.py:61-139 (the tightening does not enter, as in the reference)."""
        out = self._point(x_hat, u)
        lead = x_hat.shape[:-1]
        return out["A"].reshape(*lead, 4, 4).to(x_hat.device), out["B"].reshape(*lead, 4, 2).to(x_hat.device)


class TrackingCost:
    """l = sum Q (x - x_ref)^2 + sum R (u - u_ref)^2 + q_b b^2, phi = sum Qf (x_N - x_ref,N)^2 + q_b b_N^2
    (core/tube_mpc.py:281-303, 358-383, 823-842, 875-894; derivatives core/cost_derivs.py:58-76, 110-146).
    The reference is the constant `target` (nominal OCP) or `X_ref` [.., T+1, 3] / `U_ref` [.., T, 2] (auxiliary)."""

    def __init__(self, Q: Any = None, R: Any = None, qb: Any = None, Qf: Any = None, *, target: Any = None,
                 X_ref: Optional[Tensor] = None, U_ref: Optional[Tensor] = None, angle_wrap: bool = False,
                 theta: Optional[Theta] = None):
        if (target is None) == (X_ref is None):
            raise ValueError("give either target or X_ref / U_ref")
        self.theta = theta
        self._Q, self._R, self._qb, self._Qf = Q, R, qb, Qf
        self.target, self.X_ref, self.U_ref = target, X_ref, U_ref
        self.angle_wrap = bool(angle_wrap)

    # constrained weight values (detached), as the solve-time closures of the reference see them
    def weights(self, dtype, device):
        if self.theta is not None:
            t = self.theta
            vals = [t.Q(), t.R(), t.qb(), t.Qf()]
        else:
            vals = [self._Q, self._R, self._qb, self._Qf]
        out = [torch.as_tensor(v).detach().to(device=device, dtype=dtype) for v in vals]
        return out[0].reshape(3), out[1].reshape(2), out[2].reshape(()), out[3].reshape(3)

    def _ref(self, x: Tensor, k: Optional[int]):
        if self.X_ref is None:
            return torch.as_tensor(self.target, dtype=x.dtype, device=x.device), None
        xr = self.X_ref[..., k if k is not None else -1, :].to(x.device)
        ur = None if k is None else self.U_ref[..., k, :].to(x.device)
        return xr, ur

    def _dx(self, x_hat: Tensor, xr: Tensor) -> Tensor:
        dx = x_hat[..., :3] - xr
        if self.angle_wrap:
            w = torch.atan2(torch.sin(dx[..., 2]), torch.cos(dx[..., 2]))
            dx = torch.cat([dx[..., :2], w.unsqueeze(-1)], dim=-1)
        return dx

    # single-point evaluations (API compatibility; the solver never calls these)
    def stage_cost(self, x_hat: Tensor, u: Tensor, k: int) -> Tensor:
        Q, R, qb, _ = self.weights(x_hat.dtype, x_hat.device)
        xr, ur = self._ref(x_hat, k)
        dx = self._dx(x_hat, xr)
        du = u if ur is None else u - ur
        return (Q * dx * dx).sum(-1) + (R * du * du).sum(-1) + qb * (x_hat[..., 3] * x_hat[..., 3])

    def terminal_cost(self, x_hat_N: Tensor) -> Tensor:
        _, _, qb, Qf = self.weights(x_hat_N.dtype, x_hat_N.device)
        xr, _ = self._ref(x_hat_N, None)
        dx = self._dx(x_hat_N, xr)
        return (Qf * dx * dx).sum(-1) + qb * (x_hat_N[..., 3] * x_hat_N[..., 3])

    def stage_derivs(self, x_hat: Tensor, u: Tensor, k: int):
        Q, R, qb, _ = self.weights(x_hat.dtype, x_hat.device)
        xr, ur = self._ref(x_hat, k)
        dx = self._dx(x_hat, xr)
        du = u if ur is None else u - ur
        l_x = torch.cat([2.0 * Q * dx, (2.0 * qb * x_hat[..., 3]).unsqueeze(-1)], dim=-1)
        l_u = 2.0 * R * du
        l_xx, l_uu, l_ux = self.stage_hess(x_hat, u, k)
        return l_x, l_u, l_xx, l_uu, l_ux

    def stage_hess(self, x_hat: Tensor, u: Tensor, k: int):
        Q, R, qb, _ = self.weights(x_hat.dtype, x_hat.device)
        l_xx = torch.diag(torch.cat([2.0 * Q, (2.0 * qb).view(1)]))
        l_uu = torch.diag(2.0 * R)
        l_ux = torch.zeros(2, 4, dtype=x_hat.dtype, device=x_hat.device)
        return l_xx, l_uu, l_ux

    def terminal_derivs(self, x_hat_N: Tensor):
        _, _, qb, Qf = self.weights(x_hat_N.dtype, x_hat_N.device)
        xr, _ = self._ref(x_hat_N, None)
        dx = self._dx(x_hat_N, xr)
        phi_x = torch.cat([2.0 * Qf * dx, (2.0 * qb * x_hat_N[..., 3]).unsqueeze(-1)], dim=-1)
        return phi_x, self.terminal_hess(x_hat_N)

    def terminal_hess(self, x_hat_N: Tensor):
        _, _, qb, Qf = self.weights(x_hat_N.dtype, x_hat_N.device)
        return torch.diag(torch.cat([2.0 * Qf, (2.0 * qb).view(1)]))


def _owner(fn, cls, what: str):
    obj = fn if isinstance(fn, cls) else getattr(fn, "__self__", None)
    if not isinstance(obj, cls):
        raise TypeError(f"{what} must be a {cls.__name__} (or one of its bound methods): arbitrary Python closures cannot "
                        "run inside the CUDA kernels and there is no generic-closure / CPU path")
    return obj


def _theta_vector(dyn: DubinsDBaSDynamics, cost: TrackingCost, dtype, device) -> Tensor:
    Q, R, qb, Qf = cost.weights(dtype, device)
    a, g, s = dyn.dbas_values()
    return torch.cat([Q, R, Qf, qb.view(1), torch.tensor([a, g, s], dtype=dtype, device=device)])


def _ocp(cfg: ILQRConfig, dyn: DubinsDBaSDynamics, jac_dyn: DubinsDBaSDynamics, cost: TrackingCost) -> OcpSpec:
    return OcpSpec(ilqr=cfg, eps_dyn=dyn.db_cfg.eps, eps_jac=jac_dyn.db_cfg.eps, angle_wrap=cost.angle_wrap)


def _batched(t: Tensor, nd: int, dev, dtype) -> Tensor:
    t = t.detach().to(device=dev, dtype=dtype)
    return t.unsqueeze(0) if t.dim() == nd else t


def _cost_refs(cost: TrackingCost, B: int, T: int, dev, dtype):
    if cost.X_ref is None:
        return dict(target=torch.as_tensor(cost.target, dtype=dtype, device=dev).reshape(3))
    Xr = _batched(cost.X_ref, 2, dev, dtype)[..., :3].expand(B, T + 1, 3).contiguous()
    Ur = _batched(cost.U_ref, 2, dev, dtype).expand(B, T, 2).contiguous()
    return dict(X_ref=Xr, U_ref=Ur)


def rollout(x0: Tensor, V: Tensor, *, f: Callable[[Tensor, Tensor], Tensor]) -> Tensor:
    """x0 [4] / [B,4], V [N,2] / [B,N,2] -> X [N+1,4] / [B,N+1,4]   (core/ddp.py:89-99)."""
    dyn = _owner(f, DubinsDBaSDynamics, "f")
    dev = _cuda_device(x0, V)
    xb, Vb = _batched(x0, 1, dev, x0.dtype), _batched(V, 2, dev, x0.dtype)
    a, g, s = dyn.dbas_values()
    theta = torch.tensor([0.0] * 9 + [a, g, s], dtype=x0.dtype, device=dev)
    ocp = OcpSpec(ILQRConfig(horizon=int(Vb.shape[1])), eps_dyn=dyn.db_cfg.eps, eps_jac=dyn.db_cfg.eps)
    X = ops.rollout_batched(dyn.system(None), ocp, xb.contiguous(), Vb.contiguous(), theta)
    return (X[0] if x0.dim() == 1 else X).to(x0.device)


def ilqr_solve(*, x0: Tensor, V_init: Tensor, cfg: ILQRConfig, f, f_jac=None, ctrl: Optional[BoxClampControl] = None,
               stage_cost, terminal_cost, stage_derivs, terminal_derivs, feasible_fn=None, debug: bool = False,
               debug_name: str = "ilqr"):
    """Batched counterpart of core/ddp.py:102-307; returns (X, V).  Raises FloatingPointError when any instance
    hits a non-finite value (the reference raises from `_ensure_finite`, :138-159)."""
    if feasible_fn is not None:
        raise NotImplementedError("feasible_fn is not supported (no caller of the reference passes it, ddp.py:114)")
    dyn = _owner(f, DubinsDBaSDynamics, "f")
    jac = _owner(f_jac, DubinsDBaSDynamics, "f_jac") if f_jac is not None else dyn
    cost = _owner(stage_cost, TrackingCost, "stage_cost")
    for fn, nm in ((terminal_cost, "terminal_cost"), (stage_derivs, "stage_derivs"), (terminal_derivs, "terminal_derivs")):
        if _owner(fn, TrackingCost, nm) is not cost:
            raise TypeError("stage_cost, terminal_cost, stage_derivs and terminal_derivs must belong to one TrackingCost")
    dev = _cuda_device(x0, V_init)
    dtype = x0.dtype
    xb, Vb = _batched(x0, 1, dev, dtype).contiguous(), _batched(V_init, 2, dev, dtype).contiguous()
    B, T = xb.shape[0], int(cfg.horizon)
    theta = _theta_vector(dyn, cost, dtype, dev)
    res = ops.ilqr_solve_batched(dyn.system(ctrl), _ocp(cfg, dyn, jac, cost), xb, Vb, theta,
                                 **_cost_refs(cost, B, T, dev, dtype))
    if bool((res.status != 0).any()):
        raise FloatingPointError(ops.failure_message(debug_name, res))
    X, V = (res.X[0], res.V[0]) if x0.dim() == 1 else (res.X, res.V)
    return X.to(x0.device), V.to(x0.device)


@dataclass(frozen=True)
class SensitivityResult:
    """core/ddp.py:310-314"""
    delta_X: Tensor
    delta_V: Tensor
    delta_lambda: Tensor


def _upper(g, X: Tensor, n: int, last=None) -> Optional[Tensor]:
    """Upper-level gradients: a tensor [.., n, d], or the reference's callables g(X[k], k) (+ g_N(X[N]))."""
    if g is None or isinstance(g, Tensor):
        return g
    rows = [g(X[..., k, :], k) for k in range(n)]
    if last is not None:
        rows.append(last(X[..., n, :]))
    return torch.stack(rows, dim=-2)


def ddp_sensitivity(*, X: Tensor, V: Tensor, f, f_jac=None, ctrl: Optional[BoxClampControl] = None, stage_hess,
                    terminal_hess, upper_grad_x, upper_grad_u, upper_grad_xN=None) -> SensitivityResult:
    """Batched counterpart of core/ddp.py:317-427.  upper_grad_* are the reference's callables
    (upper_grad_x(X[k], k), upper_grad_u(V[k], k), upper_grad_xN(X[N])) or tensors [.., N+1, 4] / [.., N, 2]."""
    dyn = _owner(f, DubinsDBaSDynamics, "f")
    jac = _owner(f_jac, DubinsDBaSDynamics, "f_jac") if f_jac is not None else dyn
    cost = _owner(stage_hess, TrackingCost, "stage_hess")
    dev = _cuda_device(X, V)
    dtype = X.dtype
    single = X.dim() == 2
    Xb, Vb = _batched(X, 2, dev, dtype).contiguous(), _batched(V, 2, dev, dtype).contiguous()
    B, T = Xb.shape[0], Vb.shape[1]
    gx = _upper(upper_grad_x, Xb, T, upper_grad_xN)
    gu = _upper(upper_grad_u, Vb, T)
    gx = _batched(gx, 2, dev, dtype).expand(B, T + 1, 4).contiguous()
    gu = None if gu is None else _batched(gu, 2, dev, dtype).expand(B, T, 2).contiguous()
    cfg = ILQRConfig(horizon=T)
    theta = _theta_vector(dyn, cost, dtype, dev)
    r = ops.ddp_sensitivity_batched(dyn.system(ctrl), _ocp(cfg, dyn, jac, cost), Xb, Vb, theta, upper_grad_x=gx,
                                    upper_grad_u=gu, want_grads=False, target=torch.zeros(3, dtype=dtype, device=dev))
    out = (r.delta_X, r.delta_V, r.delta_lambda)
    if single:
        out = tuple(t[0] for t in out)
    return SensitivityResult(*(t.to(X.device) for t in out))


@dataclass(frozen=True)
class IFTInputs:
    """core/ift.py:10-19"""
    X: Tensor
    V: Tensor
    delta_X: Tensor
    delta_V: Tensor
    delta_lambda: Tensor


def ift_gradient(*, inputs: IFTInputs, theta_tensors: Sequence[Tensor], xi_fn=None, f_fn, stage_cost_fn,
                 terminal_cost_fn) -> List[Optional[Tensor]]:
    """grad_theta L = sum_k (l_theta_x dx_k + l_theta_u du_k + f_theta^T dlambda_{k+1}) + phi_theta_x dx_N
    (core/ift.py:35-92; the xi term vanishes because the reference passes a detached x0, tube_mpc.py:498, 595).

    The reference differentiates through its closures with autograd; here the closed forms run in
    one kernel.  `f_fn` / `stage_cost_fn` must be built with `theta=` (a NominalTheta / AuxiliaryTheta holding the raw
    tensors); `theta_tensors` may list those raw tensors and the cost's X_ref / U_ref, in any order.  Entries that do
    not require grad give None, like `_grad_wrt_tensors` (core/ift.py:22-32).  For a batch the gradients of the
    shared raw parameters are summed over the instances; X_ref / U_ref gradients keep their batch dimension."""
    dyn = _owner(f_fn, DubinsDBaSDynamics, "f_fn")
    cost = _owner(stage_cost_fn, TrackingCost, "stage_cost_fn")
    if _owner(terminal_cost_fn, TrackingCost, "terminal_cost_fn") is not cost:
        raise TypeError("stage_cost_fn and terminal_cost_fn must belong to one TrackingCost")
    th = cost.theta if cost.theta is not None else dyn.theta
    if th is None:
        raise TypeError("ift_gradient needs the raw parameters: build the cost / dynamics with theta=NominalTheta|AuxiliaryTheta")
    X = inputs.X
    dev = _cuda_device(X)
    dtype = X.dtype
    single = X.dim() == 2
    b2 = lambda t: _batched(t, 2, dev, dtype).contiguous()
    Xb, Vb, dX, dV, dL = b2(inputs.X), b2(inputs.V), b2(inputs.delta_X), b2(inputs.delta_V), b2(inputs.delta_lambda)
    B, T = Xb.shape[0], Vb.shape[1]
    theta = _theta_vector(dyn, cost, dtype, dev)
    want_ref = cost.X_ref is not None and any(t is cost.X_ref or t is cost.U_ref for t in theta_tensors)
    cfg = ILQRConfig(horizon=T)
    r = ops.ift_accumulate_batched(dyn.system(None), _ocp(cfg, dyn, dyn, cost), Xb, Vb, theta, dX, dV, dL,
                                   want_ref_grads=want_ref, **_cost_refs(cost, B, T, dev, dtype))
    g_raw = (ops.reduce_columns(r.g_theta) if B > 1 else r.g_theta[0]) * th.chain().to(dev)
    by_id = {id(th.Q_raw): g_raw[0:3], id(th.R_raw): g_raw[3:5], id(th.Qf_raw): g_raw[5:8], id(th.qb_raw): g_raw[8],
             id(th.alpha_raw): g_raw[9], id(th.gamma_raw): g_raw[10]}
    if isinstance(th, NominalTheta):
        by_id[id(th.tight_raw)] = g_raw[11]
    if want_ref:
        by_id[id(cost.X_ref)] = r.g_xref[0] if single else r.g_xref
        by_id[id(cost.U_ref)] = r.g_uref[0] if single else r.g_uref
    out: List[Optional[Tensor]] = []
    for t in theta_tensors:
        g = by_id.get(id(t))
        out.append(None if (g is None or not t.requires_grad) else g.reshape(t.shape).to(t.device))
    return out
