"""Batched operators over the C ABI: ilqr_solve, tube_solve, ddp_sensitivity, ift accumulation, reductions.

Every function takes CUDA tensors (instance-major, contiguous), launches on torch's current stream and returns
tensors; nothing here computes on the CPU.  Reference entry points each operator replaces:

  ilqr_solve_batched       core/ddp.py:102-307   ilqr_solve
  tube_solve_batched       core/tube_mpc.py:844-909 / :305-399 (nominal then auxiliary solve)
  ddp_sensitivity_batched  core/ddp.py:317-427   ddp_sensitivity   (+ core/ift.py:35-92 fused)
  ift_accumulate_batched   core/ift.py:35-92     ift_gradient (closed forms)
  rollout_batched          core/ddp.py:89-99     rollout
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .spec import OcpSpec, SystemSpec

_SUFFIX = {torch.float64: "f64", torch.float32: "f32"}


def _suffix(dtype) -> str:
    try:
        return _SUFFIX[dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {dtype}; the path computes in float64 or float32") from None


def _ptr(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk(t: Tensor, shape, dtype, name: str) -> Tensor:
    if t.device.type != "cuda":
        raise ValueError(f"{name} must be a CUDA tensor (no CPU path exists)")
    if t.dtype != dtype:
        raise TypeError(f"{name} has dtype {t.dtype}, expected {dtype}")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t.contiguous()


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _theta_arg(theta: Tensor, B: int, dtype) -> Tuple[Tensor, int]:
    if theta.dim() == 1:
        return _chk(theta, (_lib.NTHETA,), dtype, "theta"), 0
    return _chk(theta, (B, _lib.NTHETA), dtype, "theta"), _lib.NTHETA


def _target_arg(target: Tensor, B: int, dtype) -> Tuple[Tensor, int]:
    if target.dim() == 1:
        return _chk(target, (3,), dtype, "target"), 0
    return _chk(target, (B, 3), dtype, "target"), 3


class Workspace:
    """Caller-owned scratch for the C ABI, grown on demand and reused across calls (per device)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._buf: Dict[str, Tensor] = {}

    def get(self, key: str, nbytes: int) -> Tensor:
        b = self._buf.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._buf[key] = b
        return b


_const_cache: Dict[tuple, Tensor] = {}


def _const_tensor(values, dtype, device) -> Tensor:
    """Small host constant (the fixed target) as a device tensor, uploaded once per (values, dtype, device): a
    pageable host-to-device copy per call would serialise the stream and cannot be captured into a CUDA graph."""
    device = torch.device(device)
    key = (tuple(float(v) for v in values), dtype, device.type, device.index if device.index is not None
           else torch.cuda.current_device())
    t = _const_cache.get(key)
    if t is None:
        t = torch.tensor(list(key[0]), dtype=dtype, device=device)
        _const_cache[key] = t
    return t


_default_ws: Dict[tuple, Workspace] = {}


def default_workspace(device) -> Workspace:
    """The scratch every operator uses when the caller passes no `workspace`: one buffer set per (device, current
    stream), so calls that are not stream-ordered against each other (e.g. an autograd backward on a side stream)
    never share scratch.  Host threads that enqueue on the SAME stream concurrently must still pass their own
    `Workspace` (as `BatchedTubeMPC` does per engine)."""
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, int(torch.cuda.current_stream(idx).cuda_stream))
    if key not in _default_ws:
        _default_ws[key] = Workspace(torch.device("cuda", idx))
    return _default_ws[key]


@dataclass
class SolveResult:
    X: Tensor                  # [B,T+1,4]
    V: Tensor                  # [B,T,2]
    iters: Tensor              # [B] int32
    status: Tensor             # [B] int32
    K: Optional[Tensor] = None        # [B,T,2,4] gains of the last backward pass
    k: Optional[Tensor] = None        # [B,T,2]
    alpha_idx: Optional[Tensor] = None  # [B,max_iter] int32, -1 beyond iters
    costs: Optional[Tensor] = None      # [B,max_iter], nan beyond iters
    margins: Optional[Tensor] = None    # [B,max_iter]
    fail_at: Optional[Tensor] = None    # [B,2] int32 (iteration, step) of the first non-finite value, (-1,-1) if clean


def failure_message(name: str, res: "SolveResult", shift: int = 0) -> Optional[str]:
    """What the reference prints before it raises (core/ddp.py:138-159 `[NUMERIC-FAIL] name: ... non-finite at ...`):
    the first failing instance with the iteration / horizon step where the first non-finite value appeared."""
    bad = ((res.status >> shift) & 15) != 0
    if not bool(bad.any()):
        return None
    i = int(torch.nonzero(bad)[0])
    where = ""
    if res.fail_at is not None:
        it, k = (int(v) for v in res.fail_at[i].tolist())
        where = f" at (it={it}, k={k})" if k >= 0 else f" at it={it} (line-search / rollout cost)"
    return f"[NUMERIC-FAIL] {name}: non-finite detected{where} in instance {i} ({int(bad.sum())} of {bad.numel()} failed)"


def _alloc_solve(B, T, max_iter, dtype, device, gains, trace, status=None):
    out = dict(
        X=torch.empty(B, T + 1, 4, dtype=dtype, device=device), V=torch.empty(B, T, 2, dtype=dtype, device=device),
        iters=torch.zeros(B, dtype=torch.int32, device=device),
        status=torch.zeros(B, dtype=torch.int32, device=device) if status is None else status,
        fail_at=torch.empty(B, 2, dtype=torch.int32, device=device))
    if gains:
        out["K"] = torch.zeros(B, T, 2, 4, dtype=dtype, device=device)
        out["k"] = torch.zeros(B, T, 2, dtype=dtype, device=device)
    if trace:
        out["alpha_idx"] = torch.full((B, max_iter), -1, dtype=torch.int32, device=device)
        out["costs"] = torch.full((B, max_iter), float("nan"), dtype=dtype, device=device)
        out["margins"] = torch.full((B, max_iter), float("nan"), dtype=dtype, device=device)
    return out


def _fill_io(io: _lib.SolveIO, x0, V_init, theta, ts, target, gs, Xref, Uref, out):
    io.x0, io.V_init, io.theta, io.theta_stride = _ptr(x0), _ptr(V_init), _ptr(theta), ts
    io.target, io.target_stride, io.Xref, io.Uref = _ptr(target), gs, _ptr(Xref), _ptr(Uref)
    io.X, io.V, io.K, io.kff = _ptr(out["X"]), _ptr(out["V"]), _ptr(out.get("K")), _ptr(out.get("k"))
    io.iters, io.status, io.alpha_idx = _ptr(out["iters"]), _ptr(out["status"]), _ptr(out.get("alpha_idx"))
    io.costs, io.margins, io.fail_at = _ptr(out.get("costs")), _ptr(out.get("margins")), _ptr(out.get("fail_at"))


def ilqr_solve_batched(system: SystemSpec, ocp: OcpSpec, x0: Tensor, V_init: Tensor, theta: Tensor, *,
                       target: Optional[Tensor] = None, X_ref: Optional[Tensor] = None, U_ref: Optional[Tensor] = None,
                       gains: bool = False, trace: bool = False, workspace: Optional[Workspace] = None) -> SolveResult:
    """B independent iLQR solves.  x0 [B,4] (x, y, heading, b), V_init [B,T,2], theta [B,12] or [12] constrained
    values; tracks `target` [3] / [B,3] (nominal OCP) or X_ref [B,T+1,3], U_ref [B,T,2] (auxiliary OCP)."""
    L = _lib.load()
    dtype, device = x0.dtype, x0.device
    sfx = _suffix(dtype)
    B, T = int(x0.shape[0]), int(ocp.ilqr.horizon)
    x0 = _chk(x0, (B, 4), dtype, "x0")
    V_init = _chk(V_init, (B, T, 2), dtype, "V_init")
    theta, ts = _theta_arg(theta, B, dtype)
    gs = 0
    if X_ref is not None:
        X_ref = _chk(X_ref, (B, T + 1, 3), dtype, "X_ref")
        U_ref = _chk(U_ref, (B, T, 2), dtype, "U_ref")
    else:
        if target is None:
            target = _const_tensor(system.y_target, dtype, device)
        target, gs = _target_arg(target, B, dtype)
    out = _alloc_solve(B, T, ocp.ilqr.max_iter, dtype, device, gains, trace)
    io = _lib.SolveIO()
    _fill_io(io, x0, V_init, theta, ts, target, gs, X_ref, U_ref, out)
    env_c, ocp_c = system.to_c(), ocp.to_c()
    ws = workspace or default_workspace(device)
    with torch.cuda.device(device):
        nbytes = L.dtmpc_solve_workspace_bytes(T, B, ocp_c.n_alpha, x0.element_size())
        buf = ws.get("solve", nbytes)
        rc = getattr(L, f"dtmpc_ilqr_solve_{sfx}")(C.byref(env_c), C.byref(ocp_c), B, C.byref(io), _ptr(buf),
                                                   buf.numel(), _stream())
    _lib.check(rc, "dtmpc_ilqr_solve")
    return SolveResult(**out)


def tube_solve_batched(system: SystemSpec, nominal: OcpSpec, auxiliary: OcpSpec, x0_nom: Tensor, x0_aux: Tensor,
                       Vn_init: Tensor, Va_init: Tensor, theta_nom: Tensor, theta_aux: Tensor, *,
                       target: Optional[Tensor] = None, gains: bool = False, trace: bool = False,
                       workspace: Optional[Workspace] = None) -> Tuple[SolveResult, SolveResult]:
    """Nominal solve followed by the auxiliary solve that tracks it, fused in one persistent launch.
    status (shared tensor) = nominal bits | auxiliary bits << 4."""
    L = _lib.load()
    dtype, device = x0_nom.dtype, x0_nom.device
    sfx = _suffix(dtype)
    B, T = int(x0_nom.shape[0]), int(nominal.ilqr.horizon)
    x0_nom = _chk(x0_nom, (B, 4), dtype, "x0_nom")
    x0_aux = _chk(x0_aux, (B, 4), dtype, "x0_aux")
    Vn_init = _chk(Vn_init, (B, T, 2), dtype, "Vn_init")
    Va_init = _chk(Va_init, (B, T, 2), dtype, "Va_init")
    theta_nom, tsn = _theta_arg(theta_nom, B, dtype)
    theta_aux, tsa = _theta_arg(theta_aux, B, dtype)
    if target is None:
        target = _const_tensor(system.y_target, dtype, device)
    target, gs = _target_arg(target, B, dtype)
    on = _alloc_solve(B, T, nominal.ilqr.max_iter, dtype, device, gains, trace)
    oa = _alloc_solve(B, T, auxiliary.ilqr.max_iter, dtype, device, gains, trace, status=on["status"])
    ion, ioa = _lib.SolveIO(), _lib.SolveIO()
    _fill_io(ion, x0_nom, Vn_init, theta_nom, tsn, target, gs, None, None, on)
    _fill_io(ioa, x0_aux, Va_init, theta_aux, tsa, None, 0, None, None, oa)
    env_c, n_c, a_c = system.to_c(), nominal.to_c(), auxiliary.to_c()
    ws = workspace or default_workspace(device)
    with torch.cuda.device(device):
        nbytes = L.dtmpc_solve_workspace_bytes(T, B, max(n_c.n_alpha, a_c.n_alpha), x0_nom.element_size())
        buf = ws.get("solve", nbytes)
        rc = getattr(L, f"dtmpc_tube_solve_{sfx}")(C.byref(env_c), C.byref(n_c), C.byref(a_c), B, C.byref(ion),
                                                   C.byref(ioa), _ptr(buf), buf.numel(), _stream())
    _lib.check(rc, "dtmpc_tube_solve")
    return SolveResult(**on), SolveResult(**oa)


def tube_step_batched(system: SystemSpec, nominal: OcpSpec, auxiliary: OcpSpec, x0_nom: Tensor, x0_aux: Tensor,
                      Vn_init: Tensor, Va_init: Tensor, theta_nom: Tensor, theta_aux: Tensor, *, full: bool,
                      target: Optional[Tensor] = None, gains: bool = False, trace: bool = False,
                      workspace: Optional[Workspace] = None):
    """One closed-loop step in one library call (`dtmpc_tube_step_*`): nominal solve, auxiliary solve, auxiliary
    sensitivity + gradients with the built-in tracking loss, and (full = True) the nominal sensitivity + gradients
    driven by dL/dx_ref, dL/du_ref of the auxiliary sweep (core/tube_mpc.py:844-976, :305-602).  Same results, bit for
    bit, as tube_solve_batched + ddp_sensitivity_batched (x2); every chunk of the batch runs the whole chain on its own
    stream.  Returns (nominal, auxiliary, sens_aux, sens_nom or None)."""
    L = _lib.load()
    dtype, device = x0_nom.dtype, x0_nom.device
    sfx = _suffix(dtype)
    B, T = int(x0_nom.shape[0]), int(nominal.ilqr.horizon)
    x0_nom = _chk(x0_nom, (B, 4), dtype, "x0_nom")
    x0_aux = _chk(x0_aux, (B, 4), dtype, "x0_aux")
    Vn_init = _chk(Vn_init, (B, T, 2), dtype, "Vn_init")
    Va_init = _chk(Va_init, (B, T, 2), dtype, "Va_init")
    theta_nom, tsn = _theta_arg(theta_nom, B, dtype)
    theta_aux, tsa = _theta_arg(theta_aux, B, dtype)
    if target is None:
        target = _const_tensor(system.y_target, dtype, device)
    target, gs = _target_arg(target, B, dtype)
    on = _alloc_solve(B, T, nominal.ilqr.max_iter, dtype, device, gains, trace)
    oa = _alloc_solve(B, T, auxiliary.ilqr.max_iter, dtype, device, gains, trace, status=on["status"])
    ion, ioa = _lib.SolveIO(), _lib.SolveIO()
    _fill_io(ion, x0_nom, Vn_init, theta_nom, tsn, target, gs, None, None, on)
    _fill_io(ioa, x0_aux, Va_init, theta_aux, tsa, None, 0, None, None, oa)
    rn, ra = SolveResult(**on), SolveResult(**oa)

    def new_sens(X, V):
        return SensitivityResult(delta_X=torch.empty_like(X), delta_V=torch.empty_like(V), delta_lambda=torch.empty_like(X),
                                 g_theta=torch.empty(B, _lib.NTHETA, dtype=dtype, device=device))

    sa = new_sens(ra.X, ra.V)
    sa.loss = torch.empty(B, dtype=dtype, device=device)
    isa = _lib.SensIO()
    isa.X, isa.V, isa.theta, isa.theta_stride = _ptr(ra.X), _ptr(ra.V), _ptr(theta_aux), tsa
    isa.Xbar, isa.Xref, isa.Uref, isa.xref_row = _ptr(rn.X), _ptr(rn.X), _ptr(rn.V), 4
    isa.dX, isa.dV, isa.dL, isa.g_theta = _ptr(sa.delta_X), _ptr(sa.delta_V), _ptr(sa.delta_lambda), _ptr(sa.g_theta)
    isa.loss, isa.status = _ptr(sa.loss), _ptr(rn.status)
    sn, isn = None, None
    if full:
        sa.g_xref = torch.empty(B, T + 1, 3, dtype=dtype, device=device)
        sa.g_uref = torch.empty(B, T, 2, dtype=dtype, device=device)
        isa.gXref, isa.gUref = _ptr(sa.g_xref), _ptr(sa.g_uref)
        sn = new_sens(rn.X, rn.V)
        isn = _lib.SensIO()
        isn.X, isn.V, isn.theta, isn.theta_stride = _ptr(rn.X), _ptr(rn.V), _ptr(theta_nom), tsn
        isn.target, isn.target_stride = _ptr(target), gs
        isn.gx, isn.gx_row, isn.gu = _ptr(sa.g_xref), 3, _ptr(sa.g_uref)   # [dL/dx_ref, 0] (core/tube_mpc.py:520-535)
        isn.dX, isn.dV, isn.dL, isn.g_theta = _ptr(sn.delta_X), _ptr(sn.delta_V), _ptr(sn.delta_lambda), _ptr(sn.g_theta)
        isn.status = _ptr(rn.status)
    env_c, n_c, a_c = system.to_c(), nominal.to_c(), auxiliary.to_c()
    ws = workspace or default_workspace(device)
    with torch.cuda.device(device):
        nbytes = L.dtmpc_step_workspace_bytes(T, B, max(n_c.n_alpha, a_c.n_alpha), x0_nom.element_size())
        buf = ws.get("step", nbytes)
        rc = getattr(L, f"dtmpc_tube_step_{sfx}")(C.byref(env_c), C.byref(n_c), C.byref(a_c), B, C.byref(ion),
                                                  C.byref(ioa), C.byref(isa), C.byref(isn) if isn is not None else None,
                                                  _ptr(buf), buf.numel(), _stream())
    _lib.check(rc, "dtmpc_tube_step")
    return rn, ra, sa, sn


@dataclass
class SensitivityResult:
    """Batched counterpart of core/ddp.py:310-314 plus the fused IFT accumulation."""
    delta_X: Tensor            # [B,T+1,4]
    delta_V: Tensor            # [B,T,2]
    delta_lambda: Tensor       # [B,T+1,4]
    g_theta: Optional[Tensor] = None   # [B,12] d/d(constrained theta): Q, R, Qf, qb, alpha, gamma, tight
    g_xref: Optional[Tensor] = None    # [B,T+1,3]
    g_uref: Optional[Tensor] = None    # [B,T,2]
    loss: Optional[Tensor] = None      # [B] upper loss (built-in tracking loss only)


def ddp_sensitivity_batched(system: SystemSpec, ocp: OcpSpec, X: Tensor, V: Tensor, theta: Tensor, *,
                            upper_grad_x: Optional[Tensor] = None, upper_grad_u: Optional[Tensor] = None,
                            X_bar: Optional[Tensor] = None, target: Optional[Tensor] = None,
                            X_ref: Optional[Tensor] = None, U_ref: Optional[Tensor] = None, want_grads: bool = True,
                            want_ref_grads: bool = False, status: Optional[Tensor] = None,
                            workspace: Optional[Workspace] = None) -> SensitivityResult:
    """Sensitivity sweep at the optimum (X, V).  Upper-level gradients are either given
    (upper_grad_x [B,T+1,4], upper_grad_u [B,T,2] or None for zeros) or, with X_bar [B,T+1,4], the built-in loss
    sum|x - xbar|^2 + sum b^2 of core/tube_mpc.py:916-919.  The cost reference for the parameter gradients is
    X_ref [B,T+1,3 or 4] / U_ref, else `target`."""
    L = _lib.load()
    dtype, device = X.dtype, X.device
    sfx = _suffix(dtype)
    B, T = int(X.shape[0]), int(ocp.ilqr.horizon)
    X = _chk(X, (B, T + 1, 4), dtype, "X")
    V = _chk(V, (B, T, 2), dtype, "V")
    theta, ts = _theta_arg(theta, B, dtype)
    io = _lib.SensIO()
    io.X, io.V, io.theta, io.theta_stride = _ptr(X), _ptr(V), _ptr(theta), ts
    keep = [X, V, theta]
    if upper_grad_x is not None:
        upper_grad_x = _chk(upper_grad_x, (B, T + 1, 4), dtype, "upper_grad_x")
        io.gx = _ptr(upper_grad_x)
        if upper_grad_u is not None:
            upper_grad_u = _chk(upper_grad_u, (B, T, 2), dtype, "upper_grad_u")
            io.gu = _ptr(upper_grad_u)
    elif X_bar is not None:
        X_bar = _chk(X_bar, (B, T + 1, 4), dtype, "X_bar")
        io.Xbar = _ptr(X_bar)
    else:
        raise ValueError("either upper_grad_x or X_bar is required")
    if X_ref is not None:
        row = int(X_ref.shape[-1])
        X_ref = _chk(X_ref, (B, T + 1, row), dtype, "X_ref")
        U_ref = _chk(U_ref, (B, T, 2), dtype, "U_ref")
        io.Xref, io.Uref, io.xref_row = _ptr(X_ref), _ptr(U_ref), row
    else:
        if target is None:
            target = _const_tensor(system.y_target, dtype, device)
        target, gs = _target_arg(target, B, dtype)
        io.target, io.target_stride = _ptr(target), gs
    keep += [upper_grad_x, upper_grad_u, X_bar, X_ref, U_ref, target]
    res = SensitivityResult(delta_X=torch.empty_like(X), delta_V=torch.empty_like(V), delta_lambda=torch.empty_like(X))
    io.dX, io.dV, io.dL = _ptr(res.delta_X), _ptr(res.delta_V), _ptr(res.delta_lambda)
    if want_grads:
        res.g_theta = torch.empty(B, _lib.NTHETA, dtype=dtype, device=device)
        io.g_theta = _ptr(res.g_theta)
    if want_ref_grads:
        res.g_xref = torch.empty(B, T + 1, 3, dtype=dtype, device=device)
        res.g_uref = torch.empty(B, T, 2, dtype=dtype, device=device)
        io.gXref, io.gUref = _ptr(res.g_xref), _ptr(res.g_uref)
    if X_bar is not None and upper_grad_x is None:
        res.loss = torch.empty(B, dtype=dtype, device=device)
        io.loss = _ptr(res.loss)
    if status is not None:
        io.status = _ptr(status)
    env_c, ocp_c = system.to_c(), ocp.to_c()
    ws = workspace or default_workspace(device)
    with torch.cuda.device(device):
        nbytes = L.dtmpc_sens_workspace_bytes(T, B, X.element_size())
        buf = ws.get("sens", nbytes)
        rc = getattr(L, f"dtmpc_sensitivity_{sfx}")(C.byref(env_c), C.byref(ocp_c), B, C.byref(io), _ptr(buf),
                                                    buf.numel(), _stream())
    _lib.check(rc, "dtmpc_sensitivity")
    del keep
    return res


def ift_accumulate_batched(system: SystemSpec, ocp: OcpSpec, X: Tensor, V: Tensor, theta: Tensor, delta_X: Tensor,
                           delta_V: Tensor, delta_lambda: Tensor, *, target: Optional[Tensor] = None,
                           X_ref: Optional[Tensor] = None, U_ref: Optional[Tensor] = None,
                           want_ref_grads: bool = False) -> SensitivityResult:
    """Closed-form ift_gradient for caller-supplied sensitivities (core/ift.py:35-92)."""
    L = _lib.load()
    dtype, device = X.dtype, X.device
    sfx = _suffix(dtype)
    B, T = int(X.shape[0]), int(ocp.ilqr.horizon)
    X = _chk(X, (B, T + 1, 4), dtype, "X")
    V = _chk(V, (B, T, 2), dtype, "V")
    dX = _chk(delta_X, (B, T + 1, 4), dtype, "delta_X")
    dV = _chk(delta_V, (B, T, 2), dtype, "delta_V")
    dL = _chk(delta_lambda, (B, T + 1, 4), dtype, "delta_lambda")
    theta, ts = _theta_arg(theta, B, dtype)
    io = _lib.SensIO()
    io.X, io.V, io.theta, io.theta_stride = _ptr(X), _ptr(V), _ptr(theta), ts
    io.dX, io.dV, io.dL = _ptr(dX), _ptr(dV), _ptr(dL)
    if X_ref is not None:
        row = int(X_ref.shape[-1])
        X_ref = _chk(X_ref, (B, T + 1, row), dtype, "X_ref")
        U_ref = _chk(U_ref, (B, T, 2), dtype, "U_ref")
        io.Xref, io.Uref, io.xref_row = _ptr(X_ref), _ptr(U_ref), row
    else:
        if target is None:
            target = _const_tensor(system.y_target, dtype, device)
        target, gs = _target_arg(target, B, dtype)
        io.target, io.target_stride = _ptr(target), gs
    res = SensitivityResult(delta_X=dX, delta_V=dV, delta_lambda=dL,
                            g_theta=torch.empty(B, _lib.NTHETA, dtype=dtype, device=device))
    io.g_theta = _ptr(res.g_theta)
    if want_ref_grads:
        res.g_xref = torch.empty(B, T + 1, 3, dtype=dtype, device=device)
        res.g_uref = torch.empty(B, T, 2, dtype=dtype, device=device)
        io.gXref, io.gUref = _ptr(res.g_xref), _ptr(res.g_uref)
    env_c, ocp_c = system.to_c(), ocp.to_c()
    with torch.cuda.device(device):
        rc = getattr(L, f"dtmpc_ift_accumulate_{sfx}")(C.byref(env_c), C.byref(ocp_c), B, C.byref(io), _stream())
    _lib.check(rc, "dtmpc_ift_accumulate")
    return res


def reduce_columns(g: Tensor, workspace: Optional[Workspace] = None) -> Tensor:
    """Deterministic column sums of g [B,C] (fixed reduction tree, bitwise reproducible)."""
    L = _lib.load()
    if g.dim() == 1:
        g = g.view(-1, 1)
    g = g.contiguous()
    if g.device.type != "cuda":
        raise ValueError("g must be a CUDA tensor")
    sfx = _suffix(g.dtype)
    B, ncols = int(g.shape[0]), int(g.shape[1])
    out = torch.empty(ncols, dtype=g.dtype, device=g.device)
    ws = workspace or default_workspace(g.device)
    with torch.cuda.device(g.device):
        nbytes = L.dtmpc_reduce_workspace_bytes(B, ncols, g.element_size())
        buf = ws.get("reduce", max(nbytes, 16))
        rc = getattr(L, f"dtmpc_reduce_columns_{sfx}")(_ptr(g), B, ncols, _ptr(out), _ptr(buf), buf.numel(), _stream())
    _lib.check(rc, "dtmpc_reduce_columns")
    return out


def rollout_batched(system: SystemSpec, ocp: OcpSpec, x0: Tensor, V: Tensor, theta: Tensor) -> Tensor:
    """X[:,0] = x0, X[:,k+1] = f_hat(X[:,k], V[:,k])  (core/ddp.py:89-99 with the DBaS-augmented Dubins step)."""
    L = _lib.load()
    dtype, device = x0.dtype, x0.device
    sfx = _suffix(dtype)
    B, T = int(x0.shape[0]), int(V.shape[1])
    x0 = _chk(x0, (B, 4), dtype, "x0")
    V = _chk(V, (B, T, 2), dtype, "V")
    theta, ts = _theta_arg(theta, B, dtype)
    X = torch.empty(B, T + 1, 4, dtype=dtype, device=device)
    env_c = system.to_c()
    ocp_c = ocp.to_c()
    ocp_c.horizon = T
    with torch.cuda.device(device):
        rc = getattr(L, f"dtmpc_rollout_{sfx}")(C.byref(env_c), C.byref(ocp_c), B, _ptr(x0), _ptr(V), _ptr(theta), ts,
                                                _ptr(X), _stream())
    _lib.check(rc, "dtmpc_rollout")
    return X


def primitives_batched(system: SystemSpec, ocp: OcpSpec, xh: Tensor, u: Tensor, theta: Tensor,
                       target: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """Point evaluations of f_hat, the augmented jacobian, h, grad h, the stage cost, clamp and the active mask."""
    L = _lib.load()
    dtype, device = xh.dtype, xh.device
    sfx = _suffix(dtype)
    B = int(xh.shape[0])
    xh = _chk(xh, (B, 4), dtype, "xh")
    u = _chk(u, (B, 2), dtype, "u")
    theta, ts = _theta_arg(theta, B, dtype)
    if target is None:
        target = _const_tensor(system.y_target, dtype, device)
    target = _chk(target, (3,), dtype, "target")
    e = lambda *s: torch.empty(*s, dtype=dtype, device=device)
    out = dict(x_next=e(B, 4), A=e(B, 4, 4), B=e(B, 4, 2), h=e(B), grad_h=e(B, 3), cost=e(B), u_clamped=e(B, 2),
               active=torch.empty(B, 2, dtype=torch.int32, device=device), b0=e(B))
    env_c, ocp_c = system.to_c(), ocp.to_c()
    with torch.cuda.device(device):
        rc = getattr(L, f"dtmpc_primitives_{sfx}")(
            C.byref(env_c), C.byref(ocp_c), B, _ptr(xh), _ptr(u), _ptr(theta), ts, _ptr(target), _ptr(out["x_next"]),
            _ptr(out["A"]), _ptr(out["B"]), _ptr(out["h"]), _ptr(out["grad_h"]), _ptr(out["cost"]),
            _ptr(out["u_clamped"]), _ptr(out["active"]), _ptr(out["b0"]), _stream())
    _lib.check(rc, "dtmpc_primitives")
    out["active"] = out["active"].bool()
    return out


@dataclass
class DisturbanceSpec:
    """In-kernel disturbance w = low + (high - low) U[0,1)^3 (core/systems/dubins.py:57-65), Philox4x32-10 keyed by
    `seed` with counter (first_instance + i, step[0]): `step` is a device int64 tensor of one element, read at kernel
    run time (so a captured step can be replayed with a moving counter)."""
    low: Tuple[float, float, float]
    high: Tuple[float, float, float]
    seed: int
    step: Tensor
    first_instance: int = 0


def plant_step_batched(system: SystemSpec, nominal: OcpSpec, auxiliary: OcpSpec, Un: Tensor, Ua: Tensor,
                       theta_nom: Tensor, theta_aux: Tensor, w: Optional[Tensor], xh: Tensor, xh_bar: Tensor,
                       Vn_ws: Tensor, Va_ws: Tensor, rng: Optional[DisturbanceSpec] = None) -> None:
    """In-place closed-loop step (core/tube_mpc.py:990-1023): plant + disturbance (the tensor `w` [B,3], or drawn in
    the kernel from `rng`, or none), DBaS propagation of the undisturbed successor, nominal propagation, warm-start
    shift."""
    L = _lib.load()
    dtype, device = xh.dtype, xh.device
    sfx = _suffix(dtype)
    B, T = int(xh.shape[0]), int(nominal.ilqr.horizon)
    for t, shp, nm in ((Un, (B, T, 2), "Un"), (Ua, (B, T, 2), "Ua"), (xh, (B, 4), "xh"), (xh_bar, (B, 4), "xh_bar"),
                       (Vn_ws, (B, T, 2), "Vn_ws"), (Va_ws, (B, T, 2), "Va_ws")):
        if not t.is_contiguous():
            raise ValueError(f"{nm} must be contiguous (updated in place)")
        _chk(t, shp, dtype, nm)
    theta_nom, tsn = _theta_arg(theta_nom, B, dtype)
    theta_aux, tsa = _theta_arg(theta_aux, B, dtype)
    if w is not None:
        w = _chk(w, (B, 3), dtype, "w")
    env_c, n_c, a_c = system.to_c(), nominal.to_c(), auxiliary.to_c()
    rng_c = None
    if w is None and rng is not None:
        if rng.step.device != device or rng.step.dtype != torch.int64 or rng.step.numel() != 1:
            raise ValueError("rng.step must be a one-element int64 tensor on the compute device")
        rng_c = _lib.Disturbance()
        for c in range(3):
            rng_c.low[c], rng_c.high[c] = float(rng.low[c]), float(rng.high[c])
        rng_c.seed, rng_c.first_instance, rng_c.step = int(rng.seed) & (2 ** 64 - 1), int(rng.first_instance), rng.step.data_ptr()
    with torch.cuda.device(device):
        rc = getattr(L, f"dtmpc_plant_step_{sfx}")(C.byref(env_c), C.byref(n_c), C.byref(a_c), B, _ptr(Un), _ptr(Ua),
                                                   _ptr(theta_nom), tsn, _ptr(theta_aux), tsa, _ptr(w),
                                                   C.byref(rng_c) if rng_c is not None else None, _ptr(xh),
                                                   _ptr(xh_bar), _ptr(Vn_ws), _ptr(Va_ws), _stream())
    _lib.check(rc, "dtmpc_plant_step")


def math_selftest(x: Tensor) -> Tensor:
    """[n,6] = exp_nonpos(-|x|), log_ge1(1+|x|), log_1to2(1+frac|x|), sin x, cos x, rcp x of the library's fp64 functions."""
    L = _lib.load()
    x = _chk(x, tuple(x.shape), torch.float64, "x").reshape(-1)
    out = torch.empty(x.numel(), 6, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.dtmpc_math_selftest_f64(_ptr(x), x.numel(), _ptr(out), _stream())
    _lib.check(rc, "dtmpc_math_selftest")
    return out


def kernel_launches() -> int:
    """Number of CUDA kernels libdtmpc_b200 has enqueued in this process (dtmpc_kernel_launches)."""
    return int(_lib.load().dtmpc_kernel_launches())


def measure_fp64_peak_tflops(iters: int = 20000) -> float:
    return float(_lib.load().dtmpc_measure_fp64_peak_tflops(int(iters), _stream()))


def measure_fp32_peak_tflops(iters: int = 20000) -> float:
    return float(_lib.load().dtmpc_measure_fp32_peak_tflops(int(iters), _stream()))
