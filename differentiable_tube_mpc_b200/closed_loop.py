"""Batched closed loop on the device: B tubes advance in lock-step for H steps without a host round trip.

One step = what the body of the reference's loop does (core/tube_mpc.py:803-1023 paper mode, :251-638 full mode):
solve nominal -> solve auxiliary -> upper loss -> sensitivity -> parameter gradient -> Algorithm-2 update ->
plant step with disturbance, DBaS propagation of the undisturbed successor, nominal propagation, warm-start shift.
Parameters are shared by the batch (Algorithm 2 descends the batch-summed loss; with B = 1 this is the reference);
under torch.distributed the summed gradient is all-reduced, so every rank applies the same update.

Nothing in `step()` synchronises with the host: logs are device tensors, numerical failures are OR-ed into a device
flag and raised by `check()` / at the end of `run()`.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional

import torch
from torch import Tensor

from . import ops
from .adaptation import NAMES_THETA, NAMES_THETA_BAR, PaperAdapter, apply_update, split_raw_grads
from .engine import BatchedTubeMPC, allreduce_shared
from .spec import TubeProblem, is_paper_mode


class ClosedLoop:
    def __init__(self, cfg: Dict[str, Any], *, device, batch: Optional[int] = None, n_log: int = 1,
                 disturbance_fn: Optional[Callable[[int], Tensor]] = None, x0: Optional[Tensor] = None,
                 x_bar0: Optional[Tensor] = None):
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("differentiable_tube_mpc_b200 runs on CUDA devices only (no CPU fallback)")
        self.paper = is_paper_mode(cfg)
        self.dtype = torch.float64 if bool(cfg.get("use_float64", False)) else torch.float32
        self.prob = TubeProblem.from_config(cfg, mode="paper" if self.paper else "full")
        s = cfg["system"]
        self.N, self.H = int(s["horizon_N"]), int(s["task_horizon_H"])
        self.B = B = int(batch if batch is not None else cfg.get("batch", 1))
        self.n_log = min(int(n_log), B)
        self.eng = BatchedTubeMPC(self.prob, device=self.device, dtype=self.dtype)
        dev, dt = self.device, self.dtype
        t_ = lambda v: torch.tensor(v, dtype=dt, device=dev)
        ad = cfg.get("adaptation", {})
        if x0 is None:   # start at the origin facing upper-right (reference :95-97, 772-773)
            x0 = t_(cfg.get("x0", [0.0, 0.0, 0.7853981633974483])).expand(B, 3)
        x = x0.to(device=dev, dtype=dt).contiguous()
        x_bar = (x if x_bar0 is None else x_bar0.to(device=dev, dtype=dt)).contiguous()
        if self.paper:
            self.theta_n = t_(self.prob.theta_nom0)
            a0 = self.prob.theta_aux0
            self.adapter = PaperAdapter(Qa=t_(a0[0:3]), Ra=t_(a0[3:5]), qba=t_(a0[8]), eta=float(ad.get("lr_eta", 1e-2)),
                                        mom=float(ad.get("momentum", 0.9)))
            theta_a = self.adapter.theta()
        else:
            self.theta_bar, self.theta = self.prob.make_thetas(dev, dt)
            self.adapt_nominal = bool(ad.get("adapt_nominal", True))
            self.adapt_ancillary = bool(ad.get("adapt_ancillary", True))
            self.upd = dict(lr=float(ad.get("lr_eta", 1e-3)), momentum=float(ad.get("momentum", 0.0)),
                            clip_norm=float(ad.get("grad_clip_norm", 0.0)), project_params=bool(ad.get("project_params", False)))
            self.adapt_steps = int(ad.get("steps", 1))
            self.v_theta = [torch.zeros_like(p) for p in self.theta.tensors()]
            self.v_theta_bar = [torch.zeros_like(p) for p in self.theta_bar.tensors()]
            self.theta_n, theta_a = self.theta_bar.constrained(), self.theta.constrained()
        b = self.eng.init_barrier(x, theta_a, nominal=False)
        b_bar = self.eng.init_barrier(x_bar, self.theta_n, nominal=True)
        self.xh = torch.cat([x, b.view(B, 1)], dim=1).contiguous()
        self.xh_bar = torch.cat([x_bar, b_bar.view(B, 1)], dim=1).contiguous()
        self.Vn_ws = torch.zeros(B, self.N, 2, dtype=dt, device=dev)
        self.Va_ws = torch.zeros(B, self.N, 2, dtype=dt, device=dev)
        self.sample_w = disturbance_fn or self._device_disturbance()
        self.failed = torch.zeros((), dtype=torch.int32, device=dev)
        self.t = 0
        self.log: Dict[str, list] = {k: [] for k in ("x_real", "u_real", "x_bar", "u_bar", "b_real", "loss", "Qa", "Ra", "qba")}

    def _device_disturbance(self) -> Callable[[int], Tensor]:
        """w_t = low + (high - low) U(0,1)^3 per instance, drawn on the device (reference core/systems/dubins.py:57-65)."""
        d = self.cfg["system"]["disturbance"]
        low = torch.tensor(d["w_low"], dtype=self.dtype, device=self.device)
        high = torch.tensor(d["w_high"], dtype=self.dtype, device=self.device)
        gen = torch.Generator(device=self.device).manual_seed(int(self.cfg.get("seed", 0)))
        return lambda t: low + (high - low) * torch.rand(self.B, 3, dtype=self.dtype, device=self.device, generator=gen)

    # ------------------------------------------------------------------
    def step(self) -> None:
        p, eng, B, N, dev, dt = self.prob, self.eng, self.B, self.N, self.device, self.dtype
        if self.paper:
            theta_a = self.adapter.theta()
            st = allreduce_shared(eng.step(self.xh_bar, self.xh, self.Vn_ws, self.Va_ws, self.theta_n, theta_a, full=False))
            rn, ra, loss = st.nominal, st.auxiliary, st.loss
            self.adapter.update(st.g_aux_sum[:12])   # momentum + projection (reference :978-984)
            th_n_plant, th_a_plant = self.theta_n, theta_a   # DBaS entries are fixed in paper mode (alpha = gamma = 0)
        else:
            theta, theta_bar = self.theta, self.theta_bar
            theta_n, theta_a = theta_bar.constrained(), theta.constrained()
            rn, ra = ops.tube_solve_batched(p.system, p.nominal, p.auxiliary, self.xh_bar, self.xh, self.Vn_ws, self.Va_ws,
                                            theta_n, theta_a, workspace=eng.workspace)
            sa = ops.ddp_sensitivity_batched(p.system, p.auxiliary, ra.X, ra.V, theta_a, X_bar=rn.X, X_ref=rn.X, U_ref=rn.V,
                                             want_grads=True, want_ref_grads=True, status=rn.status, workspace=eng.workspace)
            loss = sa.loss
            for s in range(self.adapt_steps):
                if s > 0:   # later adaptation steps re-accumulate with the updated parameters at the same optimum
                    sa2 = ops.ift_accumulate_batched(p.system, p.auxiliary, ra.X, ra.V, theta.constrained(), sa.delta_X,
                                                     sa.delta_V, sa.delta_lambda, X_ref=rn.X, U_ref=rn.V, want_ref_grads=True)
                    sa.g_theta, sa.g_xref, sa.g_uref = sa2.g_theta, sa2.g_xref, sa2.g_uref
                g_sum = _allreduce(ops.reduce_columns(sa.g_theta, eng.workspace))
                if self.adapt_ancillary:
                    apply_update(theta.tensors(), split_raw_grads(g_sum * theta.chain(), False), self.v_theta, NAMES_THETA,
                                 **self.upd)
                if not self.adapt_nominal:
                    continue
                gx = torch.cat([sa.g_xref, torch.zeros(B, N + 1, 1, dtype=dt, device=dev)], dim=2)
                sn = ops.ddp_sensitivity_batched(p.system, p.nominal, rn.X, rn.V, theta_n, upper_grad_x=gx,
                                                 upper_grad_u=sa.g_uref, want_grads=(s == 0), status=rn.status,
                                                 workspace=eng.workspace)
                if s > 0:
                    sn = ops.ift_accumulate_batched(p.system, p.nominal, rn.X, rn.V, theta_bar.constrained(), sn.delta_X,
                                                    sn.delta_V, sn.delta_lambda)
                g_sum = _allreduce(ops.reduce_columns(sn.g_theta, eng.workspace))
                apply_update(theta_bar.tensors(), split_raw_grads(g_sum * theta_bar.chain(), True), self.v_theta_bar,
                             NAMES_THETA_BAR, **self.upd)
            # the plant and the nominal state propagate with the UPDATED DBaS parameters (reference :609, 614)
            th_n_plant, th_a_plant = theta_bar.constrained(), theta.constrained()
        self.failed |= (rn.status != 0).any().to(torch.int32) * (self.t + 1) * (self.failed == 0)

        # device-side log of the first n_log instances (pre-step values, reference :619-629 / :1004-1013)
        n = self.n_log
        lg = self.log
        lg["x_real"].append(self.xh[:n, :3].clone()); lg["b_real"].append(self.xh[:n, 3].clone())
        lg["x_bar"].append(self.xh_bar[:n, :3].clone())
        lg["u_real"].append(ra.V[:n, 0].clone()); lg["u_bar"].append(rn.V[:n, 0].clone())
        lg["loss"].append(loss[:n].clone())
        if self.paper:
            lg["Qa"].append(self.adapter.Qa.clone()); lg["Ra"].append(self.adapter.Ra.clone())
            lg["qba"].append(self.adapter.qba.clone())
        elif self.adapt_ancillary:
            lg["Qa"].append(self.theta.Q().detach().clone()); lg["Ra"].append(self.theta.R().detach().clone())
            lg["qba"].append(self.theta.qb().detach().clone())

        # true system step with disturbance, barrier of the undisturbed successor, nominal propagation, warm-start
        # shift (reference :607-638 / :990-1023)
        ops.plant_step_batched(p.system, p.nominal, p.auxiliary, rn.V, ra.V, th_n_plant, th_a_plant, self.sample_w(self.t),
                               self.xh, self.xh_bar, self.Vn_ws, self.Va_ws)
        self.t += 1

    def check(self) -> None:
        """Raises what the reference would have raised inside its solver (FloatingPointError, core/ddp.py:142)."""
        f = int(self.failed.item())
        if f:
            raise FloatingPointError(f"tube MPC: non-finite detected at t={f - 1}")

    def run(self, H: Optional[int] = None, progress: bool = False) -> Dict[str, Tensor]:
        H = self.H if H is None else int(H)
        for t in range(H):
            if progress and (self.t % 25) == 0:
                print(f"[step {self.t}/{H}] running...", flush=True)
            self.step()
        self.check()
        return self.logs()

    def logs(self) -> Dict[str, Tensor]:
        """Stacked device logs: x_real [t,n,3], u_real [t,n,2], x_bar, u_bar, b_real [t,n], loss [t,n], Qa [t,3], ..."""
        return {k: torch.stack(v) for k, v in self.log.items() if v}


def _allreduce(g: Tensor) -> Tensor:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
    return g
