"""Batched closed loop on the device: B tubes advance in lock-step for H steps without a host round trip.

One step = what the body of the reference's loop does (core/tube_mpc.py:803-1023 paper mode, :251-638 full mode):
solve nominal -> solve auxiliary -> upper loss -> sensitivity -> parameter gradient -> Algorithm-2 update ->
plant step with disturbance, DBaS propagation of the undisturbed successor, nominal propagation, warm-start shift.
Parameters are shared by the batch (Algorithm 2 descends the batch-summed loss; with B = 1 this is the reference);
under torch.distributed the summed gradient is summed over the ranks in a fixed order, so every rank applies the
same update, bit for bit.

Everything a step touches lives in device memory and is updated in place: states, warm starts, the packed parameter /
momentum vectors (`adaptation`), the step counter, the logs (written at the device-side step index) and the failure
flag.  With the in-kernel disturbance generator (Philox, keyed by seed / global instance / step) the whole step is
therefore a fixed launch sequence and is captured ONCE into a CUDA graph (`graph=True`, the default on one rank):
H steps = H graph launches, no synchronisation until `check()` / the end of `run()`.  A host-side `disturbance_fn`
(the reference's torch RNG stream, used by the drop-in `run_closed_loop_experiment`) keeps the plain launch path.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional

import torch
from torch import Tensor

from . import ops
from .adaptation import PaperAdapter, RawAdapter
from .engine import BatchedTubeMPC, allreduce_shared, sum_over_ranks
from .spec import TubeProblem, is_paper_mode

_LOG_KEYS = ("x_real", "u_real", "x_bar", "u_bar", "b_real", "loss", "Qa", "Ra", "qba")


class ClosedLoop:
    def __init__(self, cfg: Dict[str, Any], *, device, batch: Optional[int] = None, n_log: int = 1,
                 disturbance_fn: Optional[Callable[[int], Tensor]] = None, x0: Optional[Tensor] = None,
                 x_bar0: Optional[Tensor] = None, graph: Optional[bool] = None, first_instance: int = 0):
        """batch: tubes on THIS rank; first_instance: global index of this rank's first tube (keys the disturbance
        stream, so a sharded run draws what the single-rank run draws); graph: capture the step into a CUDA graph
        (default: yes on one rank with the in-kernel disturbance, no otherwise)."""
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
        self.adapt_steps = 1
        if self.paper:
            self.theta_n = t_(self.prob.theta_nom0)
            a0 = self.prob.theta_aux0
            self.adapter = PaperAdapter(Qa=t_(a0[0:3]), Ra=t_(a0[3:5]), qba=t_(a0[8]), eta=float(ad.get("lr_eta", 1e-2)),
                                        mom=float(ad.get("momentum", 0.9)))
            self.theta_a = self.adapter.theta
            self.adapt_nominal, self.adapt_ancillary = False, True
        else:
            upd = dict(lr=float(ad.get("lr_eta", 1e-3)), momentum=float(ad.get("momentum", 0.0)),
                       clip_norm=float(ad.get("grad_clip_norm", 0.0)), project_params=bool(ad.get("project_params", False)))
            self.theta_bar = RawAdapter(t_(self.prob.theta_nom0), with_tight=True, **upd)          # nominal, raw
            self.theta = RawAdapter(t_(list(self.prob.theta_aux0) + [0.0]), with_tight=False, **upd)  # auxiliary, raw
            self.theta_n, self.theta_a = self.theta_bar.theta, self.theta.theta
            self.adapt_nominal = bool(ad.get("adapt_nominal", True))
            self.adapt_ancillary = bool(ad.get("adapt_ancillary", True))
            self.adapt_steps = int(ad.get("steps", 1))
        b = self.eng.init_barrier(x, self.theta_a, nominal=False)
        b_bar = self.eng.init_barrier(x_bar, self.theta_n, nominal=True)
        self.xh = torch.cat([x, b.view(B, 1)], dim=1).contiguous()
        self.xh_bar = torch.cat([x_bar, b_bar.view(B, 1)], dim=1).contiguous()
        self.Vn_ws = torch.zeros(B, self.N, 2, dtype=dt, device=dev)
        self.Va_ws = torch.zeros(B, self.N, 2, dtype=dt, device=dev)
        self.t_dev = torch.zeros(1, dtype=torch.int64, device=dev)      # closed-loop step counter, device side
        self.failed = torch.zeros((), dtype=torch.int64, device=dev)    # (first failing step + 1), 0 = clean
        self.sample_w = disturbance_fn
        d = cfg["system"]["disturbance"]
        self.rng = ops.DisturbanceSpec(low=tuple(d["w_low"]), high=tuple(d["w_high"]), seed=int(cfg.get("seed", 0)),
                                       step=self.t_dev, first_instance=int(first_instance))
        import torch.distributed as dist

        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.use_graph = (disturbance_fn is None and world == 1 and self.adapt_steps == 1) if graph is None else bool(graph)
        if self.use_graph and (disturbance_fn is not None or self.adapt_steps != 1):
            raise ValueError("a captured step needs the in-kernel disturbance and adaptation.steps == 1")
        self.t = 0
        self._cap = 0
        self._logs: Dict[str, Tensor] = {}
        self._graph = None
        self.last = None          # TubeStep of the most recent step (static tensors when graphed)

    # ------------------------------------------------------------------
    def _reserve_logs(self, steps: int) -> None:
        """Device log buffers for `steps` more steps (never reallocated while a captured graph points at them)."""
        need = self.t + steps
        if need <= self._cap:
            return
        if self._graph is not None:
            raise RuntimeError("the captured step writes into fixed log buffers: pass the total number of steps to the "
                               "first run() / reserve() call")
        n, dev, dt = self.n_log, self.device, self.dtype
        shapes = dict(x_real=(n, 3), u_real=(n, 2), x_bar=(n, 3), u_bar=(n, 2), b_real=(n,), loss=(n,), Qa=(3,), Ra=(2,), qba=())
        new = {k: torch.zeros((need,) + shp, dtype=dt, device=dev) for k, shp in shapes.items()}
        for k, old in self._logs.items():
            new[k][: old.shape[0]] = old
        self._logs, self._cap = new, need

    reserve = _reserve_logs

    def _log(self, key: str, value: Tensor) -> None:
        self._logs[key].index_copy_(0, self.t_dev, value.unsqueeze(0))

    def _body(self) -> None:
        """One closed-loop step as a fixed sequence of launches on the current stream (capturable)."""
        p, eng, n = self.prob, self.eng, self.n_log
        st = eng.step(self.xh_bar, self.xh, self.Vn_ws, self.Va_ws, self.theta_n, self.theta_a, full=self.adapt_nominal)
        allreduce_shared(st)
        self.last = st
        rn, ra = st.nominal, st.auxiliary
        # pre-step values (reference :619-629 / :1004-1013)
        self._log("x_real", self.xh[:n, :3]); self._log("b_real", self.xh[:n, 3])
        self._log("x_bar", self.xh_bar[:n, :3])
        self._log("u_real", ra.V[:n, 0]); self._log("u_bar", rn.V[:n, 0])
        self._log("loss", st.loss[:n])
        if self.paper:
            self.adapter.update(st.g_aux_sum[:12])        # momentum + projection (reference :978-984)
        else:
            if self.adapt_ancillary:
                self.theta.update(st.g_aux_sum[:12])      # reference :513-514
            if self.adapt_nominal:
                self.theta_bar.update(st.g_nom_sum)       # reference :601-602
        if self.adapt_ancillary:
            self._log("Qa", self.theta_a[0:3]); self._log("Ra", self.theta_a[3:5]); self._log("qba", self.theta_a[8])
        bad = (rn.status != 0).any().to(torch.int64)
        self.failed.add_(bad * (self.failed == 0) * (self.t_dev[0] + 1))
        # true system step with disturbance, barrier of the undisturbed successor, nominal propagation, warm-start
        # shift (reference :607-638 / :990-1023); the plant and the nominal state propagate with the UPDATED DBaS
        # parameters (reference :609, 614)
        w = self.sample_w(self.t) if self.sample_w is not None else None
        ops.plant_step_batched(p.system, p.nominal, p.auxiliary, rn.V, ra.V, self.theta_n, self.theta_a, w, self.xh,
                               self.xh_bar, self.Vn_ws, self.Va_ws, rng=self.rng if w is None else None)
        self.t_dev.add_(1)

    def _body_multi_adapt(self) -> None:
        """adaptation.steps > 1 (reference :407-602 inside `for _ in range(adapt_steps)`): the later adaptation steps
        re-accumulate the gradients with the updated parameters at the same optimum.  Plain launches."""
        p, eng, B, N, dev, dt, n = self.prob, self.eng, self.B, self.N, self.device, self.dtype, self.n_log
        rn, ra = ops.tube_solve_batched(p.system, p.nominal, p.auxiliary, self.xh_bar, self.xh, self.Vn_ws, self.Va_ws,
                                        self.theta_n, self.theta_a, workspace=eng.workspace)
        sa = ops.ddp_sensitivity_batched(p.system, p.auxiliary, ra.X, ra.V, self.theta_a, X_bar=rn.X, X_ref=rn.X, U_ref=rn.V,
                                         want_grads=True, want_ref_grads=True, status=rn.status, workspace=eng.workspace)
        self._log("x_real", self.xh[:n, :3]); self._log("b_real", self.xh[:n, 3]); self._log("x_bar", self.xh_bar[:n, :3])
        self._log("u_real", ra.V[:n, 0]); self._log("u_bar", rn.V[:n, 0]); self._log("loss", sa.loss[:n])
        for s in range(self.adapt_steps):
            if s > 0:
                sa2 = ops.ift_accumulate_batched(p.system, p.auxiliary, ra.X, ra.V, self.theta_a, sa.delta_X, sa.delta_V,
                                                 sa.delta_lambda, X_ref=rn.X, U_ref=rn.V, want_ref_grads=True)
                sa.g_theta, sa.g_xref, sa.g_uref = sa2.g_theta, sa2.g_xref, sa2.g_uref
            if self.adapt_ancillary:
                self.theta.update(sum_over_ranks(ops.reduce_columns(sa.g_theta, eng.workspace)))
            if not self.adapt_nominal:
                continue
            gx = torch.cat([sa.g_xref, torch.zeros(B, N + 1, 1, dtype=dt, device=dev)], dim=2)
            sn = ops.ddp_sensitivity_batched(p.system, p.nominal, rn.X, rn.V, self.theta_n, upper_grad_x=gx,
                                             upper_grad_u=sa.g_uref, want_grads=True, status=rn.status,
                                             workspace=eng.workspace)
            self.theta_bar.update(sum_over_ranks(ops.reduce_columns(sn.g_theta, eng.workspace)))
        if self.adapt_ancillary:
            self._log("Qa", self.theta_a[0:3]); self._log("Ra", self.theta_a[3:5]); self._log("qba", self.theta_a[8])
        bad = (rn.status != 0).any().to(torch.int64)
        self.failed.add_(bad * (self.failed == 0) * (self.t_dev[0] + 1))
        w = self.sample_w(self.t) if self.sample_w is not None else None
        ops.plant_step_batched(p.system, p.nominal, p.auxiliary, rn.V, ra.V, self.theta_n, self.theta_a, w, self.xh,
                               self.xh_bar, self.Vn_ws, self.Va_ws, rng=self.rng if w is None else None)
        self.t_dev.add_(1)

    # ------------------------------------------------------------------
    def _state(self):
        ts = [self.xh, self.xh_bar, self.Vn_ws, self.Va_ws, self.t_dev, self.failed, *self._logs.values()]
        for a in ([self.adapter] if self.paper else [self.theta, self.theta_bar]):
            ts += [a.theta, a.vel] + ([a.raw] if hasattr(a, "raw") else [])
        return ts

    def _capture(self) -> None:
        """Warm-up (allocator pools, workspaces and library streams must exist before capture) on a copy of the state,
        then capture one step; the capture itself executes nothing."""
        saved = [t.clone() for t in self._state()]
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                self._body()
        cur.wait_stream(side)
        for t, s in zip(self._state(), saved):
            t.copy_(s)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._body()

    def step(self) -> None:
        self._reserve_logs(1)
        if self.adapt_steps != 1:
            self._body_multi_adapt()
        elif self.use_graph:
            if self._graph is None:
                self._capture()
            self._graph.replay()
        else:
            self._body()
        self.t += 1

    def check(self) -> None:
        """Raises what the reference would have raised inside its solver (FloatingPointError, core/ddp.py:142)."""
        f = int(self.failed.item())
        if f:
            detail = ""
            if self.last is not None and f == self.t:   # the failing step is the latest one: its location is at hand
                msg = ops.failure_message("iLQR-nominal", self.last.nominal) or \
                    ops.failure_message("iLQR-auxiliary", self.last.auxiliary, shift=4)
                detail = f": {msg}" if msg else ""
            raise FloatingPointError(f"tube MPC: non-finite detected at t={f - 1}{detail}")

    def run(self, H: Optional[int] = None, progress: bool = False) -> Dict[str, Tensor]:
        H = self.H if H is None else int(H)
        self._reserve_logs(H)
        for _ in range(H):
            if progress and (self.t % 25) == 0:
                print(f"[step {self.t}/{H}] running...", flush=True)
            self.step()
        self.check()
        return self.logs()

    def logs(self) -> Dict[str, Tensor]:
        """Device logs of the steps run so far: x_real [t,n,3], u_real [t,n,2], x_bar, u_bar, b_real [t,n], loss [t,n],
        Qa [t,3], Ra [t,2], qba [t]."""
        keys = [k for k in _LOG_KEYS if self.adapt_ancillary or k not in ("Qa", "Ra", "qba")]
        return {k: self._logs[k][: self.t] for k in keys if k in self._logs}
