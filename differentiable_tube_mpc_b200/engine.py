"""Batched two-layer tube MPC: one call = what one closed-loop step of the reference computes per instance.

  nominal ilqr_solve -> auxiliary ilqr_solve -> upper loss -> ddp_sensitivity(aux) -> theta gradient
  [-> ddp_sensitivity(nominal) driven by dL/dx_ref, dL/du_ref -> theta_bar gradient]   (full mode)

reference: core/tube_mpc.py:844-976 (paper mode) and :305-602 (full mode).  The batch is sharded across
ranks (one process per GPU); the only collective is the all-reduce of the batch-summed gradient when the
parameters are shared (Algorithm 2), see `allreduce_shared`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch
from torch import Tensor

from . import ops
from .spec import TubeProblem


@dataclass
class TubeStep:
    nominal: ops.SolveResult
    auxiliary: ops.SolveResult
    sens_aux: ops.SensitivityResult
    sens_nom: Optional[ops.SensitivityResult]
    loss: Tensor                 # [B] upper loss per instance (core/tube_mpc.py:916-919)
    g_aux_sum: Tensor            # [13] = batch sums of d/d(constrained theta) [12] and of the loss [1]
    g_nom_sum: Optional[Tensor]  # [12] batch sums of d/d(constrained theta_bar) (full mode)

    @property
    def status(self) -> Tensor:
        return self.nominal.status


class BatchedTubeMPC:
    """Solver for B independent tube-MPC instances on one GPU."""

    def __init__(self, problem: TubeProblem, device="cuda", dtype=torch.float64, gains: bool = False,
                 trace: bool = False):
        self.problem = problem
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("BatchedTubeMPC runs on a CUDA device only (no CPU fallback)")
        self.dtype = dtype
        self.gains = gains
        self.trace = trace
        self.workspace = ops.Workspace(self.device)
        self.timers: Optional[Dict[str, list]] = None   # filled with CUDA events when enabled (bench.py)
        self.unfused = False                            # True: separate library calls per stage (per-stage timers)

    @property
    def launches(self) -> int:
        """CUDA kernels the library has enqueued in this process so far (counted inside the C-ABI library)."""
        return ops.kernel_launches()

    def graphed(self, x0_nom: Tensor, x0_aux: Tensor, Vn_init: Tensor, Va_init: Tensor, theta_nom: Tensor,
                theta_aux: Tensor, full: Optional[bool] = None) -> "GraphedStep":
        """step() captured into a CUDA graph for these shapes (see GraphedStep)."""
        return GraphedStep(self, x0_nom, x0_aux, Vn_init, Va_init, theta_nom, theta_aux, full=full)

    # -- initial barrier state b0 = B(h(x0) - tight)  (core/barrier.py:111-120) -------------
    def init_barrier(self, x: Tensor, theta: Tensor, nominal: bool) -> Tensor:
        ocp = self.problem.nominal if nominal else self.problem.auxiliary
        B = x.shape[0]
        xh = torch.cat([x, torch.zeros(B, 1, dtype=x.dtype, device=x.device)], dim=1)
        u = torch.zeros(B, 2, dtype=x.dtype, device=x.device)
        return ops.primitives_batched(self.problem.system, ocp, xh, u, theta)["b0"]

    def _timed(self, name):
        eng = self

        class _T:
            def __enter__(self_inner):
                if eng.timers is not None:
                    self_inner.e0 = torch.cuda.Event(enable_timing=True)
                    self_inner.e1 = torch.cuda.Event(enable_timing=True)
                    self_inner.e0.record()
                return self_inner

            def __exit__(self_inner, *exc):
                if eng.timers is not None:
                    self_inner.e1.record()
                    eng.timers.setdefault(name, []).append((self_inner.e0, self_inner.e1))
                return False

        return _T()

    def _step_unfused(self, x0_nom, x0_aux, Vn_init, Va_init, theta_nom, theta_aux, target, full):
        """The same step as separate library calls (solve, sensitivity, sensitivity): per-section timing, custom target."""
        p = self.problem
        with self._timed("tube_solve"):
            rn, ra = ops.tube_solve_batched(p.system, p.nominal, p.auxiliary, x0_nom, x0_aux, Vn_init, Va_init,
                                            theta_nom, theta_aux, target=target, gains=self.gains, trace=self.trace,
                                            workspace=self.workspace)
        with self._timed("sens_aux"):
            sa = ops.ddp_sensitivity_batched(p.system, p.auxiliary, ra.X, ra.V, theta_aux, X_bar=rn.X, X_ref=rn.X,
                                             U_ref=rn.V, want_grads=True, want_ref_grads=full, status=rn.status,
                                             workspace=self.workspace)
        sn = None
        if full:
            B, T = rn.X.shape[0], rn.V.shape[1]
            # upper gradients of the nominal problem: [dL/dx_ref, 0] and dL/du_ref (core/tube_mpc.py:520-535)
            gx = torch.cat([sa.g_xref, torch.zeros(B, T + 1, 1, dtype=self.dtype, device=self.device)], dim=2)
            with self._timed("sens_nom"):
                sn = ops.ddp_sensitivity_batched(p.system, p.nominal, rn.X, rn.V, theta_nom, upper_grad_x=gx,
                                                 upper_grad_u=sa.g_uref, target=target, want_grads=True,
                                                 status=rn.status, workspace=self.workspace)
        return rn, ra, sa, sn

    def step(self, x0_nom: Tensor, x0_aux: Tensor, Vn_init: Tensor, Va_init: Tensor, theta_nom: Tensor,
             theta_aux: Tensor, target: Optional[Tensor] = None, full: Optional[bool] = None) -> TubeStep:
        """All tensors on self.device: x0_* [B,4], V*_init [B,T,2], theta_* [12] or [B,12] (constrained values)."""
        p = self.problem
        full = (p.mode == "full") if full is None else full
        if target is None and not self.unfused:
            # one library call: every chunk of the batch runs solves and sensitivity sweeps on its own stream
            with self._timed("tube_step"):
                rn, ra, sa, sn = ops.tube_step_batched(p.system, p.nominal, p.auxiliary, x0_nom, x0_aux, Vn_init, Va_init,
                                                       theta_nom, theta_aux, full=full, gains=self.gains,
                                                       trace=self.trace, workspace=self.workspace)
        else:
            rn, ra, sa, sn = self._step_unfused(x0_nom, x0_aux, Vn_init, Va_init, theta_nom, theta_aux, target, full)
        g_nom_sum = None
        with self._timed("reduce"):
            g_aux_sum = ops.reduce_columns(torch.cat([sa.g_theta, sa.loss.view(-1, 1)], dim=1), self.workspace)
            if full:
                g_nom_sum = ops.reduce_columns(sn.g_theta, self.workspace)
        return TubeStep(nominal=rn, auxiliary=ra, sens_aux=sa, sens_nom=sn, loss=sa.loss, g_aux_sum=g_aux_sum,
                        g_nom_sum=g_nom_sum)


class GraphedStep:
    """`BatchedTubeMPC.step` on fixed-shape inputs captured once into a CUDA graph and replayed with one launch.

    A step is ~190 kernel launches per chunk of the batch; replaying them as a graph removes the per-launch host cost
    and makes the step time independent of host jitter.  Inputs are copied into the graph's static tensors
    (`graphed(x0_nom=..., ...)`, any subset), results are the static tensors of `self.out`, overwritten by each replay.
    The library's internal chunk streams fork from and join the capturing stream, so they are part of the graph.
    """

    NAMES = ("x0_nom", "x0_aux", "Vn_init", "Va_init", "theta_nom", "theta_aux")

    def __init__(self, eng: "BatchedTubeMPC", x0_nom, x0_aux, Vn_init, Va_init, theta_nom, theta_aux, full=None):
        self.static = {k: v.clone() for k, v in zip(self.NAMES, (x0_nom, x0_aux, Vn_init, Va_init, theta_nom, theta_aux))}
        args = [self.static[k] for k in self.NAMES]
        saved, eng.timers = eng.timers, None
        try:
            side = torch.cuda.Stream(device=eng.device)
            side.wait_stream(torch.cuda.current_stream(eng.device))
            with torch.cuda.stream(side):   # warm-up: workspaces, library streams and allocator pools exist before capture
                for _ in range(2):
                    eng.step(*args, full=full)
            torch.cuda.current_stream(eng.device).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = eng.step(*args, full=full)
        finally:
            eng.timers = saved

    def __call__(self, **inputs) -> TubeStep:
        for k, v in inputs.items():
            self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.out


class DoubleBufferedSteps:
    """Two graphed steps with separate static inputs, used alternately: `prefetch(**host_inputs)` copies the NEXT
    step's inputs from (pinned) host memory on a copy stream while the current step's kernels run; `run()` waits for
    its inputs, replays the graph and returns the static results of that buffer (valid until the second next run).

        db = DoubleBufferedSteps(eng, x0_nom, x0_aux, Vn, Va, th_n, th_a)
        db.prefetch(**inputs[0])
        for i in range(n):
            if i + 1 < n: db.prefetch(**inputs[i + 1])
            out = db.run()
    """

    def __init__(self, eng: "BatchedTubeMPC", *args, full=None, overlap_steps: bool = False, buffers: int = 2):
        """overlap_steps = True gives the second buffer its own scratch and each buffer its own compute stream, so two
        consecutive (independent) steps may be in flight at once: the low-occupancy head (initial rollouts) and tail
        (sensitivity sweeps) of one step fill with the other's line search.  Only for streams of independent batches;
        a closed loop needs step i before it can start step i+1."""
        self.device = eng.device
        self.overlap = bool(overlap_steps)
        self.nb = nb = max(2, int(buffers))            # buffers = steps that may be in flight when overlap_steps
        engs = [eng] + [BatchedTubeMPC(eng.problem, device=eng.device, dtype=eng.dtype, gains=eng.gains, trace=eng.trace)
                        if self.overlap else eng for _ in range(nb - 1)]
        self.steps = [GraphedStep(e, *args, full=full) for e in engs]
        self.copy_stream = torch.cuda.Stream(device=eng.device)
        self.compute = [torch.cuda.Stream(device=eng.device) for _ in range(nb)] if self.overlap else None
        self.ready = [torch.cuda.Event() for _ in range(nb)]
        self.done = [None] * nb
        self.consumed = [None] * nb
        self.n_prefetched = 0
        self.n_run = 0

    def prefetch(self, **host_inputs) -> None:
        slot = self.n_prefetched % self.nb
        assert self.n_prefetched - self.n_run < self.nb, "every buffer already holds a step that has not run"
        with torch.cuda.stream(self.copy_stream):
            if self.done[slot] is not None:            # the graph that last read these static inputs has finished
                self.copy_stream.wait_event(self.done[slot])
            else:                                      # first use: the static tensors were made on the caller's stream
                self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            for k, v in host_inputs.items():
                self.steps[slot].static[k].copy_(v, non_blocking=True)
            self.ready[slot].record(self.copy_stream)
        self.n_prefetched += 1

    def run(self) -> TubeStep:
        """Replays the next step; the results are ordered on the CALLER's stream (read them there)."""
        slot = self.n_run % self.nb
        assert self.n_prefetched > self.n_run, "run() without a prefetch()"
        cur = torch.cuda.current_stream(self.device)
        if self.overlap:
            # the caller has by now enqueued its reads of the PREVIOUS step's results on its stream: mark them, so that
            # the graph that overwrites that buffer two steps later waits for them (and for nothing else on `cur`)
            if self.n_run > 0:
                prev = (slot - 1) % self.nb
                self.consumed[prev] = torch.cuda.Event()
                self.consumed[prev].record(cur)
            cs = self.compute[slot]
            if self.consumed[slot] is not None:
                cs.wait_event(self.consumed[slot])
            else:
                cs.wait_stream(cur)                    # first use of this buffer
            cs.wait_event(self.ready[slot])
            with torch.cuda.stream(cs):
                out = self.steps[slot]()
            self.done[slot] = torch.cuda.Event()
            self.done[slot].record(cs)
            cur.wait_event(self.done[slot])
        else:
            cur.wait_event(self.ready[slot])
            out = self.steps[slot]()
            self.done[slot] = torch.cuda.Event()
            self.done[slot].record(cur)
        self.n_run += 1
        return out


def sum_over_ranks(vec: Tensor, group=None) -> Tensor:
    """Sum of a small vector over the data-parallel ranks with a FIXED order of additions: the ranks' vectors are
    all-gathered (one collective, < 2 KB) and added in rank order on every rank.  The result is bitwise the same on
    all ranks and does not depend on the algorithm NCCL would pick for an all-reduce (ring, tree, NVLS), so runs are
    reproducible across topologies.  Returns `vec` itself without an initialised process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return vec
    world = dist.get_world_size(group)
    src = vec.contiguous()
    if src.is_cuda and dist.get_backend(group) == "gloo":   # test set-up (several ranks on one GPU): stage through the host
        src = src.cpu()
    parts = [torch.empty_like(src) for _ in range(world)]
    dist.all_gather(parts, src, group=group)
    total = parts[0].clone()
    for r in range(1, world):
        total += parts[r]
    return total.to(vec.device)


def allreduce_shared(step: TubeStep, group=None) -> TubeStep:
    """Sum the batch-summed gradients (and the loss) over the data-parallel ranks: the one collective of the path
    (Algorithm 2 with shared theta; < 200 bytes, latency bound).  The sums are overwritten IN PLACE: a graphed step
    returns the same static tensors on every replay, so rebinding the fields would detach the caller from them.
    No-op without an initialised process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return step
    n_aux = step.g_aux_sum.numel()
    parts = [step.g_aux_sum] + ([step.g_nom_sum] if step.g_nom_sum is not None else [])
    total = sum_over_ranks(torch.cat(parts), group=group)
    step.g_aux_sum.copy_(total[:n_aux])
    if step.g_nom_sum is not None:
        step.g_nom_sum.copy_(total[n_aux:])
    return step


def shard_bounds(total: int, world_size: int, rank: int):
    """Contiguous slice [lo, hi) of a batch of `total` instances owned by `rank`."""
    base, rem = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
