#!/usr/bin/env python
"""Headline benchmark: batched two-layer Dubins tube-MPC solves (nominal iLQR + auxiliary iLQR + DDP sensitivity +
IFT gradients) per second -- BASELINE.json metric, workload configs[2] (B = 65,536 instances per GPU, T = 50, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode full|paper] [--batch B]

One "step" = one pass of the hot path over one batch of synthetic instances.  `value` is measured with the inputs
resident in HBM; `e2e` goes through the public host-buffer API (pinned host -> device copies of the step's inputs and
a device -> host read of its result inside the timed region).  `--impl reference` times the CPU implementation of the
same path (the C port in oracle/, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "tube_mpc_solves_per_sec"
UNIT = "solves/s"
T_HORIZON = 50
BATCH_PER_GPU = 65536
SEED = 2024


# ---------------------------------------------------------------------------
# workload (cfg#3): x_bar0 ~ U([0,10]^2 x [-pi,pi]) with h > 0.25, x0 = x_bar0 + w, zero warm starts,
# per-instance weights x U(0.5, 2); a pure function of (seed, global instance index), never of the rank
# ---------------------------------------------------------------------------
def make_workload(cfg, mode, lo, hi, total):
    from differentiable_tube_mpc_b200 import TubeProblem
    from differentiable_tube_mpc_b200.workload import make_instances

    prob = TubeProblem.from_config(cfg, mode=mode)
    x_all, xb_all = make_instances(total, SEED, cfg)
    rng = np.random.default_rng(SEED + 1)
    scale_n = rng.uniform(0.5, 2.0, size=(total, 9))
    scale_a = rng.uniform(0.5, 2.0, size=(total, 9))
    if mode == "full":
        import torch

        th_n_raw, th_a_raw = prob.make_thetas("cpu", torch.float64)
        base_n, base_a = th_n_raw.constrained().numpy(), th_a_raw.constrained().numpy()
    else:
        base_n, base_a = np.asarray(prob.theta_nom0), np.asarray(prob.theta_aux0)
    th_n = np.tile(base_n, (total, 1))
    th_a = np.tile(base_a, (total, 1))
    th_n[:, :9] *= scale_n
    th_a[:, :9] *= scale_a
    th_a[:, 8] = np.minimum(th_a[:, 8], 1.0)
    sl = slice(lo, hi)
    return prob, dict(x=x_all[sl], xb=xb_all[sl], th_n=np.ascontiguousarray(th_n[sl]), th_a=np.ascontiguousarray(th_a[sl]))


def algorithmic_flops(T, n_alpha, iters_total, n_solves, n_sens):
    """SURVEY 8d / App. D (dense reference form, 1 flop per add/mul/div/compare) with transcendental calls
    converted at exp = 25, log = 30, sincos = 50 flop-equivalents (one f_hat evaluation = 10 exp + 2 log + 1 sincos)."""
    fhat_tr = 10 * 25 + 2 * 30 + 50
    per_iter = T * (1087 + 188 * n_alpha) + T * (1 + n_alpha) * fhat_tr
    rollout = T * 137 + T * fhat_tr
    sens = T * 1134 + T * fhat_tr
    return per_iter * iters_total + rollout * n_solves + sens * n_sens


def algorithmic_bytes_per_solve(T, full):
    """compulsory fp64 bytes per solve with per-instance parameters and all outputs kept."""
    b = 8 * (4 * T + 31) + 8 * (2 * 4 * (T + 1) + 2 * 2 * T + 2 * 10 * T + (8 * (T + 1) + 2 * T) + 11 + 1) + 16
    if full:
        b += 8 * (8 * (T + 1) + 2 * T) + 96
    return b


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: in-process NVML every 20 ms (an `nvidia-smi -lms`
    subprocess perturbed the measured step time by ~30 %); falls back to nvidia-smi if NVML is unavailable."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.sm, self.reasons, self.power = [], set(), []
        self.smax = None
        self._stop = threading.Event()
        self.thread = None
        self.h = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu_index]) if vis and vis.split(",")[0].isdigit() else self.gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for name, bit in self.REASONS:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def stop(self):
        if self.h is None:
            return self._smi_once()
        self._stop.set()
        self.thread.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None, "source": "nvml"}

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                  "-i", str(self.gpu_index)], capture_output=True, text=True, timeout=10).stdout
            a, b = [float(v) for v in out.strip().split(",")]
            return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "source": "nvidia-smi after the run"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}


def traffic_from_profile():
    """DRAM bytes (read + write) per launch of the largest kernel, from the committed ncu --set full capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01f_traffic.json")))
        return {"kernel": t["kernel"], "bytes_per_launch": t["dram_bytes_read"] + t["dram_bytes_write"], "source": t["source"]}
    except Exception:
        return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ---------------------------------------------------------------------------
# CPU arm: the C port of the reference path (oracle/), all host threads, bounded sample
# ---------------------------------------------------------------------------
def cpu_run(cfg, mode, n_per_step, steps, warmup, threads):
    import dtmpc_oracle as orc

    orc.build()
    env = orc.env_from_cfg(cfg)
    nom, aux = orc.ocps_from_cfg(cfg, mode)
    total = n_per_step
    _, w = make_workload(cfg, mode, 0, total, total)
    T = nom.T
    b0n = np.empty(total)
    b0a = np.empty(total)
    for i in range(total):
        nom.set_theta(w["th_n"][i]); aux.set_theta(w["th_a"][i])
        b0n[i], b0a[i] = orc.b0(env, nom, w["xb"][i]), orc.b0(env, aux, w["x"][i])
    x0n = np.concatenate([w["xb"], b0n[:, None]], 1)
    x0a = np.concatenate([w["x"], b0a[:, None]], 1)
    V0 = np.zeros((total, T, 2))
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        orc.tube_solve_batch(env, nom, aux, x0n, x0a, V0, V0, theta_n=w["th_n"], theta_a=w["th_a"], mode=mode,
                             n_threads=threads)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    dt = float(np.sum(times))
    return n_per_step * steps / dt, dt / steps * 1e3


_REAL_STDOUT = None


def guard_stdout():
    """The driver parses stdout as ONE JSON line: libraries that write to fd 1 (NCCL prints its version there) are
    redirected to stderr; emit() writes the result line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", type=str, default="full", choices=["full", "paper"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="instances per GPU")
    ap.add_argument("--horizon", type=int, default=T_HORIZON)
    ap.add_argument("--dtype", type=str, default="f64", choices=["f64", "f32"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sequential", action="store_true", help="one step in flight (no overlap of consecutive steps)")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--warmup-seconds", type=float, default=2.0, help="minimum warm-up wall time (clock ramp)")
    args = ap.parse_args()
    guard_stdout()

    from differentiable_tube_mpc_b200 import default_dubins_config

    cfg = default_dubins_config()
    cfg["system"]["horizon_N"] = args.horizon
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    workload = (f"two-layer tube MPC ({args.mode} mode), Dubins+DBaS, T={args.horizon}, B={args.batch}/GPU, "
                f"per-instance theta, IFT gradients")
    config = {"workload": workload, "mode": args.mode, "horizon": args.horizon, "batch_per_gpu": args.batch,
              "global_batch": args.batch * max(world, 1), "parallelism": f"dp{max(world, 1)}",
              "l2": "per-step working set (inputs+outputs+scratch > 1 GB) exceeds the 126 MB L2; no explicit flush"}

    # ---------------- reference arm (CPU) ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        n = args.cpu_sample or cores * (1280 if args.mode == "full" else 2048)  # ~10 s of CPU work per step
        val, ms = cpu_run(cfg, args.mode, n, max(args.steps, 1), max(args.warmup, 0), cores)
        line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{n} instances of the same seeded workload per step, {cores} threads; "
                                           "C port of the reference algorithm (oracle/dtmpc_oracle.c); the reference's own "
                                           "Python/PyTorch path measured 0.15-0.28 solves/s per core (BASELINE.md)"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    # ---------------- B200 arm ----------------
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    from differentiable_tube_mpc_b200 import _lib, ops
    from differentiable_tube_mpc_b200.engine import BatchedTubeMPC, DoubleBufferedSteps, allreduce_shared, shard_bounds

    _lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    total = args.batch * world
    lo, hi = shard_bounds(total, world, rank)
    prob, w = make_workload(cfg, args.mode, lo, hi, total)
    B, T = hi - lo, args.horizon
    full = args.mode == "full"
    eng = BatchedTubeMPC(prob, device=dev, dtype=dtype)

    # host (pinned) copies of the step inputs
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).pin_memory()
    h_th_n, h_th_a = pin(w["th_n"]), pin(w["th_a"])
    d_th_n, d_th_a = h_th_n.to(dev), h_th_a.to(dev)
    b0n = eng.init_barrier(torch.as_tensor(w["xb"], dtype=dtype, device=dev), d_th_n, nominal=True)
    b0a = eng.init_barrier(torch.as_tensor(w["x"], dtype=dtype, device=dev), d_th_a, nominal=False)
    d_x0n = torch.cat([torch.as_tensor(w["xb"], dtype=dtype, device=dev), b0n.view(-1, 1)], dim=1).contiguous()
    d_x0a = torch.cat([torch.as_tensor(w["x"], dtype=dtype, device=dev), b0a.view(-1, 1)], dim=1).contiguous()
    d_V0 = torch.zeros(B, T, 2, dtype=dtype, device=dev)
    h_x0n, h_x0a, h_V0n, h_V0a = d_x0n.cpu().pin_memory(), d_x0a.cpu().pin_memory(), d_V0.cpu().pin_memory(), \
        d_V0.cpu().pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in (h_x0n, h_x0a, h_V0n, h_V0a, h_th_n, h_th_a))
    # result of an MPC step read back by the caller: applied controls, loss, status and the summed gradients
    h_u0 = torch.empty(B, 2, dtype=dtype).pin_memory()
    h_ub0 = torch.empty(B, 2, dtype=dtype).pin_memory()
    h_loss = torch.empty(B, dtype=dtype).pin_memory()
    h_status = torch.empty(B, dtype=torch.int32).pin_memory()
    h_g = torch.empty(13 + (12 if full else 0), dtype=dtype).pin_memory()
    d2h = sum(t.numel() * t.element_size() for t in (h_u0, h_ub0, h_loss, h_status, h_g))

    # The step (~950 kernel launches at the default size) is captured once into a CUDA graph and replayed: one launch
    # per step, no per-launch host cost inside the timed region.  --no-graph launches the kernels one by one.
    graphed = None if args.no_graph else eng.graphed(d_x0n, d_x0a, d_V0, d_V0, d_th_n, d_th_a, full=full)

    def plain_step():
        return allreduce_shared(eng.step(d_x0n, d_x0a, d_V0, d_V0, d_th_n, d_th_a, full=full))

    def sequential_step():   # one step at a time: the next replay starts when this one has finished
        if graphed is None:
            return plain_step()
        return allreduce_shared(graphed())

    # end to end: every step copies its inputs from pinned host memory and reads its results back.  Two graphed steps
    # with separate static inputs alternate, so the copy of step i+1 (copy stream) overlaps the kernels of step i.
    h_inputs = dict(x0_nom=h_x0n, x0_aux=h_x0a, Vn_init=h_V0n, Va_init=h_V0a, theta_nom=h_th_n, theta_aux=h_th_a)
    # The bench steps are independent batches, so two consecutive steps may be in flight (separate scratch, separate
    # compute streams): the low-occupancy head and tail of one step fill with the other's line search.
    # --sequential keeps one step in flight; the JSON line always carries the sequential figure too.
    overlap = not (args.sequential or args.no_graph)
    dbuf = None if args.no_graph else DoubleBufferedSteps(eng, d_x0n, d_x0a, d_V0, d_V0, d_th_n, d_th_a, full=full,
                                                          overlap_steps=overlap, buffers=int(os.environ.get("DTMPC_BENCH_BUFFERS", "2")))

    def device_steps(n):     # inputs resident: the static tensors of the two buffers already hold them
        last = None
        if dbuf is None:
            for _ in range(n):
                last = plain_step()
            return last
        for _ in range(n):
            dbuf.prefetch()
            last = allreduce_shared(dbuf.run())
        return last

    def read_back(st):
        h_u0.copy_(st.auxiliary.V[:, 0], non_blocking=True)
        h_ub0.copy_(st.nominal.V[:, 0], non_blocking=True)
        h_loss.copy_(st.loss, non_blocking=True)
        h_status.copy_(st.status, non_blocking=True)
        h_g.copy_(torch.cat([st.g_aux_sum] + ([st.g_nom_sum] if full else [])), non_blocking=True)

    def host_steps(n):
        last = None
        if dbuf is None:
            for _ in range(n):
                dv = {k: v.to(dev, non_blocking=True) for k, v in h_inputs.items()}
                last = allreduce_shared(eng.step(dv["x0_nom"], dv["x0_aux"], dv["Vn_init"], dv["Va_init"],
                                                 dv["theta_nom"], dv["theta_aux"], full=full))
                read_back(last)
            return last
        dbuf.prefetch(**h_inputs)
        for i in range(n):
            if i + 1 < n:
                dbuf.prefetch(**h_inputs)
            last = allreduce_shared(dbuf.run())
            read_back(last)
        return last

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        if fn is host_steps or fn is device_steps:
            last = fn(steps)
        else:
            for _ in range(steps):
                last = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    # warm-up: at least W (>= 3) steps AND ~2 s of load -- the first launches pay lazy module loading and the SM
    # clocks need a few hundred ms to ramp from idle (first two steps measured 2-9x slower than steady state)
    # Every step ends in a collective when world > 1, so all ranks must run the SAME number of warm-up steps: the count
    # that fills the warm-up time is derived from all-reduced timings, never from a rank's own clock.
    n_warm, t_warm = 0, time.perf_counter()
    for _ in range(max(args.warmup, 3)):
        device_steps(1)
        torch.cuda.synchronize()
        n_warm += 1
    t1 = time.perf_counter()
    device_steps(2)
    torch.cuda.synchronize()
    n_warm += 2
    est = torch.tensor([time.perf_counter() - t_warm, (time.perf_counter() - t1) / 2.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MIN)
    spent, per_step = float(est[0]), max(float(est[1]), 1e-4)
    n_extra = int(max(0.0, args.warmup_seconds - spent) / per_step)
    for _ in range(n_extra):
        device_steps(1)
        n_warm += 1
    torch.cuda.synchronize()
    # per-kernel-group durations (CUDA events around each library call, on the launching stream) and the number of
    # kernels one step launches: measured on the plain launch path, which is what the graph replays.  The plain path
    # pays ~190 host launches per chunk, so its first steps and any host hiccup inflate a section: median of 9.
    for _ in range(3):
        plain_step()
    torch.cuda.synchronize()
    eng.timers = {}
    l0 = eng.launches
    n_plain = 9
    for _ in range(n_plain):
        plain_step()
        torch.cuda.synchronize()
    timers, eng.timers = eng.timers, None
    launches_per_step = (eng.launches - l0) // n_plain
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, last = timed(device_steps, args.steps)
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    host_steps(2)
    ms_e2e, last_h = timed(host_steps, args.steps)
    sequential_step()
    ms_seq, _ = timed(sequential_step, args.steps)
    torch.cuda.synchronize()

    n_bad = int((last.status != 0).sum())
    iters_total = int(last.nominal.iters.sum()) + int(last.auxiliary.iters.sum())
    kern_ms = {k: float(np.median([a.elapsed_time(b) for a, b in v])) for k, v in timers.items()}
    value = total * args.steps / (ms_total * 1e-3)
    e2e_val = total * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        n_alpha = prob.nominal.ilqr.line_search_alphas
        # dominant work item: the iLQR solve = launch sequence init + max_iter x (ric, fwd, commit) for the
        # nominal and the auxiliary OCP; fwd_kernel (line search) is its largest kernel (ncu share in profiles/)
        # one library call per step (dtmpc_tube_step): both solves and the sensitivity sweeps of every chunk
        flops = algorithmic_flops(T, len(n_alpha), iters_total, 2 * B, B * (2 if full else 1))
        peak64 = ops.measure_fp64_peak_tflops() if dtype == torch.float64 else ops.measure_fp32_peak_tflops()
        # over the timed region (CUDA events around the K replays): the step is the dtmpc_tube_step launch sequence
        # plus two reduction kernels (0.3 % of it); kernel_ms holds the sections of a plain sequential step
        ach = flops / (ms_total / args.steps * 1e-3) / 1e12
        peaks = measured_peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        step_ms = ms_total / args.steps
        bytes_solve = algorithmic_bytes_per_solve(T, full) * (1 if dtype == torch.float64 else 0.5)
        hbm_ach = bytes_solve * B / (step_ms * 1e-3) / 1e9
        roofline = {"bound": "fp64" if dtype == torch.float64 else "fp32", "kernel": "dtmpc_tube_step launch sequence (per OCP: init + max_iter x [ric, fwd, commit]; then sens; largest: fwd_kernel)",
                    "achieved": ach, "peak": peak64, "unit": "TFLOP/s", "frac": ach / peak64 if peak64 > 0 else None,
                    "peak_source": "DFMA loop measured live on this GPU (MEASURED_PEAKS.json has no FP64 figure)",
                    "flop_model": "SURVEY 8d dense reference form x actual iterations; exp=25 log=30 sincos=50 flop-eq",
                    "kernel_ms": kern_ms, "iters_per_solve": iters_total / B, "traffic": traffic_from_profile(),
                    "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                            "bytes_per_solve": bytes_solve}}
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            n = args.cpu_sample or cores * (1280 if full else 2048)  # ~10-20 s of CPU work
            cval, _ = cpu_run(cfg, args.mode, n, 1, 0, cores)
            cpu_baseline = {"value": cval, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{n} instances of the same seeded workload, {cores} threads, C port of the "
                                      "reference algorithm (oracle/dtmpc_oracle.c)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": n_warm, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": config,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "launch_mode": "plain" if graphed is None else "cuda_graph_replay",
                "steps_in_flight": (dbuf.nb if overlap else 1),
                "sequential": {"value": total * args.steps / (ms_seq * 1e-3), "unit": UNIT,
                               "ms_per_step": ms_seq / args.steps,
                               "note": "same K steps with one step in flight (each replay starts after the previous one)"},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
                "nonfinite_instances": n_bad}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
