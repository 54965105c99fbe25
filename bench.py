#!/usr/bin/env python
"""Headline benchmark: batched two-layer Dubins tube-MPC solves (nominal iLQR + auxiliary iLQR + DDP sensitivity +
IFT gradients) per second -- BASELINE.json metric, workload configs[2] (B = 65,536 instances per GPU, T = 50, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode full|paper] [--batch B]
                    [--horizon T] [--dtype f64|f32] [--workload tube|nominal|closedloop] [--no-extras]

One "step" = one pass of the hot path over one batch of synthetic instances.  `value` is measured with the inputs
resident in HBM; `e2e` goes through the public host-buffer API (pinned host -> device copies of the step's inputs and
a device -> host read of its result inside the timed region).  `--impl reference` times the CPU implementation of the
same path (the C port in oracle/, all host threads) on a bounded sample of the same workload.

The other BASELINE configs are measured by the same program: `--workload nominal` (configs[1]: nominal-only iLQR with
run_nominal.py semantics, B = 4,096), `--horizon / --batch / --dtype` (configs[3] sweep points), `--workload closedloop`
(configs[0] at batch scale and configs[4]: shared-theta Algorithm 2, gradient summed over the ranks, update applied).
The default run appends short measurements of them to its JSON line under `extras` (`--no-extras` skips that).
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
NOMINAL_BATCH = 4096
SEED = 2024
COUNTERS = os.path.join(ROOT, "profiles", "r02_step_counters.json")
COUNTERS_F32 = os.path.join(ROOT, "profiles", "r02_step_counters_f32.json")


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


def make_nominal_workload(cfg, lo, hi, total):
    """SURVEY 8d cfg#2: nominal OCP alone with run_nominal.py semantics (wrapped heading error :297-324, reg = ilqr_reg,
    tol 1e-3, YAML step sizes, warm start v = v_max, omega = 0 :336-338); x_bar0 ~ U([0,10]^2 x [-pi,pi]), h > 0.25,
    seed 1234."""
    from differentiable_tube_mpc_b200 import TubeProblem
    from differentiable_tube_mpc_b200.run_nominal import _nominal_ocp
    from differentiable_tube_mpc_b200.workload import make_instances

    prob = TubeProblem.from_config(cfg, mode="paper")
    ocp = _nominal_ocp(cfg, prob)
    _, xb = make_instances(total, 1234, cfg)
    T = int(cfg["system"]["horizon_N"])
    V0 = np.zeros((hi - lo, T, 2))
    V0[:, :, 0] = float(cfg["system"]["control_bounds"]["v_max"])
    return prob, ocp, np.ascontiguousarray(xb[lo:hi]), V0


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


def load_counters(f32=False):
    try:
        return json.load(open(COUNTERS_F32 if f32 else COUNTERS))
    except Exception:
        return None


def executed_work(counters, T, tiles, ocps, n_sens):
    """EXECUTED work of one step from the committed ncu opcode counters (profiles/r02_step_counters.json: fp64-pipe warp
    instructions, issued warp instructions and DRAM bytes per tile-step of every kernel, measured on this code) times
    the launch / iteration counts the step actually ran.  ocps = [(n_eval, sum of iterations over the instances / 32)]:
    a launch only works on the tiles that still hold running instances (compacted slots)."""
    pu = counters["per_tile_step"]
    out = {}
    for key in ("fp64_inst", "inst", "dram_bytes"):
        c = lambda f: pu.get(f, {}).get(key, 0.0)
        v = 0.0
        for (n_eval, tile_iters), o in zip(ocps, ("n", "a")):   # nominal, auxiliary (the latter also streams its reference)
            v += tiles * T * c(f"init_{o}") + tile_iters * T * (c(f"ric_{o}") + n_eval * c(f"fwdc_{o}"))
        v += tiles * T * (c("sens_a") + (c("sens_n") if n_sens > 1 else 0.0))
        out[key] = v
    return out


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


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and the pinned host buffers it is about to allocate) to the CPUs of the NUMA node the GPU
    hangs off: at N = 8 every rank otherwise stages its 120 MB / step of host inputs through one socket
    (round 1: e2e 6 % below 8 x the single-GPU figure).  Best effort; returns what was done."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis and vis.split(",")[0].isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinityWithinScope(h, n_words, pynvml.NVML_AFFINITY_SCOPE_NODE)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return {"cpus": f"{cpus[0]}-{cpus[-1]}", "n": len(cpus)}
        return {"cpus": "all", "n": len(os.sched_getaffinity(0))}
    except Exception as e:   # noqa: BLE001
        return {"cpus": "unchanged", "error": type(e).__name__}


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


def cpu_run_nominal(cfg, n, threads):
    """The C port on the nominal-only problem of configs[1] (one ilqr_solve per instance, `threads` host threads)."""
    from concurrent.futures import ThreadPoolExecutor

    import dtmpc_oracle as orc

    orc.build()
    env = orc.env_from_cfg(cfg)
    nom, _ = orc.ocps_from_cfg(cfg, "paper")
    nom.angle_wrap = 1
    nom.reg = float(cfg["system"].get("ilqr_reg", 1e-6))
    _, _, xb, V0 = make_nominal_workload(cfg, 0, n, n)
    x0 = np.concatenate([xb, np.array([orc.b0(env, nom, v) for v in xb])[:, None]], 1)
    bounds = np.linspace(0, n, threads + 1).astype(int)

    def work(j):
        o = nom.copy()
        for i in range(bounds[j], bounds[j + 1]):
            orc.ilqr_solve(env, o, x0[i], V0[i])

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))
    return n / (time.perf_counter() - t0)


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


# ---------------------------------------------------------------------------
# secondary workloads (each returns a small dict; device-timed with CUDA events, max over ranks by the caller)
# ---------------------------------------------------------------------------
def _timed_ms(fn, n, dev, world):
    import torch
    import torch.distributed as dist

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def bench_nominal(cfg, dev, dtype, B, steps, warmup, world=1, rank=0):
    """configs[1]: B nominal iLQR solves (run_nominal.py problem) per step, graph replay, inputs resident."""
    import torch
    from differentiable_tube_mpc_b200 import ops

    lo, hi = rank * B, (rank + 1) * B
    prob, ocp, xb, V0 = make_nominal_workload(cfg, lo, hi, B * world)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=dev)
    theta = t(prob.theta_nom0)
    x = t(xb)
    xh0 = torch.cat([x, torch.zeros(B, 1, dtype=dtype, device=dev)], 1)
    b = ops.primitives_batched(prob.system, ocp, xh0, torch.zeros(B, 2, dtype=dtype, device=dev), theta)["b0"]
    xh, V = torch.cat([x, b.view(B, 1)], 1).contiguous(), t(V0)
    ws = ops.Workspace(dev)
    run = lambda: ops.ilqr_solve_batched(prob.system, ocp, xh, V, theta, workspace=ws)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            run()
    torch.cuda.current_stream(dev).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        res = run()
    for _ in range(max(warmup, 3)):
        g.replay()
    ms = _timed_ms(g.replay, steps, dev, world)
    it = res.iters.float()
    return {"workload": f"nominal-only iLQR, run_nominal.py semantics (wrapped heading, reg=ilqr_reg), T={ocp.ilqr.horizon}, "
                        f"B={B}/GPU, shared theta", "metric": "nominal_ilqr_solves_per_sec",
            "value": B * world * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
            "iters_mean": float(it.mean()), "iters_max": int(it.max()), "nonfinite": int((res.status != 0).sum()),
            "dtype": "f64" if dtype == torch.float64 else "f32"}, res


def bench_closed_loop(cfg_in, dev, B, H, mode, world=1, rank=0, dtype_f64=True):
    """configs[0] at batch scale + configs[4]: B tubes per rank in closed loop, SHARED theta, Algorithm-2 update applied
    every step from the gradient summed over the batch and the ranks; one captured step replayed H times (one rank) or
    plain launches around the collective (several ranks)."""
    import copy

    import torch
    from differentiable_tube_mpc_b200.closed_loop import ClosedLoop
    from differentiable_tube_mpc_b200.workload import make_instances

    cfg = copy.deepcopy(cfg_in)
    cfg["use_float64"] = bool(dtype_f64)
    cfg["adaptation"]["adapt_nominal"] = mode == "full"
    cfg["adaptation"]["lr_eta"] = (1e-3 if mode == "full" else 0.05) / (B * world)   # the summed gradient grows with B
    total = B * world
    x_all, xb_all = make_instances(total, SEED, cfg)
    dt = torch.float64 if dtype_f64 else torch.float32
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    lo, hi = rank * B, (rank + 1) * B
    loop = ClosedLoop(cfg, device=dev, batch=B, n_log=1, x0=t(x_all[lo:hi]), x_bar0=t(xb_all[lo:hi]), first_instance=lo)
    loop.reserve(H + 3)
    th0 = loop.theta_a.clone()
    for _ in range(3):
        loop.step()
    ms = _timed_ms(loop.step, H, dev, world)
    failed = int(loop.failed.item())
    return {"workload": f"closed loop ({mode} mode), shared theta, Algorithm-2 update every step (gradient summed over batch"
                        f"{' and ranks' if world > 1 else ''}), in-kernel disturbance, T={loop.N}, B={B}/GPU, H={H} timed steps",
            "metric": "closed_loop_tube_steps_per_sec", "value": total * H / (ms * 1e-3), "unit": "tube-steps/s",
            "ms_per_step": ms / H, "steps": H, "launch": "cuda_graph_replay" if loop._graph is not None else "plain",
            "theta_moved": float((loop.theta_a - th0).abs().max()), "failed_at": failed - 1 if failed else None,
            "loss_mean_last": float(loop.logs()["loss"][-1].mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", type=str, default="full", choices=["full", "paper"])
    ap.add_argument("--workload", type=str, default="tube", choices=["tube", "nominal", "closedloop"])
    ap.add_argument("--adapt", action="store_true", help="alias of --workload closedloop (shared theta, update applied)")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (default 65536; nominal workload 4096)")
    ap.add_argument("--horizon", type=int, default=T_HORIZON)
    ap.add_argument("--dtype", type=str, default="f64", choices=["f64", "f32"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short measurements of the other BASELINE configs")
    ap.add_argument("--sequential", action="store_true", help="one step in flight (no overlap of consecutive steps)")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--preheat-seconds", "--warmup-seconds", dest="preheat_seconds", type=float, default=2.0,
                    help="load applied BEFORE the W warm-up steps so that the SM clocks have ramped (not part of W)")
    args = ap.parse_args()
    guard_stdout()
    if args.adapt:
        args.workload = "closedloop"
    if args.batch <= 0:
        args.batch = NOMINAL_BATCH if args.workload == "nominal" else BATCH_PER_GPU

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
    W = max(args.warmup, 3)

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
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else {"cpus": "unchanged (single rank)"}
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
    full = args.mode == "full"
    peaks = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    def finish():
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    # ---- secondary workloads as the main line ----
    if args.workload == "nominal":
        l0 = ops.kernel_launches()
        r, _ = bench_nominal(cfg, dev, dtype, args.batch, args.steps, W, world, rank)
        if rank == 0:
            emit({"metric": r["metric"], "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
                  "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                  "dtype": args.dtype, "data": "synthetic", "config": dict(config, workload=r["workload"]),
                  "gpu_launches": (1 + 3 * 10) * args.steps, "launches_enqueued_at_capture": ops.kernel_launches() - l0,
                  "detail": r})
        return finish()
    if args.workload == "closedloop":
        r = bench_closed_loop(cfg, dev, args.batch, args.steps, args.mode, world, rank, dtype_f64=args.dtype == "f64")
        if rank == 0:
            emit({"metric": r["metric"], "value": r["value"], "unit": r["unit"], "n_gpus": world, "steps": args.steps,
                  "warmup": 3, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                  "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": dict(config, workload=r["workload"]),
                  "detail": r})
        return finish()

    # ---- main workload: configs[2] ----
    total = args.batch * world
    lo, hi = shard_bounds(total, world, rank)
    prob, w = make_workload(cfg, args.mode, lo, hi, total)
    B, T = hi - lo, args.horizon
    eng = BatchedTubeMPC(prob, device=dev, dtype=dtype, gains=True)   # every output of the path is written, K and k included

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

    # Preheat (NOT the warm-up): the first launches pay lazy module loading and the SM clocks need a few hundred ms of
    # load to ramp from idle (first steps measured 2-9x slower than steady state).  Every step ends in a collective when
    # world > 1, so all ranks must run the SAME number of steps: the count is derived from all-reduced timings.
    t_pre = time.perf_counter()
    device_steps(2)
    torch.cuda.synchronize()
    est = torch.tensor([(time.perf_counter() - t_pre) / 2.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MIN)
    n_pre = 2 + int(max(0.0, args.preheat_seconds - 2 * float(est[0])) / max(float(est[0]), 1e-4))
    device_steps(n_pre - 2)
    torch.cuda.synchronize()
    # per-kernel-group durations (CUDA events around each library call, on the launching stream) and the number of
    # kernels one step launches: measured on the plain launch path, which is what the graph replays.
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
    # the W warm-up steps, immediately before the timed region
    device_steps(W)
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
    it_n, it_a = int(last.nominal.iters.sum()), int(last.auxiliary.iters.sum())
    iters_total = it_n + it_a
    kern_ms = {k: float(np.median([a.elapsed_time(b) for a, b in v])) for k, v in timers.items()}
    value = total * args.steps / (ms_total * 1e-3)
    e2e_val = total * args.steps / (ms_e2e * 1e-3)
    step_ms = ms_total / args.steps

    # ---- short measurements of the other BASELINE configs (every rank takes part; rank 0 reports) ----
    extras = None
    if not args.no_extras and not args.no_graph:
        del dbuf, graphed                      # release the graphs' private pools before the next captures
        torch.cuda.empty_cache()
        extras = {}
        try:
            if world == 1:
                cfg2 = default_dubins_config()
                r2, _ = bench_nominal(cfg2, dev, torch.float64, NOMINAL_BATCH, 20, 3)
                extras["cfg2_nominal_B4096"] = r2
                sweep = []
                for Tq, Bq, dq in ((20, 65536, "f64"), (100, 65536, "f64"), (200, 65536, "f64"), (50, 65536, "f32")):
                    sweep.append(sweep_point(default_dubins_config(), dev, args.mode, Tq, Bq, dq, steps=4))
                extras["cfg4_sweep"] = sweep
                extras["cfg1_closed_loop_B65536"] = bench_closed_loop(default_dubins_config(), dev, 65536, 10, "paper")
            extras["cfg5_adapt_shared_theta"] = bench_closed_loop(default_dubins_config(), dev, args.batch, 5, "full", world, rank)
        except Exception as e:   # noqa: BLE001  (the headline must not die with a secondary measurement)
            extras["error"] = f"{type(e).__name__}: {e}"

    if rank == 0:
        n_alpha = prob.nominal.ilqr.line_search_alphas
        flops = algorithmic_flops(T, len(n_alpha), iters_total, 2 * B, B * (2 if full else 1))
        is64 = dtype == torch.float64
        peak = ops.measure_fp64_peak_tflops() if is64 else ops.measure_fp32_peak_tflops()
        ach_model = flops / (step_ms * 1e-3) / 1e12
        bytes_solve = algorithmic_bytes_per_solve(T, full) * (1 if is64 else 0.5)
        hbm_ach = bytes_solve * B / (step_ms * 1e-3) / 1e9
        # executed work: fp64-pipe warp instructions (ncu opcode counters of THIS code, profiles/) x what the step ran
        counters = load_counters(f32=not is64)
        ne = lambda o: sum(1 for a in o.ilqr.line_search_alphas if a != 0.0)
        executed = None
        if counters:
            ex = executed_work(counters, T, (B + 31) // 32, [(ne(prob.nominal), it_n / 32.0), (ne(prob.auxiliary), it_a / 32.0)],
                               2 if full else 1)
            # one arithmetic-pipe warp instruction = 32 lanes x (1 FMA = 2 flop); the fp32 counters hold the fp32 pipes
            ach_exec = ex["fp64_inst"] * 64.0 / (step_ms * 1e-3) / 1e12
            executed = {"fp64_warp_inst_per_solve": ex["fp64_inst"] / B, "issued_warp_inst_per_solve": ex["inst"] / B,
                        "achieved": ach_exec, "frac": ach_exec / peak if peak > 0 else None,
                        "issue_frac": ex["inst"] / (step_ms * 1e-3) / (148 * 4 * (clocks or {}).get("sm_mhz", 1965.0) * 1e6)
                        if clocks and clocks.get("sm_mhz") else None,
                        "hbm_actual": {"bytes_per_solve": ex["dram_bytes"] / B, "bytes_per_step": ex["dram_bytes"],
                                       "ratio_to_algorithmic": ex["dram_bytes"] / B / bytes_solve,
                                       "achieved": ex["dram_bytes"] / (step_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                       "frac": ex["dram_bytes"] / (step_ms * 1e-3) / 1e9 / hbm_peak,
                                       "source": "dram__bytes_read+write per launch (ncu, whole batch per launch, cold L2 per "
                                                 "launch) x the launches the step ran; " + os.path.relpath(COUNTERS, ROOT)},
                        "source": os.path.relpath(COUNTERS if is64 else COUNTERS_F32, ROOT)}
        frac_exec = executed["frac"] if executed else None
        hbm_frac_act = executed["hbm_actual"]["frac"] if executed else None
        binds = None
        if executed:
            binds = ("fp64 pipe" if frac_exec >= hbm_frac_act else "hbm") + \
                f": executed fp64 pipe {frac_exec:.2f} of peak, measured DRAM traffic {hbm_frac_act:.2f} of HBM peak"
        roofline = {"bound": "fp64" if is64 else "fp32",
                    "kernel": "dtmpc_tube_step launch sequence (per OCP: init + max_iter x [ric_kernel, fwdc_kernel]; then sens_kernel; largest: fwdc_kernel, ~60 % of the step)",
                    "achieved": executed["achieved"] if executed else ach_model, "peak": peak, "unit": "TFLOP/s",
                    "frac": frac_exec if executed else (ach_model / peak if peak > 0 else None),
                    "frac_basis": "executed fp64-pipe warp instructions x 64 flop (ncu counters) / time" if executed else
                                  "flop model (no counters for this dtype)",
                    "frac_model": ach_model / peak if peak > 0 else None, "achieved_model": ach_model,
                    "peak_source": "DFMA loop measured live on this GPU (MEASURED_PEAKS.json has no FP64 figure)",
                    "flop_model": "SURVEY 8d dense reference form x actual iterations; exp=25 log=30 sincos=50 flop-eq",
                    "executed": executed, "binds": binds,
                    "traffic": executed["hbm_actual"]["bytes_per_step"] if executed else None,
                    "kernel_ms": kern_ms, "iters_per_solve": iters_total / B,
                    "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                            "bytes_per_solve": bytes_solve, "basis": "algorithmic (compulsory) bytes"}}
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            n = args.cpu_sample or cores * (1280 if full else 2048)  # ~10-20 s of CPU work
            cval, _ = cpu_run(cfg, args.mode, n, 1, 0, cores)
            cpu_baseline = {"value": cval, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{n} instances of the same seeded workload, {cores} threads, C port of the "
                                      "reference algorithm (oracle/dtmpc_oracle.c)"}
            if extras is not None and "cfg2_nominal_B4096" in extras:
                extras["cfg2_nominal_B4096"]["cpu_baseline"] = {
                    "value": cpu_run_nominal(default_dubins_config(), NOMINAL_BATCH, cores), "unit": UNIT, "cores": cores,
                    "kind": "port", "sample": f"the same {NOMINAL_BATCH} instances"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": W, "ms_per_step": step_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": dict(config, preheat_steps=n_pre + 3 + n_plain, numa=numa),
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "launch_mode": "plain" if args.no_graph else "cuda_graph_replay",
                "steps_in_flight": (int(os.environ.get("DTMPC_BENCH_BUFFERS", "2")) if overlap else 1),
                "sequential": {"value": total * args.steps / (ms_seq * 1e-3), "unit": UNIT,
                               "ms_per_step": ms_seq / args.steps,
                               "note": "same K steps with one step in flight (each replay starts after the previous one): "
                                       "the MPC latency figure; `value` overlaps two independent batches"},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
                "nonfinite_instances": n_bad, "extras": extras}
        emit(line)
    finish()


def sweep_point(cfg, dev, mode, T, B, dtype_name, steps=4):
    """One configs[3] point: the tube step at horizon T, batch B, dtype, graph replay, one step in flight."""
    import torch
    from differentiable_tube_mpc_b200.engine import BatchedTubeMPC

    cfg["system"]["horizon_N"] = T
    dtype = torch.float64 if dtype_name == "f64" else torch.float32
    prob, w = make_workload(cfg, mode, 0, B, B)
    eng = BatchedTubeMPC(prob, device=dev, dtype=dtype)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=dev)
    thn, tha = t(w["th_n"]), t(w["th_a"])
    x0n = torch.cat([t(w["xb"]), eng.init_barrier(t(w["xb"]), thn, True).view(-1, 1)], 1).contiguous()
    x0a = torch.cat([t(w["x"]), eng.init_barrier(t(w["x"]), tha, False).view(-1, 1)], 1).contiguous()
    V0 = torch.zeros(B, T, 2, dtype=dtype, device=dev)
    g = eng.graphed(x0n, x0a, V0, V0, thn, tha, full=(mode == "full"))
    for _ in range(3):
        out = g()
    ms = _timed_ms(g, steps, dev, 1)
    res = {"T": T, "B": B, "dtype": dtype_name, "mode": mode, "value": B * steps / (ms * 1e-3), "unit": UNIT,
           "ms_per_step": ms / steps, "steps": steps, "iters_per_solve": float((out.nominal.iters.sum() + out.auxiliary.iters.sum()) / B),
           "nonfinite": int((out.status != 0).sum()), "launch": "cuda_graph_replay, one step in flight"}
    counters = load_counters(f32=dtype_name != "f64")
    if counters:
        ne = lambda o: sum(1 for a in o.ilqr.line_search_alphas if a != 0.0)
        ex = executed_work(counters, T, (B + 31) // 32, [(ne(prob.nominal), float(out.nominal.iters.sum()) / 32.0),
                                                        (ne(prob.auxiliary), float(out.auxiliary.iters.sum()) / 32.0)],
                           2 if mode == "full" else 1)
        # the arithmetic pipe of this dtype (fp64: DFMA/DADD/DMUL pipe; fp32: FMA pipe) against ITS peak, measured here
        from differentiable_tube_mpc_b200 import ops
        f64 = dtype_name == "f64"
        pipe_peak = ops.measure_fp64_peak_tflops() if f64 else ops.measure_fp32_peak_tflops()
        res["pipe"] = "fp64" if f64 else "fp32 fma"
        res["pipe_tflops_executed"] = ex["fp64_inst"] * 64.0 / (ms / steps * 1e-3) / 1e12
        res["pipe_peak_tflops"] = pipe_peak
        res["pipe_frac_executed"] = res["pipe_tflops_executed"] / pipe_peak if pipe_peak > 0 else None
        res["dram_gbs_from_counters"] = ex["dram_bytes"] / (ms / steps * 1e-3) / 1e9
        res["hbm_frac_from_counters"] = res["dram_gbs_from_counters"] / float(measured_peaks().get("hbm_gbs", 6455.6))
    del g, eng
    torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    main()
