#!/usr/bin/env python
"""BASELINE configs[3]: horizon / batch / precision sweep of the tube step on ONE GPU (graph replay, one step in flight,
inputs resident; per-instance theta, full mode).  Prints a markdown table (committed as profiles/r02_sweep.md).

    python tools/sweep.py [--quick]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from differentiable_tube_mpc_b200 import default_dubins_config, ops

ap = argparse.ArgumentParser(); ap.add_argument("--quick", action="store_true"); ap.add_argument("--mode", default="full")
a = ap.parse_args()
dev = torch.device("cuda:0")
grid = []
for T in (20, 50, 100, 200):
    for B in ((4096, 65536) if a.quick else (4096, 16384, 65536, 262144)):
        grid.append((T, B, "f64"))
for B in (4096, 16384, 65536, 262144):
    grid.append((50, B, "f32"))
if not a.quick:
    grid += [(50, 1048576, "f64"), (50, 1048576, "f32"), (200, 65536, "f32")]
peak64, peak32 = ops.measure_fp64_peak_tflops(), ops.measure_fp32_peak_tflops()
hbm = float(bench.measured_peaks().get("hbm_gbs", 6455.6))
print(f"peaks measured on this GPU: fp64 DFMA {peak64:.1f} TFLOP/s, fp32 FFMA {peak32:.1f} TFLOP/s; HBM {hbm:.0f} GB/s (MEASURED_PEAKS.json)\n")
print("| T | B | dtype | ms / step | solves/s | iterations / solve | arithmetic pipe (executed) | DRAM (from counters) | non-finite |")
print("|---|---|---|---|---|---|---|---|---|")
for T, B, dt in grid:
    r = bench.sweep_point(default_dubins_config(), dev, a.mode, T, B, dt, steps=3 if B >= 262144 else 5)
    fr = f"{r['pipe_tflops_executed']:.1f} TFLOP/s = {r['pipe_frac_executed']:.2f}" if "pipe_frac_executed" in r else "-"
    dr = f"{r['dram_gbs_from_counters'] / 1e3:.2f} TB/s = {r['hbm_frac_from_counters']:.2f}" if "dram_gbs_from_counters" in r else "-"
    print(f"| {T} | {B} | {dt} | {r['ms_per_step']:.2f} | {r['value'] / 1e6:.2f} M | {r['iters_per_solve']:.1f} | {fr} | {dr} | {r['nonfinite']} |", flush=True)
