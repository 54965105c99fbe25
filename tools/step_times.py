"""Diagnostic: per-step and per-library-call CUDA-event times of the bench workload (not a bench; prints a table:
rows tube_step = dtmpc_tube_step (both solves + sensitivity sweeps), reduce)."""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from differentiable_tube_mpc_b200 import default_dubins_config
from differentiable_tube_mpc_b200.engine import BatchedTubeMPC

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="full"); ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--steps", type=int, default=10); ap.add_argument("--horizon", type=int, default=50)
ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
a = ap.parse_args()
cfg = default_dubins_config(); cfg["system"]["horizon_N"] = a.horizon
dev = torch.device("cuda:0"); dt = torch.float64 if a.dtype == "f64" else torch.float32
prob, w = bench.make_workload(cfg, a.mode, 0, a.batch, a.batch)
eng = BatchedTubeMPC(prob, device=dev, dtype=dt)
cu = lambda v: torch.as_tensor(np.ascontiguousarray(v), dtype=dt, device=dev)
thn, tha = cu(w["th_n"]), cu(w["th_a"])
x0n = torch.cat([cu(w["xb"]), eng.init_barrier(cu(w["xb"]), thn, True).view(-1, 1)], 1).contiguous()
x0a = torch.cat([cu(w["x"]), eng.init_barrier(cu(w["x"]), tha, False).view(-1, 1)], 1).contiguous()
V0 = torch.zeros(a.batch, a.horizon, 2, dtype=dt, device=dev)
eng.timers = {}
tot = []
for s in range(a.steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st = eng.step(x0n, x0a, V0, V0, thn, tha, full=(a.mode == "full")); e1.record()
    torch.cuda.synchronize(); tot.append(e0.elapsed_time(e1))
print("step ms:", " ".join(f"{t:.1f}" for t in tot))
for k, v in eng.timers.items():
    print(f"{k:12s}", " ".join(f"{x.elapsed_time(y):.2f}" for x, y in v))
print("iters nom/aux mean:", float(st.nominal.iters.float().mean()), float(st.auxiliary.iters.float().mean()),
      "max", int(st.nominal.iters.max()), int(st.auxiliary.iters.max()))
