"""Real (non-serialised) per-kernel GPU time of one bench step, from torch.profiler (CUPTI activity records)."""
import argparse, os, sys, collections
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from differentiable_tube_mpc_b200 import default_dubins_config
from differentiable_tube_mpc_b200.engine import BatchedTubeMPC
ap = argparse.ArgumentParser(); ap.add_argument("--mode", default="full"); ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--horizon", type=int, default=50); ap.add_argument("--timeline", action="store_true")
a = ap.parse_args()
cfg = default_dubins_config(); cfg["system"]["horizon_N"] = a.horizon
dev = torch.device("cuda:0"); dt = torch.float64
prob, w = bench.make_workload(cfg, a.mode, 0, a.batch, a.batch)
eng = BatchedTubeMPC(prob, device=dev, dtype=dt)
cu = lambda v: torch.as_tensor(np.ascontiguousarray(v), dtype=dt, device=dev)
thn, tha = cu(w["th_n"]), cu(w["th_a"])
x0n = torch.cat([cu(w["xb"]), eng.init_barrier(cu(w["xb"]), thn, True).view(-1, 1)], 1).contiguous()
x0a = torch.cat([cu(w["x"]), eng.init_barrier(cu(w["x"]), tha, False).view(-1, 1)], 1).contiguous()
V0 = torch.zeros(a.batch, a.horizon, 2, dtype=dt, device=dev)
for _ in range(6): eng.step(x0n, x0a, V0, V0, thn, tha, full=(a.mode == "full"))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng.step(x0n, x0a, V0, V0, thn, tha, full=(a.mode == "full")); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.OrderedDict(); 
for e in evs:
    n = e.name.split("<")[0].replace("void dtmpc::", "").replace("void ", "")[:30]
    d = agg.setdefault(n, [0, 0.0]); d[0] += 1; d[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
print(f"span {(t1 - t0) / 1e3:.2f} ms, sum of kernel durations {tot / 1e3:.2f} ms")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]: print(f"{t / 1e3:9.3f} ms {c:5d}x {100 * t / tot:5.1f}%  {n}")
if a.timeline:
    for e in sorted(evs, key=lambda e: e.time_range.start)[:140]:
        print(f"{(e.time_range.start - t0):9.1f} us  +{e.time_range.end - e.time_range.start:8.1f}  {e.name.split('<')[0][-24:]}")
