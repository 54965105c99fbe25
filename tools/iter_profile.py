"""Diagnostic: per-iLQR-iteration kernel durations (CUPTI records) next to the number of still-active instances."""
import argparse, os, sys
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from differentiable_tube_mpc_b200 import default_dubins_config
from differentiable_tube_mpc_b200.engine import BatchedTubeMPC
ap = argparse.ArgumentParser(); ap.add_argument("--mode", default="full"); ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--horizon", type=int, default=50)
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
for _ in range(6): st = eng.step(x0n, x0a, V0, V0, thn, tha, full=(a.mode == "full"))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    st = eng.step(x0n, x0a, V0, V0, thn, tha, full=(a.mode == "full")); torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
short = lambda e: e.name.split("<")[0].replace("void dtmpc::", "").replace("void ", "")
rows = {"ric_kernel": [], "fwdc_kernel": []}
for e in evs:
    if short(e) in rows: rows[short(e)].append(e.time_range.end - e.time_range.start)
# one chunk per call (DTMPC_CHUNKS=1): the launches come in the order nominal iterations, auxiliary iterations
n_nom, n_aux = prob.nominal.ilqr.max_iter, prob.auxiliary.ilqr.max_iter
assert len(rows["ric_kernel"]) == n_nom + n_aux, "run with DTMPC_CHUNKS=1"
for name, res, first, n_it in (("nominal", st.nominal, 0, n_nom), ("auxiliary", st.auxiliary, n_nom, n_aux)):
    it = res.iters.cpu().numpy()
    print(f"== {name}: mean iters {it.mean():.2f}, max {it.max()}")
    print(" it   active   ric_us  fwdc_us   ric ns/inst  fwdc ns/inst")
    for j in range(n_it):
        act = int((it > j).sum()) if j > 0 else a.batch   # an instance runs iteration j iff it did not stop before it
        r, f = (rows[k][first + j] for k in ("ric_kernel", "fwdc_kernel"))
        m = max(act, 1)
        print(f"{j:3d} {act:8d} {r:8.1f} {f:8.1f}   {1e3 * r / m:9.2f}  {1e3 * f / m:9.2f}")
