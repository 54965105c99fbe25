#!/usr/bin/env python
"""Data-parallel Algorithm 2 with SHARED theta on WORLD ranks (BASELINE configs[4]): every rank runs its contiguous
shard of one seeded global batch through `ClosedLoop`, the batch-summed gradient is summed over the ranks, and every
rank applies the same momentum / projection update (reference core/tube_mpc.py:187-249, 978-984).

Launched by tests/test_gpu_loop.py (two processes on one GPU, gloo) and by hand / gpurun on N GPUs:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/dist_adapt.py --world N ...
Each rank writes <out>/w<WORLD>_r<RANK>.npz; `--check` (rank 0) compares against <out>/w1_r0.npz from a 1-rank run.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="paper", choices=["paper", "full"])
    ap.add_argument("--batch", type=int, default=4096, help="GLOBAL batch")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--horizon", type=int, default=50)
    ap.add_argument("--world", type=int, default=1)
    ap.add_argument("--out", required=True)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    assert world == a.world
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        backend = os.environ.get("DTMPC_DIST_BACKEND", "nccl")
        dist.init_process_group(backend, rank=rank, world_size=world, **({"device_id": dev} if backend == "nccl" else {}))
    from differentiable_tube_mpc_b200 import default_dubins_config
    from differentiable_tube_mpc_b200.closed_loop import ClosedLoop
    from differentiable_tube_mpc_b200.engine import shard_bounds
    from differentiable_tube_mpc_b200.workload import make_instances

    cfg = default_dubins_config()
    cfg["system"]["horizon_N"] = a.horizon
    cfg["use_float64"] = True
    cfg["adaptation"]["adapt_nominal"] = a.mode == "full"
    cfg["adaptation"]["lr_eta"] = 1e-3 if a.mode == "full" else 0.05 / a.batch   # the summed gradient grows with the batch
    lo, hi = shard_bounds(a.batch, world, rank)
    x_all, xb_all = make_instances(a.batch, 2024, cfg)
    t = lambda v: torch.as_tensor(np.ascontiguousarray(v), dtype=torch.float64, device=dev)
    loop = ClosedLoop(cfg, device=dev, batch=hi - lo, n_log=hi - lo, x0=t(x_all[lo:hi]), x_bar0=t(xb_all[lo:hi]),
                      first_instance=lo)
    th_a0 = loop.theta_a.clone()
    g_hist = []
    loop.reserve(a.steps)
    for _ in range(a.steps):
        loop.step()
        st = loop.last
        g_hist.append(torch.cat([st.g_aux_sum] + ([st.g_nom_sum] if st.g_nom_sum is not None else [])).cpu().numpy())
    loop.check()
    logs = loop.logs()
    os.makedirs(a.out, exist_ok=True)
    np.savez(os.path.join(a.out, f"w{world}_r{rank}.npz"), theta_a=loop.theta_a.cpu().numpy(), theta_n=loop.theta_n.cpu().numpy(),
             theta_a0=th_a0.cpu().numpy(), g_hist=np.stack(g_hist), u_real=logs["u_real"].cpu().numpy(),
             loss=logs["loss"].cpu().numpy())
    if world > 1:
        dist.barrier()
    if a.check and rank == 0:
        one = np.load(os.path.join(a.out, "w1_r0.npz"))
        g = np.stack(g_hist)
        dg = float(np.max(np.abs(g - one["g_hist"]) / np.abs(one["g_hist"]).max(axis=1, keepdims=True)))
        dth = float(max(np.max(np.abs(loop.theta_a.cpu().numpy() - one["theta_a"])), np.max(np.abs(loop.theta_n.cpu().numpy() - one["theta_n"]))))
        same = [np.array_equal(np.load(os.path.join(a.out, f"w{world}_r{r}.npz"))["theta_a"], loop.theta_a.cpu().numpy())
                for r in range(world)]
        print(f"[dist_adapt] world {world} vs 1 rank, global batch {a.batch}, {a.steps} steps, {a.mode} mode: "
              f"summed gradient rel. diff {dg:.2e}, theta abs. diff {dth:.2e}, ranks bitwise identical: {all(same)}", flush=True)
        assert dg < 1e-11 and dth < 1e-11 and all(same)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
