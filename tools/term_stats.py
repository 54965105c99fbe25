import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bench
from differentiable_tube_mpc_b200 import default_dubins_config
from differentiable_tube_mpc_b200.engine import BatchedTubeMPC
cfg = default_dubins_config(); dev = torch.device("cuda:0"); dt = torch.float64
for mode in ("full", "paper"):
    prob, w = bench.make_workload(cfg, mode, 0, 65536, 65536)
    eng = BatchedTubeMPC(prob, device=dev, dtype=dt, trace=True)
    cu = lambda v: torch.as_tensor(np.ascontiguousarray(v), dtype=dt, device=dev)
    thn, tha = cu(w["th_n"]), cu(w["th_a"])
    x0n = torch.cat([cu(w["xb"]), eng.init_barrier(cu(w["xb"]), thn, True).view(-1, 1)], 1).contiguous()
    x0a = torch.cat([cu(w["x"]), eng.init_barrier(cu(w["x"]), tha, False).view(-1, 1)], 1).contiguous()
    V0 = torch.zeros(65536, 50, 2, dtype=dt, device=dev)
    st = eng.step(x0n, x0a, V0, V0, thn, tha, full=(mode == "full"))
    for name, r, ocp in (("nom", st.nominal, prob.nominal), ("aux", st.auxiliary, prob.auxiliary)):
        it = r.iters.cpu().numpy(); ai = r.alpha_idx.cpu().numpy()
        al = np.asarray(ocp.ilqr.line_search_alphas)
        last = ai[np.arange(len(it)), it - 1]
        zero_idx = [i for i, a in enumerate(al) if a == 0.0]
        print(mode, name, "alphas", al, "mean iters", it.mean(), "max_iter", ocp.ilqr.max_iter,
              "ended at max_iter", (it == ocp.ilqr.max_iter).mean(), "last alpha==0", np.isin(last, zero_idx).mean(),
              "last alpha hist", np.bincount(last, minlength=len(al)))
        allz = np.isin(ai, zero_idx) & (np.arange(ai.shape[1])[None, :] < it[:, None])
        print("   alpha==0 picks per instance (mean)", allz.sum(1).mean(), " total alpha hist", np.bincount(ai[np.arange(ai.shape[1])[None, :] < it[:, None]], minlength=len(al)))
