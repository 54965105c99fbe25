import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch, bench, dtmpc_oracle as orc
from differentiable_tube_mpc_b200 import default_dubins_config, ops
orc.build()
cfg = default_dubins_config(); B = 4096
prob, ocp, xb, V0 = bench.make_nominal_workload(cfg, 0, B, B)
env = orc.env_from_cfg(cfg); nom, _ = orc.ocps_from_cfg(cfg, "paper"); nom.angle_wrap = 1; nom.reg = float(cfg["system"]["ilqr_reg"])
x0 = np.concatenate([xb, np.array([orc.b0(env, nom, v) for v in xb])[:, None]], 1)
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
r = ops.ilqr_solve_batched(prob.system, ocp, cu(x0), cu(V0), cu(prob.theta_nom0), gains=True, trace=True)
it, ai, co = r.iters.cpu().numpy(), r.alpha_idx.cpu().numpy(), r.costs.cpu().numpy()
nbad = 0
for i in range(B):
    w = orc.ilqr_solve(env, nom, x0[i], V0[i])
    if w["iters"] != it[i] or not np.array_equal(w["alpha_idx"][: it[i]], ai[i][: it[i]]):
        nbad += 1
        if nbad <= 6:
            print(i, xb[i], "oracle", w["iters"], w["alpha_idx"][:w["iters"]], np.round(w["costs"][:4], 3), "gpu", it[i], ai[i][:it[i]], np.round(co[i][:4], 3))
print("mismatches", nbad, "of", B, "lib", os.environ.get("DTMPC_LIB", "in-tree"))
