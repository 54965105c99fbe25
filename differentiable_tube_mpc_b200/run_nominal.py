"""Nominal MPC alone, receding horizon, no disturbances: drop-in for the reference's `run_nominal.py`.

`run_nominal_receding(cfg, *, device, run_dir)` has the reference's signature, termination rules, `.npy` files and
summary (run_nominal.py:204-412): angle-wrapped heading error in the cost (:297-324), `reg = ilqr_reg`, tol 1e-3, warm
start v = v_max, omega = 0 (:336-338), per step one nominal iLQR solve (:353-364), collision test on the true minimum
over the obstacles (:376-382), success within 0.25 of the target (:385-388), warm-start shift (:391).  The solves run
in the batched CUDA path; `cfg["batch"] = B` (or an explicit `x0` of shape [B,3]) runs B start states in lock-step,
each with its own termination, and instance 0 is what the reference files hold.
"""
from __future__ import annotations

import argparse
import json
import os
from dataclasses import replace
from datetime import datetime
from typing import Any, Dict, Optional

import numpy as np
import torch
from torch import Tensor

from . import ops
from .spec import TubeProblem

SUCCESS_RADIUS = 0.25   # run_nominal.py:344


def _nominal_ocp(cfg, prob: TubeProblem):
    s = cfg["system"]
    il = replace(prob.nominal.ilqr, horizon=int(s["horizon_N"]), max_iter=int(s.get("nominal_max_iter", 10)), tol=1e-3,
                 reg=float(s.get("ilqr_reg", 1e-6)),
                 line_search_alphas=tuple(float(a) for a in s.get("line_search_alphas", [1.0, 0.5, 0.25, 0.1])))
    return replace(prob.nominal, angle_wrap=True, ilqr=il)


def run_nominal_receding(cfg: Dict[str, Any], *, device, run_dir: str, x0: Optional[Tensor] = None) -> Dict[str, Any]:
    system_cfg = cfg["system"]
    assert system_cfg["name"] == "dubins"
    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError("differentiable_tube_mpc_b200 runs on CUDA devices only (no CPU fallback)")
    dt = torch.float64                                     # run_nominal.py:225
    N, H = int(system_cfg["horizon_N"]), int(system_cfg["task_horizon_H"])
    prob = TubeProblem.from_config(cfg, mode="paper")      # fixed DBaS parameters and nominal weights from the YAML
    ocp = _nominal_ocp(cfg, prob)
    theta = torch.tensor(prob.theta_nom0, dtype=dt, device=device)
    if x0 is None:
        B = int(cfg.get("batch", 1))
        x = torch.tensor([0.0, 0.0, float(np.pi / 4)], dtype=dt, device=device).expand(B, 3).contiguous()  # :266
    else:
        x = x0.to(device=device, dtype=dt).contiguous()
        B = x.shape[0]
    xh0 = torch.cat([x, torch.zeros(B, 1, dtype=dt, device=device)], dim=1)
    b = ops.primitives_batched(prob.system, ocp, xh0, torch.zeros(B, 2, dtype=dt, device=device), theta)["b0"]
    xh = torch.cat([x, b.view(B, 1)], dim=1).contiguous()
    U_ws = torch.zeros(B, N, 2, dtype=dt, device=device)
    U_ws[:, :, 0] = float(system_cfg["control_bounds"]["v_max"])
    target = torch.tensor(list(system_cfg["target"]), dtype=dt, device=device)
    obs = cfg.get("environment", {})
    circles = obs.get("obstacles", [obs["obstacle"]] if "obstacle" in obs else [])
    centers = torch.tensor([list(o["center"]) for o in circles], dtype=dt, device=device).view(-1, 2)
    radii2 = torch.tensor([float(o["radius"]) ** 2 for o in circles], dtype=dt, device=device)

    alive = torch.ones(B, dtype=torch.bool, device=device)
    ran = torch.zeros(B, dtype=torch.int32, device=device)
    success = torch.zeros(B, dtype=torch.bool, device=device)
    collided = torch.zeros(B, dtype=torch.bool, device=device)
    success_t = torch.full((B,), -1, dtype=torch.int32, device=device)
    failed = torch.zeros((), dtype=torch.bool, device=device)
    xs, us, bs = [], [], []
    for t in range(H):
        if (t % 25) == 0:
            print(f"[nominal step {t}/{H}] ...", flush=True)
            if t > 0 and not bool(alive.any()):    # the only host synchronisation of the loop
                break
        r = ops.ilqr_solve_batched(prob.system, ocp, xh, U_ws, theta)
        failed |= ((r.status != 0) & alive).any()
        xs.append(xh[:, :3].clone()); us.append(r.V[:, 0].clone()); bs.append(xh[:, 3].clone())
        ran += alive.to(torch.int32)
        if centers.numel():
            d = xh[:, None, :2] - centers[None]
            hit = ((d * d).sum(-1) - radii2[None]).min(dim=1).values <= 0.0
        else:
            hit = torch.zeros(B, dtype=torch.bool, device=device)
        collided |= hit & alive
        alive = alive & ~hit
        near = torch.linalg.vector_norm(xh[:, :2] - target[:2], dim=1) <= SUCCESS_RADIUS
        success_t = torch.where(near & alive, torch.full_like(success_t, t), success_t)
        success |= near & alive
        alive = alive & ~near
        # x+ = f(x, u0), b+ = dbas_step(...): exactly the first step of the returned rollout (core/ddp.py:89-99)
        xh = torch.where(alive[:, None], r.X[:, 1], xh).contiguous()
        U_ws = torch.cat([r.V[:, 1:], r.V[:, -1:]], dim=1).contiguous()
    if bool(failed):
        raise FloatingPointError("nominal MPC: non-finite value in an iLQR solve")   # core/ddp.py:142

    ran_h = ran.cpu().numpy()
    n0 = int(ran_h[0])
    xs_np = torch.stack(xs)[:n0, 0].cpu().numpy().astype(np.float64)
    us_np = torch.stack(us)[:n0, 0].cpu().numpy().astype(np.float64)
    bs_np = torch.stack(bs)[:n0, 0].cpu().numpy().astype(np.float64)
    os.makedirs(run_dir, exist_ok=True)
    for name, a in (("x_bar", xs_np), ("u_bar", us_np), ("x_real", xs_np), ("u_real", us_np), ("b_real", bs_np)):
        np.save(os.path.join(run_dir, f"{name}.npy"), a)
    np.save(os.path.join(run_dir, "loss.npy"), np.zeros((xs_np.shape[0],), dtype=np.float64))
    st = int(success_t[0])
    summary = {"system": "dubins", "mode": "nominal_receding", "H_ran": n0, "success": bool(success[0]),
               "success_t": None if st < 0 else st, "collided": bool(collided[0]),
               "final_state": xs_np[-1].tolist() if xs_np.size else x[0].cpu().numpy().tolist()}
    out = {"summary": summary}
    if B > 1:   # batch extension: per-instance outcomes and the padded logs (rows >= H_ran[i] are the frozen state)
        out["batch"] = {"H_ran": ran_h, "success": success.cpu().numpy(), "collided": collided.cpu().numpy(),
                        "success_t": success_t.cpu().numpy(), "x": torch.stack(xs).cpu().numpy(),
                        "u": torch.stack(us).cpu().numpy(), "b": torch.stack(bs).cpu().numpy()}
    return out


def main() -> None:
    """`python -m differentiable_tube_mpc_b200.run_nominal --config cfg.yaml`: what run_nominal.py:415-448 does."""
    from .config import load_config

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, required=True)
    ap.add_argument("--out-dir", type=str, default=None)
    args = ap.parse_args()
    cfg = load_config(args.config)
    torch.manual_seed(int(cfg.get("seed", 0)))
    out_dir = args.out_dir or cfg.get("out_dir", "outputs")
    run_dir = os.path.join(out_dir, f"{cfg.get('run_name', 'dubins')}_nominal_{datetime.now().strftime('%Y%m%d_%H%M%S')}")
    res = run_nominal_receding(cfg, device=torch.device("cuda"), run_dir=run_dir)
    with open(os.path.join(run_dir, "config_used.json"), "w", encoding="utf-8") as f:
        json.dump(cfg, f, indent=2, ensure_ascii=False)
    with open(os.path.join(run_dir, "results_summary.json"), "w", encoding="utf-8") as f:
        json.dump(res["summary"], f, indent=2, ensure_ascii=False)
    print(f"Saved run to: {run_dir}")
    print(json.dumps(res["summary"], indent=2, ensure_ascii=False))


if __name__ == "__main__":
    main()
