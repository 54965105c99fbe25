"""Closed-loop tube-MPC experiment: drop-in for the reference's `run_closed_loop_experiment`.

Same signature, same mode dispatch, same `.npy` files and summary dict as reference core/tube_mpc.py:40-663 (full mode)
and :666-1048 (paper mode); the per-step solves, sensitivities and gradients run in the batched CUDA path.  Extension:
`cfg["batch"] = B` runs B tubes in lock-step (shared parameters, per-instance disturbances; Algorithm 2 then descends
the batch-summed loss); instance 0 reproduces the reference run and is what gets logged.
"""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, Optional

import numpy as np
import torch
from torch import Tensor

from . import ops
from .adaptation import NAMES_THETA, NAMES_THETA_BAR, PaperAdapter, apply_update, split_raw_grads
from .engine import BatchedTubeMPC, allreduce_shared
from .spec import TubeProblem, is_paper_mode


def _default_disturbance(cfg, B: int, dtype, device) -> Callable[[int], Tensor]:
    """w_t = low + (high - low) * U(0,1)^3 (reference core/systems/dubins.py:57-65).  Instance 0 draws from torch's
    global CPU generator exactly like a reference run on CPU; further instances use their own seeded stream."""
    low = torch.tensor(cfg["system"]["disturbance"]["w_low"], dtype=dtype)
    high = torch.tensor(cfg["system"]["disturbance"]["w_high"], dtype=dtype)
    gen = torch.Generator().manual_seed(int(cfg.get("seed", 0)) + 7919)

    def sample(t: int) -> Tensor:
        w0 = low + (high - low) * torch.rand(3, dtype=dtype)
        if B == 1:
            return w0.view(1, 3).to(device)
        rest = low + (high - low) * torch.rand(B - 1, 3, dtype=dtype, generator=gen)
        return torch.cat([w0.view(1, 3), rest]).to(device)

    return sample


def _raise_if_failed(status: Tensor, t: int) -> None:
    bad = int((status != 0).sum())
    if bad:
        st = int(status[status != 0][0])
        which = "iLQR-nominal" if (st & 15) else ("iLQR-aux" if ((st >> 4) & 15) else "ddp_sensitivity")
        raise FloatingPointError(f"{which}: non-finite detected at t={t} in {bad} instance(s)")


def run_closed_loop_experiment(cfg: Dict[str, Any], *, device: torch.device, run_dir: str,
                               disturbance_fn: Optional[Callable[[int], Tensor]] = None) -> Dict[str, Any]:
    system_cfg = cfg["system"]
    if system_cfg["name"] != "dubins":
        raise NotImplementedError("Only dubins is wired in the skeleton; other systems are added next.")
    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError("differentiable_tube_mpc_b200 runs on CUDA devices only (no CPU fallback)")
    paper = is_paper_mode(cfg)
    dtype = torch.float64 if bool(cfg.get("use_float64", False)) else torch.float32
    prob = TubeProblem.from_config(cfg, mode="paper" if paper else "full")
    N, H = int(system_cfg["horizon_N"]), int(system_cfg["task_horizon_H"])
    B = int(cfg.get("batch", 1))
    eng = BatchedTubeMPC(prob, device=device, dtype=dtype)
    t_ = lambda v: torch.tensor(v, dtype=dtype, device=device)
    adapt_cfg = cfg.get("adaptation", {})

    # initial conditions: start at the origin facing upper-right (reference :95-97, 772-773)
    x = t_(cfg.get("x0", [0.0, 0.0, float(np.pi / 4)])).expand(B, 3).contiguous()
    x_bar = x.clone()

    if paper:
        theta_n = t_(prob.theta_nom0)
        a0 = prob.theta_aux0
        adapter = PaperAdapter(Qa=t_(a0[0:3]), Ra=t_(a0[3:5]), qba=t_(a0[8]), eta=float(adapt_cfg.get("lr_eta", 1e-2)),
                               mom=float(adapt_cfg.get("momentum", 0.9)))
        theta_a = adapter.theta()
        theta_bar = theta = None
    else:
        theta_bar, theta = prob.make_thetas(device, dtype)
        adapt_nominal = bool(adapt_cfg.get("adapt_nominal", True))
        adapt_ancillary = bool(adapt_cfg.get("adapt_ancillary", True))
        lr = float(adapt_cfg.get("lr_eta", 1e-3))
        adapt_steps = int(adapt_cfg.get("steps", 1))
        clip_norm = float(adapt_cfg.get("grad_clip_norm", 0.0))
        momentum = float(adapt_cfg.get("momentum", 0.0))
        project_params = bool(adapt_cfg.get("project_params", False))
        v_theta = [torch.zeros_like(p) for p in theta.tensors()]
        v_theta_bar = [torch.zeros_like(p) for p in theta_bar.tensors()]
        theta_n, theta_a = theta_bar.constrained(), theta.constrained()

    b = eng.init_barrier(x, theta_a, nominal=False)
    b_bar = eng.init_barrier(x_bar, theta_n, nominal=True)
    xh = torch.cat([x, b.view(B, 1)], dim=1).contiguous()
    xh_bar = torch.cat([x_bar, b_bar.view(B, 1)], dim=1).contiguous()
    Vn_ws = torch.zeros(B, N, 2, dtype=dtype, device=device)
    Va_ws = torch.zeros(B, N, 2, dtype=dtype, device=device)
    sample_w = disturbance_fn or _default_disturbance(cfg, B, dtype, device)

    log: Dict[str, list] = {k: [] for k in ("x_real", "u_real", "x_bar", "u_bar", "b_real", "loss", "Qa", "Ra", "qba")}
    p = prob
    for t in range(H):
        if (t % 25) == 0:
            print(f"[step {t}/{H}] running...", flush=True)
        if paper:
            theta_a = adapter.theta()
            st = allreduce_shared(eng.step(xh_bar, xh, Vn_ws, Va_ws, theta_n, theta_a, full=False))
            _raise_if_failed(st.status, t)
            rn, ra, loss = st.nominal, st.auxiliary, st.loss
            adapter.update(st.g_aux_sum[:12])          # momentum + projection (reference :978-984)
            theta_a_plant = theta_a  # the plant step reads only the DBaS entries, fixed in paper mode (alpha = gamma = 0)
        else:
            theta_n, theta_a = theta_bar.constrained(), theta.constrained()
            rn, ra = ops.tube_solve_batched(p.system, p.nominal, p.auxiliary, xh_bar, xh, Vn_ws, Va_ws, theta_n, theta_a,
                                            workspace=eng.workspace)
            sa = ops.ddp_sensitivity_batched(p.system, p.auxiliary, ra.X, ra.V, theta_a, X_bar=rn.X, X_ref=rn.X,
                                             U_ref=rn.V, want_grads=True, want_ref_grads=True, status=rn.status,
                                             workspace=eng.workspace)
            loss = sa.loss
            for s in range(adapt_steps):
                if s > 0:
                    # later adaptation steps re-accumulate with the updated parameters at the same optimum
                    sa2 = ops.ift_accumulate_batched(p.system, p.auxiliary, ra.X, ra.V, theta.constrained(), sa.delta_X,
                                                     sa.delta_V, sa.delta_lambda, X_ref=rn.X, U_ref=rn.V,
                                                     want_ref_grads=True)
                    sa.g_theta, sa.g_xref, sa.g_uref = sa2.g_theta, sa2.g_xref, sa2.g_uref
                g_raw_aux = ops.reduce_columns(sa.g_theta, eng.workspace) * theta.chain()
                g_xref, g_uref = sa.g_xref, sa.g_uref
                if adapt_ancillary:
                    apply_update(theta.tensors(), split_raw_grads(g_raw_aux, False), v_theta, NAMES_THETA, lr=lr,
                                 momentum=momentum, clip_norm=clip_norm, project_params=project_params)
                if not adapt_nominal:
                    continue
                gx = torch.cat([g_xref, torch.zeros(B, N + 1, 1, dtype=dtype, device=device)], dim=2)
                sn = ops.ddp_sensitivity_batched(p.system, p.nominal, rn.X, rn.V, theta_n, upper_grad_x=gx,
                                                 upper_grad_u=g_uref, want_grads=(s == 0), status=rn.status,
                                                 workspace=eng.workspace)
                if s > 0:
                    sn = ops.ift_accumulate_batched(p.system, p.nominal, rn.X, rn.V, theta_bar.constrained(), sn.delta_X,
                                                    sn.delta_V, sn.delta_lambda)
                g_raw_nom = ops.reduce_columns(sn.g_theta, eng.workspace) * theta_bar.chain()
                apply_update(theta_bar.tensors(), split_raw_grads(g_raw_nom, True), v_theta_bar, NAMES_THETA_BAR, lr=lr,
                             momentum=momentum, clip_norm=clip_norm, project_params=project_params)
            _raise_if_failed(rn.status, t)

        # logging of instance 0 (pre-step values, reference :619-629 / :1004-1013)
        log["x_real"].append(xh[0, :3].cpu().numpy().copy())
        log["u_real"].append(ra.V[0, 0].cpu().numpy().copy())
        log["x_bar"].append(xh_bar[0, :3].cpu().numpy().copy())
        log["u_bar"].append(rn.V[0, 0].cpu().numpy().copy())
        log["b_real"].append(xh[0, 3].cpu().numpy().copy())
        log["loss"].append(float(loss[0].item()))
        if paper:
            log["Qa"].append(adapter.Qa.cpu().numpy().copy())
            log["Ra"].append(adapter.Ra.cpu().numpy().copy())
            log["qba"].append(float(adapter.qba.item()))
            th_n_plant, th_a_plant = theta_n, theta_a_plant
        else:
            if adapt_ancillary:
                log["Qa"].append(theta.Q().detach().cpu().numpy().copy())
                log["Ra"].append(theta.R().detach().cpu().numpy().copy())
                log["qba"].append(float(theta.qb().detach().item()))
            # the plant and the nominal state propagate with the UPDATED DBaS parameters (reference :609, 614)
            th_n_plant, th_a_plant = theta_bar.constrained(), theta.constrained()

        # true system step with disturbance, barrier of the undisturbed successor, nominal propagation, warm-start
        # shift (reference :607-638 / :990-1023)
        w = sample_w(t)
        ops.plant_step_batched(p.system, p.nominal, p.auxiliary, rn.V, ra.V, th_n_plant, th_a_plant, w, xh, xh_bar,
                               Vn_ws, Va_ws)

    os.makedirs(run_dir, exist_ok=True)
    np.save(os.path.join(run_dir, "x_real.npy"), np.stack(log["x_real"], axis=0))
    np.save(os.path.join(run_dir, "u_real.npy"), np.stack(log["u_real"], axis=0))
    np.save(os.path.join(run_dir, "x_bar.npy"), np.stack(log["x_bar"], axis=0))
    np.save(os.path.join(run_dir, "u_bar.npy"), np.stack(log["u_bar"], axis=0))
    np.save(os.path.join(run_dir, "b_real.npy"), np.stack(log["b_real"], axis=0))
    np.save(os.path.join(run_dir, "loss.npy"), np.asarray(log["loss"], dtype=np.float64))
    if len(log["Qa"]) > 0:
        np.save(os.path.join(run_dir, "Qa_history.npy"), np.stack(log["Qa"], axis=0))
        np.save(os.path.join(run_dir, "Ra_history.npy"), np.stack(log["Ra"], axis=0))
        np.save(os.path.join(run_dir, "qba_history.npy"), np.asarray(log["qba"], dtype=np.float64))

    summary = {
        "system": "dubins" if paper else system_cfg["name"],
        "H": H,
        "N": N,
        "final_state": log["x_real"][-1].tolist(),
        "final_barrier_state": float(np.array(log["b_real"][-1]).reshape(-1)[0]),
        "final_loss": float(log["loss"][-1]),
    }
    if paper:
        summary["note"] = "Dubins run aligned to paper: nominal fixed, ancillary adapts, alpha=0, gamma=0."
    return {"summary": summary}
