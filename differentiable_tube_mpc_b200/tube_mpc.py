"""Closed-loop tube-MPC experiment: drop-in for the reference's `run_closed_loop_experiment`.

Same signature, same mode dispatch, same `.npy` files and summary dict as reference core/tube_mpc.py:40-663 (full mode)
and :666-1048 (paper mode); the per-step solves, sensitivities, gradients, Algorithm-2 updates and the plant step run
in the batched CUDA path (`closed_loop.ClosedLoop`, no host round trip inside the loop).  Extension:
`cfg["batch"] = B` runs B tubes in lock-step (shared parameters, per-instance disturbances; Algorithm 2 then descends
the batch-summed loss); instance 0 reproduces the reference run and is what gets saved.
"""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, Optional

import numpy as np
import torch
from torch import Tensor

from .closed_loop import ClosedLoop
from .spec import is_paper_mode


def _reference_disturbance(cfg, B: int, dtype, device) -> Callable[[int], Tensor]:
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
    B = int(cfg.get("batch", 1))
    loop = ClosedLoop(cfg, device=device, batch=B, n_log=1,
                      disturbance_fn=disturbance_fn or _reference_disturbance(cfg, B, dtype, device))
    H, N = loop.H, loop.N
    logs = loop.run(H, progress=True)

    g = lambda k: logs[k].detach().cpu().numpy()
    os.makedirs(run_dir, exist_ok=True)
    x_real, b_real, loss = g("x_real")[:, 0], g("b_real")[:, 0], g("loss")[:, 0].astype(np.float64)
    np.save(os.path.join(run_dir, "x_real.npy"), x_real)
    np.save(os.path.join(run_dir, "u_real.npy"), g("u_real")[:, 0])
    np.save(os.path.join(run_dir, "x_bar.npy"), g("x_bar")[:, 0])
    np.save(os.path.join(run_dir, "u_bar.npy"), g("u_bar")[:, 0])
    np.save(os.path.join(run_dir, "b_real.npy"), b_real)
    np.save(os.path.join(run_dir, "loss.npy"), loss)
    if "Qa" in logs:
        np.save(os.path.join(run_dir, "Qa_history.npy"), g("Qa"))
        np.save(os.path.join(run_dir, "Ra_history.npy"), g("Ra"))
        np.save(os.path.join(run_dir, "qba_history.npy"), g("qba").astype(np.float64))

    summary = {
        "system": "dubins" if paper else system_cfg["name"],
        "H": H,
        "N": N,
        "final_state": x_real[-1].tolist(),
        "final_barrier_state": float(np.array(b_real[-1]).reshape(-1)[0]),
        "final_loss": float(loss[-1]),
    }
    if paper:
        summary["note"] = "Dubins run aligned to paper: nominal fixed, ancillary adapts, alpha=0, gamma=0."
    return {"summary": summary}
