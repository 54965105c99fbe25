"""Golden runs of the reference's `run_nominal.run_nominal_receding` (run_nominal.py:204-412) for the drop-in test.

Runs here only (the reference is imported from /root/reference through oracle/_ref/pyref, see make_golden.py); writes
tests/golden/nominal_receding.npz.  The only in-memory change to the reference is the overridable start state
(run_nominal.py:266); termination rules, warm start and files are the reference's own.

    python oracle/make_golden_nominal.py
"""
from __future__ import annotations

import contextlib
import copy
import importlib
import io
import math
import os
import sys
import types
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

# name -> (start state, H, N): a free run, a run that reaches the target, a start inside an obstacle, a longer approach
# to the target.  Starts whose first solves run through an
# obstacle (candidate costs ~1e22) are left out: their trajectories are chaotic for any implementation.
CASES = {
    "free": ([0.0, 0.0, math.pi / 4], 14, 20),
    "success": ([9.7, 9.65, 0.9], 40, 20),
    "collision": ([4.3, 2.4, 0.3], 5, 20),
    "approach": ([7.0, 8.5, 0.2], 40, 20),
}
HMAX = max(c[1] for c in CASES.values())


def _job(item):
    import torch

    name, (x0, H, N) = item
    _, _, cfg = mg.load_instrumented()
    rn = importlib.import_module(f"{mg.PKG}.run_nominal")
    src = open(rn.__file__, encoding="utf-8").read()
    src = mg._replace_once(src, "    x = torch.tensor([0.0, 0.0, float(np.pi / 4)], device=device, dtype=dtype)\n",
                           "    x = torch.tensor(cfg['_x0'], device=device, dtype=dtype)\n")
    mod = types.ModuleType(f"{mg.PKG}._run_nominal_startstate")
    mod.__package__ = mg.PKG
    sys.modules[mod.__name__] = mod
    exec(compile(src, "<run_nominal start state>", "exec"), mod.__dict__)
    cfg = copy.deepcopy(cfg)
    cfg["_x0"] = x0
    cfg["system"]["task_horizon_H"] = H
    cfg["system"]["horizon_N"] = N
    run_dir = os.path.join(mg.RUNS, f"nominal_receding_{name}")
    with contextlib.redirect_stdout(io.StringIO()):
        res = mod.run_nominal_receding(cfg, device=torch.device("cpu"), run_dir=run_dir)
    s = res["summary"]
    pad = lambda a, w: np.concatenate([a, np.full((HMAX - a.shape[0],) + a.shape[1:], np.nan)]) if a.shape[0] < HMAX else a
    return dict(x0=np.asarray(x0), H=H, N=N, H_ran=s["H_ran"], success=s["success"], collided=s["collided"],
                success_t=-1 if s["success_t"] is None else s["success_t"],
                x=pad(np.load(os.path.join(run_dir, "x_bar.npy")), 3), u=pad(np.load(os.path.join(run_dir, "u_bar.npy")), 2),
                b=pad(np.load(os.path.join(run_dir, "b_real.npy")).reshape(-1), 1))


def main():
    os.makedirs(mg.RUNS, exist_ok=True)
    with ProcessPoolExecutor(len(CASES)) as ex:
        res = list(ex.map(_job, CASES.items()))
    out = {k: np.stack([np.asarray(r[k]) for r in res]) for k in res[0]}
    out["names"] = np.array(list(CASES))
    np.savez_compressed(os.path.join(mg.GOLDEN, "nominal_receding.npz"), **out)
    for n, r in zip(CASES, res):
        print(n, "H_ran", r["H_ran"], "success", r["success"], r["success_t"], "collided", r["collided"])


if __name__ == "__main__":
    main()
