"""GPU test of the `run_nominal.py` drop-in: receding-horizon nominal MPC with the reference's
termination rules, against runs of the reference itself (tests/golden/nominal_receding.npz, made by
oracle/make_golden_nominal.py)."""
import numpy as np
import pytest
import torch

from helpers import load_golden

pytestmark = pytest.mark.gpu


def _cfg(H, N):
    from differentiable_tube_mpc_b200 import default_dubins_config

    cfg = default_dubins_config()
    cfg["system"]["task_horizon_H"] = int(H)
    cfg["system"]["horizon_N"] = int(N)
    return cfg


def test_single_runs_match_the_reference(tmp_path, capsys):
    from differentiable_tube_mpc_b200.run_nominal import run_nominal_receding

    g = load_golden("nominal_receding")
    for c, name in enumerate(g["names"]):
        run_dir = str(tmp_path / str(name))
        res = run_nominal_receding(_cfg(g["H"][c], g["N"][c]), device="cuda", run_dir=run_dir,
                                   x0=torch.tensor(g["x0"][c]).view(1, 3))
        s, n = res["summary"], int(g["H_ran"][c])
        assert s["mode"] == "nominal_receding" and s["system"] == "dubins"
        assert s["H_ran"] == n and s["success"] == bool(g["success"][c]) and s["collided"] == bool(g["collided"][c])
        assert (s["success_t"] if s["success_t"] is not None else -1) == int(g["success_t"][c])
        x, u, b = (np.load(f"{run_dir}/{k}.npy") for k in ("x_bar", "u_bar", "b_real"))
        assert x.shape == (n, 3) and u.shape == (n, 2) and b.shape == (n,)
        assert np.array_equal(x, np.load(f"{run_dir}/x_real.npy")) and np.load(f"{run_dir}/loss.npy").shape == (n,)
        assert np.abs(x - g["x"][c][:n]).max() < 1e-8, name
        assert np.abs(u - g["u"][c][:n]).max() < 1e-6, name     # controls of a flat optimum: looser than states
        assert np.abs(b - g["b"][c][:n]).max() < 1e-8 * max(1.0, np.abs(b).max()), name
        assert np.allclose(s["final_state"], g["x"][c][n - 1], atol=1e-8)
    assert "[nominal step 0/" in capsys.readouterr().out


def test_batched_run_has_per_instance_termination(tmp_path):
    from differentiable_tube_mpc_b200.run_nominal import run_nominal_receding

    g = load_golden("nominal_receding")
    H = int(g["H"].max())
    res = run_nominal_receding(_cfg(H, g["N"][0]), device="cuda", run_dir=str(tmp_path / "b"), x0=torch.tensor(g["x0"]))
    bt = res["batch"]
    for c, name in enumerate(g["names"]):
        n = int(g["H_ran"][c])
        stopped = bool(g["success"][c]) or bool(g["collided"][c])
        assert bool(bt["success"][c]) == bool(g["success"][c]) and bool(bt["collided"][c]) == bool(g["collided"][c])
        if stopped:     # the golden run ended by its own rule, so the step counts agree
            assert int(bt["H_ran"][c]) == n and int(bt["success_t"][c]) == int(g["success_t"][c])
        assert np.abs(bt["x"][:n, c] - g["x"][c][:n]).max() < 1e-8, name
    # instance 0 is what the reference files hold
    assert np.array_equal(np.load(str(tmp_path / "b" / "x_bar.npy")), bt["x"][: int(bt["H_ran"][0]), 0])
    with pytest.raises(ValueError):
        run_nominal_receding(_cfg(2, 10), device="cpu", run_dir=str(tmp_path / "c"))
