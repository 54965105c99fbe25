"""CPU tests of host-side pieces added around the hot path: the closed-form gradient forms (core/gradients.py:27-87
mirror), the run_nominal configuration mapping and its refusal to run without a CUDA device."""
import numpy as np
import pytest
import torch


def test_gradient_forms_against_explicit_loops():
    from differentiable_tube_mpc_b200.gradients import grads_auxiliary_from_deltas, grads_nominal_from_deltas

    g = torch.Generator().manual_seed(5)
    B, N = 3, 7
    X = torch.randn(B, N + 1, 4, generator=g, dtype=torch.float64)
    U = torch.randn(B, N, 2, generator=g, dtype=torch.float64)
    dX = torch.randn(B, N + 1, 4, generator=g, dtype=torch.float64)
    dU = torch.randn(B, N, 2, generator=g, dtype=torch.float64)
    Xr = torch.randn(B, N + 1, 3, generator=g, dtype=torch.float64)
    Ur = torch.randn(B, N, 2, generator=g, dtype=torch.float64)
    tgt = torch.tensor([10.0, 10.0, 0.7], dtype=torch.float64)
    ga = grads_auxiliary_from_deltas(X_hat=X, U=U, dX=dX, dU=dU, X_ref=Xr, U_ref=Ur)
    gn = grads_nominal_from_deltas(X_hat=X, U=U, dX=dX, dU=dU, target_x=tgt)
    for b in range(B):
        for grads, xr, ur in ((ga, Xr[b], Ur[b]), (gn, tgt.expand(N + 1, 3), torch.zeros(N, 2, dtype=torch.float64))):
            q = np.zeros(3); r = np.zeros(2); qb = 0.0
            for k in range(N):          # stage terms stop at k = N-1 (core/gradients.py:48-50)
                q += (2.0 * (X[b, k, :3] - xr[k]) * dX[b, k, :3]).numpy()
                r += (2.0 * (U[b, k] - ur[k]) * dU[b, k]).numpy()
                qb += float(2.0 * X[b, k, 3] * dX[b, k, 3])
            qf = (2.0 * (X[b, N, :3] - xr[N]) * dX[b, N, :3]).numpy()
            assert np.allclose(grads.Q_diag[b].numpy(), q, rtol=1e-13, atol=1e-13)
            assert np.allclose(grads.R_diag[b].numpy(), r, rtol=1e-13, atol=1e-13)
            assert np.allclose(grads.Qf_diag[b].numpy(), qf, rtol=1e-13, atol=1e-13)
            assert abs(float(grads.q_b[b]) - qb) < 1e-12
    # unbatched inputs (the reference's shapes) work too
    g1 = grads_nominal_from_deltas(X_hat=X[0], U=U[0], dX=dX[0], dU=dU[0], target_x=tgt)
    assert g1.Q_diag.shape == (3,) and torch.allclose(g1.Q_diag, gn.Q_diag[0])


def test_run_nominal_configuration_and_device_check(tmp_path):
    from differentiable_tube_mpc_b200 import TubeProblem, default_dubins_config
    from differentiable_tube_mpc_b200.run_nominal import SUCCESS_RADIUS, _nominal_ocp, run_nominal_receding

    cfg = default_dubins_config()
    cfg["system"]["horizon_N"] = 17
    cfg["system"]["nominal_max_iter"] = 6
    ocp = _nominal_ocp(cfg, TubeProblem.from_config(cfg, mode="paper"))
    assert ocp.angle_wrap and ocp.ilqr.horizon == 17 and ocp.ilqr.max_iter == 6
    assert ocp.ilqr.tol == 1e-3 and ocp.ilqr.reg == float(cfg["system"]["ilqr_reg"])      # run_nominal.py:326-334
    assert tuple(ocp.ilqr.line_search_alphas) == tuple(cfg["system"]["line_search_alphas"])
    assert SUCCESS_RADIUS == 0.25                                                          # run_nominal.py:344
    with pytest.raises(ValueError):
        run_nominal_receding(cfg, device="cpu", run_dir=str(tmp_path))
