"""GPU tests of the execution plumbing around the kernels: CUDA-graph replay of a step, chunked multi-stream solves and
the kernel launch counter.  None of them may change a single bit of the results."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import make_instances

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(B, T, mode):
    from differentiable_tube_mpc_b200 import TubeProblem, default_dubins_config
    from differentiable_tube_mpc_b200.engine import BatchedTubeMPC

    cfg = default_dubins_config()
    cfg["system"]["horizon_N"] = T
    prob = TubeProblem.from_config(cfg, mode=mode)
    eng = BatchedTubeMPC(prob, device=DEV, dtype=torch.float64)
    x, xb = make_instances(B, 21, cfg)
    cu = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)
    if mode == "full":
        th_n_raw, th_a_raw = prob.make_thetas(DEV, torch.float64)
        thn, tha = th_n_raw.constrained().detach(), th_a_raw.constrained().detach()
    else:
        thn, tha = cu(prob.theta_nom0), cu(prob.theta_aux0)
    thn, tha = thn.expand(B, 12).contiguous(), tha.expand(B, 12).contiguous()
    x0n = torch.cat([cu(xb), eng.init_barrier(cu(xb), thn, True).view(-1, 1)], 1).contiguous()
    x0a = torch.cat([cu(x), eng.init_barrier(cu(x), tha, False).view(-1, 1)], 1).contiguous()
    V0 = torch.zeros(B, T, 2, dtype=torch.float64, device=DEV)
    return eng, (x0n, x0a, V0, V0.clone(), thn, tha)


def _same(a, b):
    for r, s in ((a.nominal, b.nominal), (a.auxiliary, b.auxiliary)):
        assert torch.equal(r.X, s.X) and torch.equal(r.V, s.V) and torch.equal(r.iters, s.iters)
        assert torch.equal(r.status, s.status)
    assert torch.equal(a.loss, b.loss) and torch.equal(a.g_aux_sum, b.g_aux_sum)
    assert torch.equal(a.sens_aux.g_theta, b.sens_aux.g_theta)
    if a.g_nom_sum is not None:
        assert torch.equal(a.g_nom_sum, b.g_nom_sum)


@pytest.mark.parametrize("mode", ["paper", "full"])
def test_graph_replay_is_the_plain_step(mode):
    B, T = 16384 + 77, 12      # large enough for two chunks (8192 instances each at least), ragged last tile
    eng, args = _problem(B, T, mode)
    plain = eng.step(*args, full=(mode == "full"))
    g = eng.graphed(*args, full=(mode == "full"))
    out = g()
    torch.cuda.synchronize()
    _same(plain, out)
    # new inputs go through the static tensors; results follow
    x0a2 = args[1].clone()
    x0a2[:, 0] += 0.05
    plain2 = eng.step(args[0], x0a2, *args[2:], full=(mode == "full"))
    keep = out.auxiliary.X.clone()
    out2 = g(x0_aux=x0a2)
    torch.cuda.synchronize()
    _same(plain2, out2)
    assert not torch.equal(keep, out2.auxiliary.X)


def test_launch_counter_counts_kernels():
    from differentiable_tube_mpc_b200 import ops

    eng, args = _problem(512, 8, "paper")
    p = eng.problem
    n0 = ops.kernel_launches()
    eng.step(*args, full=False)
    torch.cuda.synchronize()
    # init + max_iter x (ric, fwdc = line search + selection + commit) per OCP, one sensitivity sweep, two reduction kernels
    expect = (1 + 2 * p.nominal.ilqr.max_iter) + (1 + 2 * p.auxiliary.ilqr.max_iter) + 1 + 2   # B = 512: one chunk
    assert ops.kernel_launches() - n0 == expect


def test_results_do_not_depend_on_chunking_or_specialisation(tmp_path):
    """DTMPC_CHUNKS / DTMPC_NO_FIXED_NOBS are read once per process: run the same solve in fresh interpreters."""
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
from test_gpu_graph import _problem
eng, args = _problem(20000, 10, "full")
st = eng.step(*args, full=True)
torch.cuda.synchronize()
np.savez(sys.argv[1], nX=st.nominal.X.cpu().numpy(), aX=st.auxiliary.X.cpu().numpy(), aV=st.auxiliary.V.cpu().numpy(),
         it=st.auxiliary.iters.cpu().numpy(), g=st.g_aux_sum.cpu().numpy(), gn=st.g_nom_sum.cpu().numpy())
''' % (ROOT, os.path.join(ROOT, "tests"))
    outs = []
    for k, env in enumerate(({"DTMPC_CHUNKS": "1"}, {"DTMPC_CHUNKS": "2"}, {"DTMPC_CHUNKS": "1", "DTMPC_NO_FIXED_NOBS": "1"})):
        f = str(tmp_path / f"o{k}.npz")
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, "-c", code, f], check=True, env=e, timeout=600)
        outs.append(np.load(f))
    for o in outs[1:]:
        for k in outs[0].files:
            assert np.array_equal(outs[0][k], o[k]), k


@pytest.mark.parametrize("overlap", [False, True])
def test_double_buffered_steps_overlap_copies_without_changing_results(overlap):
    from differentiable_tube_mpc_b200.engine import DoubleBufferedSteps

    B, T = 4096, 10
    eng, args = _problem(B, T, "paper")
    names = ("x0_nom", "x0_aux", "Vn_init", "Va_init", "theta_nom", "theta_aux")
    variants = []
    for s in range(4):       # four different steps' inputs in pinned host memory
        a = [t.clone() for t in args]
        a[1][:, 0] += 0.01 * s
        a[1][:, 1] -= 0.02 * s
        variants.append(a)
    host = [{k: v.cpu().pin_memory() for k, v in zip(names, a)} for a in variants]
    want = [eng.step(*a, full=False) for a in variants]
    want = [(w.auxiliary.X.clone(), w.loss.clone(), w.g_aux_sum.clone()) for w in want]
    db = DoubleBufferedSteps(eng, *args, full=False, overlap_steps=overlap)   # overlap: two steps in flight
    db.prefetch(**host[0])
    gots = []
    for i in range(4):
        if i + 1 < 4:
            db.prefetch(**host[i + 1])
        out = db.run()
        gots.append((out.auxiliary.X.clone(), out.loss.clone(), out.g_aux_sum.clone()))   # read on the caller's stream
    torch.cuda.synchronize()       # no host synchronisation inside the loop: steps and copies really overlap
    for i, got in enumerate(gots):
        for g, w in zip(got, want[i]):
            assert torch.equal(g, w), i
    with pytest.raises(AssertionError):
        db.run()


@pytest.mark.parametrize("mode", ["paper", "full"])
def test_fused_tube_step_is_the_separate_calls(mode):
    """dtmpc_tube_step (solves + sensitivity sweeps per chunk, one library call) against dtmpc_tube_solve followed by
    dtmpc_sensitivity (x2): the same bits, gradients included."""
    eng, args = _problem(16384 + 40, 10, mode)
    full = mode == "full"
    fused = eng.step(*args, full=full)
    eng.unfused = True
    sep = eng.step(*args, full=full)
    eng.unfused = False
    torch.cuda.synchronize()
    _same(fused, sep)
    for a, b in ((fused.sens_aux, sep.sens_aux), (fused.sens_nom, sep.sens_nom)):
        if a is None:
            assert b is None
            continue
        assert torch.equal(a.delta_X, b.delta_X) and torch.equal(a.delta_V, b.delta_V)
        assert torch.equal(a.delta_lambda, b.delta_lambda) and torch.equal(a.g_theta, b.g_theta)
    if full:
        assert torch.equal(fused.sens_aux.g_xref, sep.sens_aux.g_xref)
        assert torch.equal(fused.sens_aux.g_uref, sep.sens_aux.g_uref)
