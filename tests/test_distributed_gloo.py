"""Multi-rank host logic on CPU (gloo, world_size 2): sharding of the synthetic batch, the one collective of the path
(all-reduce of the batch-summed adaptation gradient, Algorithm 2 with shared theta) and rank-identical updates."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from differentiable_tube_mpc_b200 import default_dubins_config
from differentiable_tube_mpc_b200.adaptation import PaperAdapter
from differentiable_tube_mpc_b200.engine import TubeStep, allreduce_shared, shard_bounds
from differentiable_tube_mpc_b200.workload import make_instances


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_step(g_inst: torch.Tensor, loss: torch.Tensor, full: bool) -> TubeStep:
    """A TubeStep whose sums come from per-instance gradients (stands in for the CUDA engine on CPU)."""
    g13 = torch.cat([g_inst.sum(0), loss.sum().view(1)])
    return TubeStep(nominal=None, auxiliary=None, sens_aux=None, sens_nom=None, loss=loss, g_aux_sum=g13,
                    g_nom_sum=(2.0 * g_inst).sum(0) if full else None)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = default_dubins_config()
        total = 1000
        lo, hi = shard_bounds(total, world, rank)
        # every rank derives its slice from (seed, global index): the union is the single-process batch
        x_all, _ = make_instances(total, 2024, cfg)
        x_loc = x_all[lo:hi]
        gen = torch.Generator().manual_seed(5)
        g_all = torch.randn(total, 12, dtype=torch.float64, generator=gen)
        loss_all = torch.rand(total, dtype=torch.float64, generator=gen)
        out = {}
        for full in (False, True):
            st = allreduce_shared(_fake_step(g_all[lo:hi], loss_all[lo:hi], full))
            out[full] = (st.g_aux_sum.clone(), None if st.g_nom_sum is None else st.g_nom_sum.clone())
        # a graphed step hands out the SAME static tensors on every replay, refilled in place: repeated reductions of
        # one TubeStep must reduce the fresh values, not an already reduced copy (round-1 bug: the sums grew world x)
        st = _fake_step(g_all[lo:hi], loss_all[lo:hi], True)
        keep_aux, keep_nom = st.g_aux_sum, st.g_nom_sum
        reps = []
        for rep in range(3):
            fresh = _fake_step(g_all[lo:hi] * (rep + 1), loss_all[lo:hi], True)
            keep_aux.copy_(fresh.g_aux_sum)
            keep_nom.copy_(fresh.g_nom_sum)
            allreduce_shared(st)
            assert st.g_aux_sum is keep_aux and st.g_nom_sum is keep_nom
            reps.append((keep_aux.clone().numpy(), keep_nom.clone().numpy()))
        ad = PaperAdapter(Qa=torch.ones(3, dtype=torch.float64), Ra=torch.ones(2, dtype=torch.float64),
                          qba=torch.tensor(1.0, dtype=torch.float64), eta=0.05, mom=0.9)
        ad.update(out[False][0][:12] / total)
        q.put((rank, lo, hi, float(x_loc.sum()), out[False][0].numpy(), out[True][0].numpy(), out[True][1].numpy(),
               ad.theta.numpy(), reps))
    finally:
        dist.destroy_process_group()


def test_world_size_2_sharding_and_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = default_dubins_config()
    x_all, _ = make_instances(1000, 2024, cfg)
    gen = torch.Generator().manual_seed(5)
    g_all = torch.randn(1000, 12, dtype=torch.float64, generator=gen)
    loss_all = torch.rand(1000, dtype=torch.float64, generator=gen)
    want13 = torch.cat([g_all.sum(0), loss_all.sum().view(1)]).numpy()
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 1000
    assert abs(res[0][3] + res[1][3] - float(x_all.sum())) < 1e-9
    for r in res:
        assert np.allclose(r[4], want13, rtol=1e-13, atol=1e-12)          # paper mode: 12 gradients + loss
        assert np.allclose(r[5], want13, rtol=1e-13, atol=1e-12)          # full mode, theta part
        assert np.allclose(r[6], 2.0 * g_all.sum(0).numpy(), rtol=1e-13, atol=1e-12)  # full mode, theta_bar part
        for rep, (ga, gn) in enumerate(r[8]):
            assert np.allclose(ga[:12], (rep + 1) * want13[:12], rtol=1e-13, atol=1e-12), rep
            assert np.allclose(gn, 2.0 * (rep + 1) * g_all.sum(0).numpy(), rtol=1e-13, atol=1e-12), rep
    assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(res[0][8], res[1][8]))
    # identical reduced gradient -> bitwise identical parameter update on every rank
    assert np.array_equal(res[0][4], res[1][4]) and np.array_equal(res[0][7], res[1][7])


def test_allreduce_is_a_noop_without_a_process_group():
    g = torch.arange(12, dtype=torch.float64).repeat(4, 1)
    st = _fake_step(g, torch.ones(4, dtype=torch.float64), True)
    before = st.g_aux_sum.clone()
    assert allreduce_shared(st) is st and torch.equal(st.g_aux_sum, before)
