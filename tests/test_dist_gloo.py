"""N > 1 path on CPU: world_size-2 gloo all-gather of the per-shard records + the log-sum-exp merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from m3pc_b200 import dist as mdist
from m3pc_b200._native import PARTIAL_FLOATS


def test_shard_range_partitions_everything():
    for n in (1, 7, 625, 1024, 16384):
        for world in (1, 2, 3, 4, 8):
            r = [mdist.shard_range(n, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        mdist.shard_range(10, 2, 2)


def _problem(n=1000, A=6, seed=0):
    rs = np.random.RandomState(seed)
    J = (rs.randn(n) * 3 + 100).astype(np.float32)
    a0 = rs.uniform(-1, 1, size=(n, A)).astype(np.float32)
    q = rs.exponential(1.0, n).astype(np.float32)
    return J, a0, q


def _unsharded(J, a0, q, tau):
    w = np.exp((J.astype(np.float64) - J.max()) * tau)
    p = w / w.sum()
    return (a0 * p[:, None]).sum(0) / p.sum(), int(np.argmax(J)), int(np.argmax(p / q))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("tau", [1.0, 0.01])
def test_merge_is_shard_count_invariant(world, tau):
    J, a0, q = _problem()
    recs = []
    for g in range(world):
        lo, hi = mdist.shard_range(len(J), g, world)
        recs.append(mdist.make_partial_host(J[lo:hi], a0[lo:hi], q[lo:hi], tau, lo))
    ev, sm, amax, sidx = mdist.merge_partials_host(np.stack(recs), a0.shape[1], tau)
    ref_ev, ref_amax, ref_sidx = _unsharded(J, a0, q, tau)
    np.testing.assert_allclose(ev, ref_ev, rtol=2e-5, atol=2e-6)
    assert amax == ref_amax and sidx == ref_sidx
    assert np.array_equal(sm, a0[ref_sidx])


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, lr, w = mdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    J, a0, q = _problem()
    lo, hi = mdist.shard_range(len(J), rank, world)
    rec = torch.from_numpy(mdist.make_partial_host(J[lo:hi], a0[lo:hi], q[lo:hi], 1.0, lo))
    g = mdist.gather_partials(rec)
    assert g.shape == (world, PARTIAL_FLOATS)
    ev, sm, amax, sidx = mdist.merge_partials_host(g.numpy(), a0.shape[1], 1.0)
    # every rank must hold the same merged answer
    t = torch.from_numpy(np.concatenate([ev, sm, [amax, sidx]]).astype(np.float64))
    lo_t, hi_t = t.clone(), t.clone()
    dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
    assert torch.equal(lo_t, hi_t)
    if rank == 0:
        ret["ev"], ret["amax"], ret["sidx"] = ev, amax, sidx
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_all_gather_and_merge():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        J, a0, q = _problem()
        ref_ev, ref_amax, ref_sidx = _unsharded(J, a0, q, 1.0)
        np.testing.assert_allclose(ret["ev"], ref_ev, rtol=2e-5, atol=2e-6)
        assert ret["amax"] == ref_amax and ret["sidx"] == ref_sidx


def _worker_plumbing(rank, world, port, ret):
    """Host-side plumbing of the peer exchange and of the weight broadcast on gloo: every rank ends up with all handles in rank
    order, and with rank 0's parameters."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    mdist.init_from_env("gloo")
    got = mdist.all_gather_bytes(bytes([rank]) * 64)
    assert got == [bytes([g]) * 64 for g in range(world)]

    class FakeEngine:  # records what connect_exchange hands to m3pc_exchange_connect
        def exchange_local(self):
            return bytes([100 + rank]) * 64, 0

        def exchange_connect(self, r, w, ipc_handles=None, device_ptrs=None):
            self.args = (r, w, list(ipc_handles))

    fe = FakeEngine()
    assert mdist.connect_exchange(fe) == (rank, world)
    assert fe.args == (rank, world, [bytes([100 + g]) * 64 for g in range(world)])
    torch.manual_seed(rank)  # different initial weights per rank
    m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.LayerNorm(7))
    m.register_buffer("pos", torch.randn(3, 2))
    nbytes = mdist.broadcast_parameters(m, src=0)
    assert nbytes == 4 * (5 * 7 + 7 + 7 + 7 + 6)
    flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()] + [m.pos.reshape(-1)]).double()
    lo_t, hi_t = flat.clone(), flat.clone()
    dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
    assert torch.equal(lo_t, hi_t)
    if rank == 0:
        torch.manual_seed(0)
        ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.LayerNorm(7))
        assert torch.equal(ref[0].weight, m[0].weight)  # the source keeps its own values
        ret["ok"] = True
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_exchange_plumbing_and_weight_broadcast():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_plumbing, args=(2, port, ret), nprocs=2, join=True)
        assert ret["ok"]
