"""world_size-2 gloo tests (CPU) of the multi-process plumbing used for chain sharding: rank-order gather of
end states, weight broadcast, and the bit-exact replay of the bookkeeping on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nested as onested


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from nnest_b200 import dist as nd
        from nnest_b200.networks import SingleSpeedNVP
        dev = torch.device('cpu')
        assert nd.is_distributed() and nd.rank_world() == (rank, world)
        assert nd.chain_offset(8) == rank * 8
        # (1) rank-order gather == single-process concatenate (nested.py:425-427)
        rng = np.random.RandomState(100)
        full_last = rng.normal(size=(world * 8, 3)).astype(np.float32)
        full_first = full_last + (rng.uniform(size=full_last.shape) > 0.3).astype(np.float32)
        full_logl = rng.normal(size=(world * 8,))
        sl = slice(rank * 8, (rank + 1) * 8)
        g_first = nd.allgather_rows(torch.from_numpy(full_first[sl])).numpy()
        g_last = nd.allgather_rows(torch.from_numpy(full_last[sl])).numpy()
        g_logl = nd.allgather_rows(torch.from_numpy(full_logl[sl])).numpy()
        assert np.array_equal(g_first, full_first) and np.array_equal(g_last, full_last)
        assert np.array_equal(g_logl, full_logl)
        # (2) live points come from rank 0 (nested.py:199-205)
        u = np.random.RandomState(rank).uniform(-1, 1, size=(16, 3))
        u = nd.broadcast_array(u, dev)
        assert np.array_equal(u, np.random.RandomState(0).uniform(-1, 1, size=(16, 3)))
        # (3) flow weights: one flat broadcast, every rank ends with rank 0's parameters
        torch.manual_seed(rank)
        net = SingleSpeedNVP(3, 16, 3, 1)
        nd.broadcast_parameters(net)
        torch.manual_seed(0)
        ref = SingleSpeedNVP(3, 16, 3, 1)
        for a, b in zip(net.parameters(), ref.parameters()):
            assert torch.equal(a, b)
        assert nd.allreduce_sum_int(rank + 5, dev) == sum(r + 5 for r in range(world))
        # (3b) ragged shares of the initial live-point likelihoods (contiguous shards, rank order)
        full = np.random.RandomState(7).normal(size=13)
        lo, hi = nd.shard_bounds(13, rank, world)
        assert [nd.shard_bounds(13, r, world) for r in range(world)][-1][1] == 13
        assert np.array_equal(nd.allgather_ragged(full[lo:hi], 13, dev), full)
        # (4) every rank replays the same bookkeeping on the gathered batch -> identical state everywhere
        logl0 = np.random.RandomState(1).normal(size=16)
        st = onested.NSState(16)
        au, al = u.copy(), logl0.copy()
        av = 5 * au
        worst, loglstar = onested.iteration_head(st, av, al)
        samples = np.stack([g_first, g_last], axis=1)
        loglikes = np.stack([np.zeros_like(g_logl), g_logl], axis=1)
        onested.consume_mcmc(st, samples, loglikes, worst, loglstar, au, av, al, lambda x: 5 * x)
        digest = torch.tensor([float(st.nb), float(worst), float(al.sum()), float(au.sum())], dtype=torch.float64)
        parts = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(parts, digest)
        assert all(torch.equal(p, parts[0]) for p in parts)
        q.put((rank, 'ok'))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, 'ok'), (1, 'ok')], res
