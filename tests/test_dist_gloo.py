"""CPU, world_size 2 over gloo: the exchange logic of hypercube-sharded scoring (all-gather of top-K
lists / map slices + merge) reproduces the single-process result exactly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from acousticswarms_speech_b200 import dist as adist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _full_map(B, G):
    rng = np.random.default_rng(3)
    m = rng.random((B, G)).astype(np.float32)
    m[m < 0.3] = 0
    m[0, 5] = m[0, G - 2] = m[0].max()       # a tie straddling the shard boundary
    return torch.from_numpy(m)


def _topk_cpu(m, K, off):
    order = np.lexsort((np.arange(m.shape[1])[None].repeat(m.shape[0], 0), -m.numpy()), axis=-1)[:, :K]
    return torch.from_numpy(np.take_along_axis(m.numpy(), order, 1)), torch.from_numpy((order + off).astype(np.int32))


def _worker(rank, world, port, B, G, K, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = _full_map(B, G)
    g0, g1 = adist.shard_range(G, rank, world)
    sh = adist.HypercubeShardedSRP(G, lambda mix: full[:, g0:g1].clone(), _topk_cpu)
    assert (sh.g0, sh.g1) == (g0, g1)
    val, idx = sh.topk(None, K)
    fm = sh.full_map(None)
    q.put((rank, val.numpy(), idx.numpy(), fm.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_hypercube_sharding_two_ranks():
    B, G, K, world = 3, 1001, 16, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, G, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = _full_map(B, G)
    want_v, want_i = _topk_cpu(full, K, 0)
    for rank, val, idx, fm in res:
        assert np.array_equal(fm, full.numpy())                 # parity path: all-gathered map slices
        assert np.array_equal(idx, want_i.numpy())              # merged top-K identical to single process
        assert np.array_equal(val, want_v.numpy())


def _exchange_worker(rank, world, port, B, G, L, q):
    """TableExchangeSRP with stand-in stages: 'tables' = a fixed linear map of each mixture, 'gather' = a fixed
    linear map of the tables onto this rank's hypercubes.  Integer-valued float32, so sums are exact."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(9)
    mix = torch.from_numpy(rng.integers(-4, 5, (B, 3, 40)).astype(np.float32))
    A = torch.from_numpy(rng.integers(-3, 4, (40, L)).astype(np.float32))
    W = torch.from_numpy(rng.integers(-3, 4, (L, G)).astype(np.float32))
    g0, g1 = adist.shard_range(G, rank, world)
    seen = []

    def gcc_local(m):
        seen.append(m.shape[0])
        return m.sum(1) @ A

    ex = adist.TableExchangeSRP(gcc_local, lambda t: t @ W[:, g0:g1])
    sl = ex.score_slice(mix)
    sh = adist.HypercubeShardedSRP(G, ex.score_slice, _topk_cpu)
    fm = sh.full_map(mix)
    q.put((rank, sl.numpy(), fm.numpy(), seen[0], (mix.sum(1) @ A @ W).numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_table_exchange_sharding_two_ranks():
    """Transform stage sharded over mixtures, gather stage over hypercubes, one all-gather of the tables between:
    each rank transforms only its share of the (ragged: 5 over 2 ranks) batch and the result is the unsharded map."""
    B, G, L, world = 5, 37, 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, B, G, L, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, sl, fm, n_local, full in res:
        g0, g1 = adist.shard_range(G, rank, world)
        b0, b1 = adist.shard_range(B, rank, world)
        assert n_local == b1 - b0
        assert np.array_equal(sl, full[:, g0:g1])
        assert np.array_equal(fm, full)
