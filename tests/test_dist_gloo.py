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
