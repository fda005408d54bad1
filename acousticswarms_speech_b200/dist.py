"""Multi-GPU sharding of the scoring path (one process per GPU, torch.distributed for the plumbing).

The reference has no distributed code (only nn.DataParallel over the patch batch of the spot model,
sep/training/JointModel/network.py:30).  The path shards along two independent axes:

* mixtures  -- every rank holds the full lag table and scores its own contiguous slice of the batch;
               no data-path communication at all.
* hypercubes -- rank r owns hypercubes [g0, g1) of the lag table and scores ALL mixtures on that slice
               (the cheap STFT/cross-spectrum stage is recomputed redundantly per rank).  Each map value
               is computed wholly on one rank, so results are bit-identical to a single GPU.  The one
               collective is an all-gather of the per-rank top-K (value, global index) lists -- or of
               the map slices themselves when the host pruning needs the full map -- followed by a merge.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced split of range(n): the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def merge_topk(vals, idxs, K):
    """Merge candidate lists (..., n) -> the K best, descending by value, ties to the lower global index
    (the same order asw_map_topk produces).  Entries with index < 0 are padding."""
    v = vals.clone()
    v[idxs < 0] = float("-inf")
    # stable two-key sort: by index ascending first, then by value descending (stable keeps index order in ties)
    order = torch.argsort(idxs.to(torch.int64), dim=-1, stable=True)
    v1 = torch.gather(v, -1, order)
    i1 = torch.gather(idxs, -1, order)
    order2 = torch.argsort(v1, dim=-1, descending=True, stable=True)
    K = min(K, v.shape[-1])
    order2 = order2[..., :K]
    return torch.gather(v1, -1, order2), torch.gather(i1, -1, order2)


def allgather_topk(val, idx, K, group=None):
    """All-gather every rank's (B, K) top-K lists and merge them into the global (B, K) list."""
    world = dist.get_world_size(group)
    if world == 1:
        return val, idx
    vs = [torch.empty_like(val) for _ in range(world)]
    ix = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(vs, val.contiguous(), group=group)
    dist.all_gather(ix, idx.contiguous(), group=group)
    return merge_topk(torch.cat(vs, dim=-1), torch.cat(ix, dim=-1), K)


def allgather_map(map_slice, G, group=None):
    """All-gather hypercube slices (B, g1 - g0) of every rank into the full (B, G) map (parity path:
    the reference's voxel-neighbourhood peak test needs every value)."""
    world = dist.get_world_size(group)
    if world == 1:
        return map_slice
    B = map_slice.shape[0]
    width = max(shard_range(G, r, world)[1] - shard_range(G, r, world)[0] for r in range(world))
    pad = torch.zeros((B, width), device=map_slice.device, dtype=map_slice.dtype)
    pad[:, :map_slice.shape[1]] = map_slice
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.empty((B, G), device=map_slice.device, dtype=map_slice.dtype)
    for r in range(world):
        g0, g1 = shard_range(G, r, world)
        out[:, g0:g1] = parts[r][:, :g1 - g0]
    return out


class HypercubeShardedSRP:
    """Scores mixtures on this rank's slice of the hypercubes and exchanges results.

    ``score_slice(mix) -> (B, g1 - g0)`` and ``topk_slice(map_slice, K, idx_offset) -> (val, idx)`` are the
    device kernels (native.NativeSRP.score / native.map_topk on a GPU); they are injected so the exchange
    logic is testable with gloo on CPU."""

    def __init__(self, G, score_slice, topk_slice, group=None):
        self.G = G
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.g0, self.g1 = shard_range(G, self.rank, self.world)
        self.score_slice = score_slice
        self.topk_slice = topk_slice

    def topk(self, mix, K):
        m = self.score_slice(mix)
        val, idx = self.topk_slice(m, K, self.g0)
        if self.world == 1:
            return val, idx
        return allgather_topk(val, idx, K, self.group)

    def full_map(self, mix):
        m = self.score_slice(mix)
        if self.world == 1:
            return m
        return allgather_map(m, self.G, self.group)


def native_sharded_srp(lag_samples, num_mic, device, group=None, **kw):
    """HypercubeShardedSRP over libasw.so: this rank's NativeSRP holds lag rows [g0, g1)."""
    from . import native
    from .constants import window_length
    G = lag_samples.shape[0]
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    g0, g1 = shard_range(G, rank, world)
    h = native.NativeSRP(lag_samples[g0:g1], num_mic, device=device, **kw)

    def score(mix):
        return h.score(mix, window_length(mix.shape[-1]))

    return HypercubeShardedSRP(G, score, native.map_topk, group), h
