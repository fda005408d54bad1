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
    (the same order asw_map_topk produces).  Entries with index < 0 are padding.
    One sort of 64-bit keys (order-preserving integer image of the float32 value, inverted index)."""
    v = vals.to(torch.float32).clone()
    v[idxs < 0] = float("-inf")
    v = v + 0.0                                   # -0.0 -> +0.0: equal values must have equal bit patterns
    bits = v.contiguous().view(torch.int32)
    bits = bits ^ ((bits >> 31) & 0x7FFFFFFF)     # negative floats: flip the magnitude bits so that ints order like floats
    key = (bits.to(torch.int64) << 32) | (0xFFFFFFFF - idxs.to(torch.int64).clamp(min=0))
    K = min(K, v.shape[-1])
    order = torch.sort(key, dim=-1, descending=True).indices[..., :K]
    return torch.gather(v, -1, order).to(vals.dtype), torch.gather(idxs, -1, order)


def allgather_topk(val, idx, K, group=None):
    """All-gather every rank's (B, K) top-K lists (one collective: values and indices packed side by side) and
    merge them into the global (B, K) list."""
    world = dist.get_world_size(group)
    if world == 1:
        return val, idx
    B, k = val.shape
    pack = torch.empty((B, 2 * k), device=val.device, dtype=torch.float32)
    pack[:, :k] = val
    pack[:, k:] = idx.to(torch.int32).view(torch.float32)
    flat = torch.empty((world * B, 2 * k), device=val.device, dtype=torch.float32)   # concatenated along dim 0
    dist.all_gather_into_tensor(flat, pack, group=group)
    gathered = flat.view(world, B, 2 * k)
    vs = gathered[:, :, :k].permute(1, 0, 2).reshape(B, world * k)
    ix = gathered[:, :, k:].contiguous().view(torch.int32).permute(1, 0, 2).reshape(B, world * k)
    return merge_topk(vs, ix, K)


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


class TableExchangeSRP:
    """Hypercube sharding that does not recompute the transform stage on every rank.

    With plain hypercube sharding every rank runs the STFT / cross-spectrum / GCC stage for ALL mixtures, which is
    two thirds of the scoring time at 7 mics -- two GPUs are then slower than one.  Here the transform stage is
    sharded over MIXTURES and the gather stage over HYPERCUBES, with one all-gather of the GCC lag tables in
    between (1.2 MB per mixture at 7 mics / 3 s):

        rank r:  tables[b0:b1] = gcc_local(mix[b0:b1])          (its share of the mixtures)
        all:     tables[0:B]   = all_gather(tables[b0:b1])      (NCCL over NVLink)
        rank r:  map[:, g0:g1] = gather_slice(tables[0:B])      (its share of the hypercubes)

    ``gcc_local`` / ``gather_slice`` are injected (native.NativeSRP.gcc / .gather on a GPU) so that the exchange
    logic is testable with gloo on CPU.  Every rank's handle must use the same table layout: see
    ``native_table_exchange_srp``."""

    def __init__(self, gcc_local, gather_slice, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.gcc_local = gcc_local
        self.gather_slice = gather_slice

    def score_slice(self, mix):
        """mix (B, M, T), the same on every rank -> this rank's (B, g1 - g0) slice of the map."""
        B = mix.shape[0]
        b0, b1 = shard_range(B, self.rank, self.world)
        local = self.gcc_local(mix[b0:b1].contiguous())
        if self.world == 1:
            return self.gather_slice(local)
        per = -(-B // self.world)
        buf = local
        if B % self.world:                      # ragged split: every rank contributes `per` rows, the last ones padding
            buf = local.new_zeros((per, local.shape[1]))
            buf[:b1 - b0] = local
        gathered = local.new_empty((self.world * per, local.shape[1]))
        dist.all_gather_into_tensor(gathered, buf, group=self.group)
        if B % self.world:                      # drop every rank's padding rows
            rows = [r * per + i for r in range(self.world)
                    for i in range(shard_range(B, r, self.world)[1] - shard_range(B, r, self.world)[0])]
            gathered = gathered[torch.as_tensor(rows, device=gathered.device)]
        return self.gather_slice(gathered)


def native_table_exchange_srp(lag_samples, num_mic, device, group=None, **kw):
    """HypercubeShardedSRP whose slices come from a TableExchangeSRP over libasw.so.  Every rank's handle is built
    from its lag rows [g0, g1) plus two extra rows holding the per-pair minimum and maximum lag of ALL hypercubes, so
    that the lag-table layout (first lag, length) is the one a single-GPU handle would use: tables are then
    exchangeable between ranks and every map value is bit-identical to the single-GPU one."""
    import numpy as np
    from . import native
    from .constants import window_length
    lag = np.ascontiguousarray(lag_samples, dtype=np.float64)
    G = lag.shape[0]
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    g0, g1 = shard_range(G, rank, world)
    rows = np.concatenate([lag[g0:g1], lag.min(0, keepdims=True), lag.max(0, keepdims=True)])
    h = native.NativeSRP(rows, num_mic, device=device, **kw)
    state = {}

    def gcc_local(mix):
        state["T"] = mix.shape[-1]
        return h.gcc(mix, window_length(mix.shape[-1]))

    def gather_slice(tables):
        T = state["T"]
        return h.gather(tables, h.num_windows(T, window_length(T)))[:, :g1 - g0].contiguous()

    ex = TableExchangeSRP(gcc_local, gather_slice, group)
    return HypercubeShardedSRP(G, ex.score_slice, native.map_topk, group), h
