"""Batched front end: SRP-PHAT scoring of all hypercubes for a batch of mixtures that share one
array geometry, MAX_POWER / top-K, and the per-hypercube shift-stack into reused network-batch
buffers -- the device-resident form of
    Mic_Array.Apply_SRP_PHAT  ->  Spotform_Big_Patch's shift_and_sep
(sep/Mic_Array.py:152-222, sep/training/JointModel/network.py:37-104) for B mixtures at once.

The reference processes one mixture at a time and reuses one (128, M, T) device buffer for every
batch of patches (network.py:58); here a small ring of such buffers is written batch by batch so the
consumer (the PyTorch separator) can overlap with the next batch's shift-stack.
"""
import numpy as np
import torch

from . import native
from .constants import SPOT_BATCH_SIZE, window_length


class FrontEnd:
    def __init__(self, srp_node, device=None, net_batch=SPOT_BATCH_SIZE, ring=2, topk=128, launch_batches=1):
        """``srp_node``: an ``acousticswarms_speech_b200.srp_phat.SRP_PHAT`` (geometry + device handle).
        ``ring``: how many (net_batch, M, T) network-input buffers are cycled through; ``launch_batches``: how many
        of them one shift-stack launch fills (the reference fills one 128-patch buffer at a time,
        sep/training/JointModel/network.py:58; with 180 GB of HBM a deeper ring lets one launch stack several network
        batches ahead of the consumer and saves the launch gaps)."""
        self.node = srp_node
        self.h = srp_node.native
        self.device = srp_node.device if device is None else torch.device(device)
        self.net_batch = net_batch
        self.ring = max(ring, launch_batches)
        self.launch_batches = max(1, launch_batches)
        self.topk = topk
        self._bufs = None
        self._map = None

    # ---- scoring ---------------------------------------------------------------------------------
    def score(self, mix_dev, out=None):
        """(B, M, T) float32 CUDA -> (map (B, G), top values (B, K), top indices (B, K)).
        ``out``: caller-owned map buffer (for pipelines that keep several steps in flight)."""
        B, M, T = mix_dev.shape
        if out is None:
            if self._map is None or self._map.shape[0] != B:
                self._map = torch.empty((B, self.h.G), device=self.device, dtype=torch.float32)
            out = self._map
        self.h.score(mix_dev, window_length(T), out=out)
        val, idx = native.map_topk(out, min(self.topk, 1024))
        return out, val, idx

    # ---- shift-stack -----------------------------------------------------------------------------
    def _ring(self, M, T):
        """The ring as ``ring // launch_batches`` contiguous segments of ``launch_batches`` network batches each."""
        rows = self.net_batch * self.launch_batches
        if self._bufs is None or self._bufs[0].shape != (rows, M, T):
            self._bufs = [torch.empty((rows, M, T), device=self.device, dtype=torch.float32)
                          for _ in range(max(1, self.ring // self.launch_batches))]
        return self._bufs

    def stack(self, mix_dev, shifts_dev, mix_index_dev, consumer=None, fused_norm=False, events=None):
        """Write every patch's shifted (M, T) block, ``net_batch`` patches per launch, into the ring.
        ``consumer(batch_view, first_patch, n)`` stands for the separator; ``events`` (a list) receives
        (start, end) CUDA event pairs around each shift-stack launch for kernel-level timing."""
        B, M, T = mix_dev.shape
        N = shifts_dev.shape[0]
        bufs = self._ring(M, T)
        launches = 0
        for k, i in enumerate(range(0, N, self.net_batch)):
            n = min(self.net_batch, N - i)
            buf = bufs[k % len(bufs)][:self.net_batch]
            if events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if fused_norm:
                out = native.shift_stack_norm(mix_dev, shifts_dev[i:i + n], mix_index_dev[i:i + n], out=buf)
                launches += 2
            else:
                out = native.shift_stack(mix_dev, shifts_dev[i:i + n], mix_index_dev[i:i + n], out=buf)
                launches += 1
            if events is not None:
                e1.record()
                events.append((e0, e1, n))
            if consumer is not None:
                consumer(out, i, n)
        return launches

    # ---- pruning ---------------------------------------------------------------------------------
    def find_peaks(self, map_dev):
        """Device peak picking for the whole batch -> per mixture (ids list, values array, MAX_POWER)."""
        peaks, count, mx = self.node.native_peaks.find(map_dev)
        vals = torch.gather(map_dev, 1, peaks.clamp(min=0).to(torch.int64))
        peaks_h, count_h, mx_h, vals_h = peaks.cpu().numpy(), count.cpu().numpy(), mx.cpu().numpy(), vals.cpu().numpy()
        out = []
        for b in range(map_dev.shape[0]):
            n = min(int(count_h[b]), peaks_h.shape[1])
            out.append(([int(i) for i in peaks_h[b, :n]], vals_h[b, :n], float(mx_h[b])))
        return out

    def prune(self, map_dev):
        """Apply_SRP_PHAT's pruning for every mixture of the batch, on the device (peak picking + greedy
        hypercube selection, SRP_Prunning.py:500-643) -> list (per mixture) of list[Patch]."""
        n, off, wid, pk = self.select(map_dev)
        n_h, off_h, wid_h, pk_h = n.cpu().numpy(), off.cpu().numpy(), wid.cpu().numpy(), pk.cpu().numpy()
        if n_h.max(initial=0) > off_h.shape[1] or self._last_peak_count.max().item() > self.node.native_peaks.max_peaks:
            raise native._lib.AswError("peak or patch list capacity exceeded: results would be truncated "
                                       "(raise max_peaks / max_patches)")
        return [self.node.patches_from_device(min(int(n_h[b]), off_h.shape[1]), off_h[b], wid_h[b], pk_h[b])
                for b in range(map_dev.shape[0])]

    def prune_host_greedy(self, map_dev):
        """Same with the greedy selection on the host (device peaks): the cross-check of `prune`."""
        return [self.node.local_source_adaptive(ids, vals) for ids, vals, _ in self.find_peaks(map_dev)]

    def select(self, map_dev):
        """Device peak picking + greedy selection -> (n (B,), offsets (B, P, D), widths (B, P), peak ids (B, P))."""
        peaks, count, _ = self.node.native_peaks.find(map_dev)
        self._last_peak_count = count
        return self.node.native_select.select(map_dev, peaks, count)

    def shift_table(self, n, offsets, capacity):
        """Dense device shift table of all selected patches of the batch (no host round trip)."""
        return native.build_shift_table(n, offsets, capacity)

    def stack_counted(self, mix_dev, shifts, mix_index, n_total, capacity, consumer=None, events=None):
        """`stack` driven by a device-resident patch count: launches ceil(capacity / net_batch) batches,
        rows beyond n_total are skipped on the device."""
        B, M, T = mix_dev.shape
        bufs = self._ring(M, T)
        rows = self.net_batch * self.launch_batches
        for k, i in enumerate(range(0, capacity, rows)):
            n = min(rows, capacity - i)
            buf = bufs[k % len(bufs)]
            if events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            native.shift_stack_counted(mix_dev, shifts, mix_index, n_total, i, n, buf)
            if events is not None:
                e1.record()
                events.append((e0, e1, n))
            if consumer is not None:
                consumer(buf, i, n)

    # ---- fine stage --------------------------------------------------------------------------------
    def upper_bound_pairwise(self):
        """sep/Mic_Array.py:113-115: largest physical TDoA (+ 8 cm) of every mic against mic 0, in samples."""
        mic = self.node.mic_pos
        return (np.linalg.norm(mic[1:] - mic[:1], axis=1) + 0.08) / self.node.C * self.node.FS

    def fine_table(self, n_sel, offsets, widths, capacity, max_leaves=128):
        """Spotform_Small_Patch_Parallel's patch-list assembly (sep/Mic_Array.py:244-262) for a whole batch on the
        device: every selected coarse patch of every mixture is a candidate, all of them are subdivided in one
        asw_subdivide launch (search_area, local_utils_3d.py:212-335) and their leaves + centre patches become one
        dense shift table.  ``n_sel`` (B,), ``offsets`` (B, P, D), ``widths`` (B, P) as returned by ``select``.
        -> (shifts (capacity, M), mix_index, cand_index, cand_start (B * P + 1,), n_total (1,), status (B * P,),
        leaf_count (B * P,)); nothing is copied to the host."""
        B, P, D = offsets.shape
        dev = offsets.device
        slot = torch.arange(P, device=dev, dtype=torch.int32)
        valid = slot[None, :] < n_sel.clamp(max=P)[:, None]
        w = (widths * valid).reshape(-1).to(torch.int32).contiguous()          # width 0 = empty slot
        owner = torch.arange(B, device=dev, dtype=torch.int32)[:, None].expand(B, P).reshape(-1).contiguous()
        cnt, off, _, root, status = native.subdivide_device(self.node.native_select, offsets.reshape(-1, D).contiguous(),
                                                            w, self.upper_bound_pairwise(), max_leaves=max_leaves)
        return native.build_fine_table(cnt, off, root, w, owner, capacity) + (status, cnt)

    def stack_norm_counted(self, mix_dev, shifts, mix_index, n_total, n_rows, tables=None, max_lag=0, consumer=None,
                           events=None, grouped=False):
        """Fused shift-stack + normalize_input of rows [0, n_rows) of a device-built table, ``net_batch`` patches per
        launch into the ring; rows >= n_total[0] are skipped on the device.  ``tables``: CorrTables.compute(mix_dev)
        (many patches per mixture: the fine stage); ``grouped``: statistics from one tiled pass per mixture (a few dozen
        patches per mixture: the coarse stage)."""
        B, M, T = mix_dev.shape
        bufs = self._ring(M, T)
        rows = self.net_batch * self.launch_batches
        for k, i in enumerate(range(0, n_rows, rows)):
            n = min(rows, n_rows - i)
            buf = bufs[k % len(bufs)]
            if events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            out, mu, sd = native.shift_stack_norm(mix_dev, shifts, mix_index, out=buf, tables=tables, max_lag=max_lag,
                                                  n_total=n_total, n_base=i, N=n, grouped=grouped)
            if events is not None:
                e1.record()
                events.append((e0, e1, n))
            if consumer is not None:
                consumer(out, mu, sd, i, n)

    # ---- host helpers ----------------------------------------------------------------------------
    def prune_host(self, srp_map_host):
        """The reference's pruning (SRP_Prunning.py:347-357, 500-643) on one mixture's map -> list[Patch]."""
        self.node.load_map(srp_map_host)
        return self.node.local_source_adaptive()

    @staticmethod
    def patch_table(patch_lists):
        """list (per mixture) of list[Patch] -> (shifts (N, M) int32, mix_index (N,) int32) numpy."""
        offs, mi = [], []
        for b, pl in enumerate(patch_lists):
            for p in pl:
                offs.append(p.sample_offset)
                mi.append(b)
        if not offs:
            return np.zeros((0, 1), dtype=np.int32), np.zeros((0,), dtype=np.int32)
        return native.offsets_to_shifts(np.stack(offs)), np.asarray(mi, dtype=np.int32)
