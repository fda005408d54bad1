"""Thin torch-facing wrappers over the C ABI (include/asw.h).

torch is used for device memory and streams only; every computation below is a
libasw.so kernel.  All tensors handed to the library must live on the handle's
CUDA device; anything else raises (no CPU fallback)."""
import ctypes

import numpy as np
import torch

from . import _lib
from .constants import HOP, PHAT_TOL, n_fft


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(device):
    """The calling thread's current CUDA stream on ``device`` as a raw pointer (what torch itself launches on)."""
    if _raw_stream is not None:           # no Stream object per call: 0.3 us instead of 13 us, on every entry point
        dev = torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        return ctypes.c_void_p(_raw_stream(idx))
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t, name, dtype):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.AswError(f"{name} must be a CUDA tensor (this path has no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.AswError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.AswError(f"{name} must be contiguous")


def pair_list(M):
    return [(i, j) for i in range(M) for j in range(i + 1, M)]


def pair_lags(grids, mic_pos, fs, C):
    """Fractional pair lags in samples, (G, P) float64: the phase slope of the reference's steering
    table (sep/Traditional_SP/SRP_Prunning.py:368-381, :228).  Mic height is ignored (:375)."""
    grids = np.asarray(grids, dtype=np.float64)
    mic_pos = np.asarray(mic_pos, dtype=np.float64)
    d = np.sqrt((grids[None, :, 0] - mic_pos[:, None, 0]) ** 2 + (grids[None, :, 1] - mic_pos[:, None, 1]) ** 2
                + grids[None, :, 2] ** 2)
    prs = pair_list(mic_pos.shape[0])
    return np.ascontiguousarray(np.stack([fs * (d[i] - d[j]) / C for (i, j) in prs], axis=1))


class NativeSRP:
    """Device-resident scoring handle (asw_srp_t) for one geometry."""

    def __init__(self, lag_samples, num_mic, device=None, bin0=2, bin1=200, tol=PHAT_TOL, oversample=0,
                 pad_tail=False):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.AswError("no CUDA device: the SRP-PHAT path is CUDA-only (sm_100a), there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        lag = np.ascontiguousarray(lag_samples, dtype=np.float64)
        self.M = int(num_mic)
        self.P = self.M * (self.M - 1) // 2
        if lag.ndim != 2 or lag.shape[1] != self.P:
            raise _lib.AswError(f"lag_samples must be (G, {self.P})")
        self.G = lag.shape[0]
        self.F = bin1 - bin0
        self._h = ctypes.c_void_p()
        _lib.check(self.lib.asw_srp_create(ctypes.byref(self._h), self.device.index or 0, self.M, self.G,
                                           lag.ctypes.data_as(ctypes.c_void_p), n_fft, HOP, bin0, bin1,
                                           ctypes.c_float(tol), oversample))
        self._last = (0, 0)
        self.pad_tail = False
        if pad_tail:
            self.set_pad_tail(True)

    def set_pad_tail(self, enabled):
        """STFT framing of an analysis window: False = floor((win - nfft) / hop) + 1 frames (assumption A1),
        True = ceil(...) + 1 frames with the ragged last frame zero padded (see include/asw.h)."""
        _lib.check(self.lib.asw_srp_set_frame_mode(self._h, 1 if enabled else 0))
        self.pad_tail = bool(enabled)

    def set_stft_path(self, path):
        """"auto" (fused kernel for <= 8 mics, split beyond), "split" or "generic": see asw_srp_set_stft_path."""
        _lib.check(self.lib.asw_srp_set_stft_path(self._h, {"auto": 0, "split": 1, "generic": 2}[path]))

    def num_frames(self, win_len):
        return int(self.lib.asw_srp_num_frames_mode(int(win_len), n_fft, HOP, 1 if self.pad_tail else 0))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.asw_srp_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def num_windows(self, T, win_len):
        return int(self.lib.asw_srp_num_windows(T, win_len))

    def score(self, mix, win_len, out=None):
        """mix (B, M, T) or (M, T) float32 CUDA -> map (B, G) float32 CUDA."""
        if mix.dim() == 2:
            mix = mix.unsqueeze(0)
        _require_cuda(mix, "mix", torch.float32)
        B, M, T = mix.shape
        if M != self.M:
            raise _lib.AswError(f"mix has {M} channels, handle was built for {self.M}")
        if out is None:
            out = torch.empty((B, self.G), device=mix.device, dtype=torch.float32)
        else:
            _require_cuda(out, "out", torch.float32)
        _lib.check(self.lib.asw_srp_score(self._h, _ptr(mix), B, T, int(win_len), _ptr(out), _stream(mix.device)))
        self._last = (B, self.num_windows(T, win_len))
        return out

    def table_len(self):
        return self.gcc_layout()[3]

    def gcc(self, mix, win_len, out=None):
        """Transform stage only (asw_srp_gcc): mix (B, M, T) float32 CUDA -> GCC lag tables (B, Nw * table_len)."""
        _require_cuda(mix, "mix", torch.float32)
        B, M, T = mix.shape
        if M != self.M:
            raise _lib.AswError(f"mix has {M} channels, handle was built for {self.M}")
        Nw = self.num_windows(T, win_len)
        if out is None:
            out = torch.empty((B, Nw * self.table_len()), device=mix.device, dtype=torch.float32)
        else:
            _require_cuda(out, "out", torch.float32)
        if B:
            _lib.check(self.lib.asw_srp_gcc(self._h, _ptr(mix), B, T, int(win_len), _ptr(out), _stream(mix.device)))
        return out

    def gather(self, gcc, Nw, out=None):
        """Gather stage only (asw_srp_gather): tables (B, Nw * table_len) -> map (B, G) of this handle's hypercubes."""
        _require_cuda(gcc, "gcc", torch.float32)
        B = gcc.shape[0]
        if gcc.shape[1] != Nw * self.table_len():
            raise _lib.AswError("gcc must be (B, Nw * table_len)")
        if out is None:
            out = torch.empty((B, self.G), device=gcc.device, dtype=torch.float32)
        else:
            _require_cuda(out, "out", torch.float32)
        _lib.check(self.lib.asw_srp_gather(self._h, _ptr(gcc), B, int(Nw), _ptr(out), _stream(gcc.device)))
        return out

    def read_cc(self):
        """CC_flat of the last score call: (B, Nw, F, P) complex64."""
        B, Nw = self._last
        buf = torch.empty((B, Nw, self.F, self.P, 2), device=self.device, dtype=torch.float32)
        _lib.check(self.lib.asw_srp_read_cc(self._h, _ptr(buf), _stream(self.device)))
        return torch.view_as_complex(buf)

    def gcc_layout(self):
        lo = (ctypes.c_int * self.P)()
        n = (ctypes.c_int * self.P)()
        off = (ctypes.c_int * self.P)()
        tl, U = ctypes.c_int(), ctypes.c_int()
        _lib.check(self.lib.asw_srp_gcc_layout(self._h, lo, n, off, ctypes.byref(tl), ctypes.byref(U)))
        return np.array(lo), np.array(n), np.array(off), tl.value, U.value

    def read_gcc(self):
        """GCC lag tables of the last score call: flat (B, Nw * table_len) float32 (pair-major)."""
        B, Nw = self._last
        _, _, _, tl, _ = self.gcc_layout()
        buf = torch.empty((B, Nw * tl), device=self.device, dtype=torch.float32)
        _lib.check(self.lib.asw_srp_read_gcc(self._h, _ptr(buf), _stream(self.device)))
        return buf


class NativePeaks:
    """Device peak picking (asw_peaks_t): fill_powermap_torch + MAX_POWER + find_valid_peak_new
    (sep/Traditional_SP/SRP_Prunning.py:347-357, :432, :500-544) for a batch of maps."""

    def __init__(self, power_index, member_mask, dis_matrix, n_grids, thresholds, ratio=4.0, device=None,
                 max_peaks=1024):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.AswError("no CUDA device: peak picking on the device needs one (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        idx = np.where(member_mask, power_index, -1).astype(np.int32)
        idx = np.ascontiguousarray(idx)
        dis = np.ascontiguousarray(dis_matrix, dtype=np.float64)
        thr = np.ascontiguousarray(thresholds, dtype=np.float64)
        self.G = int(n_grids)
        self.max_peaks = int(max_peaks)
        Lx, Ly, Lz = idx.shape
        self._h = ctypes.c_void_p()
        _lib.check(self.lib.asw_peaks_create(ctypes.byref(self._h), self.device.index or 0, Lx, Ly, Lz, self.G,
                                             idx.ctypes.data, dis.ctypes.data, thr.ctypes.data,
                                             ctypes.c_double(ratio)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.asw_peaks_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def find(self, srp_map):
        """(B, G) float32 CUDA -> (peaks (B, max_peaks) int32 padded with -1, count (B,) int32, MAX_POWER (B,))."""
        if srp_map.dim() == 1:
            srp_map = srp_map.unsqueeze(0)
        _require_cuda(srp_map, "srp_map", torch.float32)
        B, G = srp_map.shape
        if G != self.G:
            raise _lib.AswError(f"map has {G} hypercubes, handle was built for {self.G}")
        peaks = torch.empty((B, self.max_peaks), device=srp_map.device, dtype=torch.int32)
        count = torch.empty((B,), device=srp_map.device, dtype=torch.int32)
        mx = torch.empty((B,), device=srp_map.device, dtype=torch.float32)
        _lib.check(self.lib.asw_peaks_find(self._h, _ptr(srp_map), B, _ptr(peaks), self.max_peaks, _ptr(count),
                                           _ptr(mx), _stream(srp_map.device)))
        return peaks, count, mx


class NativeSelect:
    """Device greedy selection of coarse patches (asw_select_t): SRP_PHAT.local_source_adaptive
    (sep/Traditional_SP/SRP_Prunning.py:547-643) for a batch of mixtures."""

    def __init__(self, cluster_offsets, Offset_5, Offset_1, Range_spk, Axis_range, width, device=None,
                 max_patches=64):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.AswError("no CUDA device: patch selection on the device needs one (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        cl = np.ascontiguousarray(cluster_offsets, dtype=np.int32)
        self.G, self.D = cl.shape
        Ny5, Nx5, Nz, D = Offset_5.shape
        Ny1, Nx1, Nz1, _ = Offset_1.shape
        assert Nz1 == Nz and D == self.D
        o5 = Offset_5.reshape(-1, D)
        # bucket grid over the integer parts of the first two TDoA coordinates
        f0 = np.floor(o5[:, 0]).astype(np.int64)
        f1 = np.floor(o5[:, 1]).astype(np.int64) if D >= 2 else np.zeros(o5.shape[0], dtype=np.int64)
        b0min, b1min = int(f0.min()), int(f1.min())
        NB0, NB1 = int(f0.max()) - b0min + 1, int(f1.max()) - b1min + 1
        key = (f0 - b0min) * NB1 + (f1 - b1min)
        order = np.argsort(key, kind="stable")
        bucket_start = np.ascontiguousarray(
            np.searchsorted(key[order], np.arange(NB0 * NB1 + 1), side="left").astype(np.int32))
        iy, ix, iz = np.unravel_index(np.arange(o5.shape[0]), (Ny5, Nx5, Nz))
        ok = (5 * iy < Ny1) & (5 * ix < Nx1)
        o1 = np.full(o5.shape, np.nan)
        o1[ok] = Offset_1[5 * iy[ok], 5 * ix[ok], iz[ok]]
        off5_sorted = np.ascontiguousarray(o5[order])
        off1_at5 = np.ascontiguousarray(o1[order])
        vox5 = np.ascontiguousarray((iy * Nx5 + ix)[order].astype(np.int32))
        r = Range_spk
        xx5 = np.ascontiguousarray(np.arange(r[0], r[1], 0.05))
        yy5 = np.ascontiguousarray(np.arange(r[2], r[3], 0.05))
        assert xx5.shape[0] == Nx5 and yy5.shape[0] == Ny5
        axis = np.array([Axis_range[0][0], Axis_range[0][1], Axis_range[1][0], Axis_range[1][1]], dtype=np.float64)
        off1 = np.ascontiguousarray(Offset_1, dtype=np.float64)
        self.max_patches = int(max_patches)
        self._h = ctypes.c_void_p()
        _lib.check(self.lib.asw_select_create(ctypes.byref(self._h), self.device.index or 0, self.G, self.D, int(width),
                                              cl.ctypes.data, off5_sorted.ctypes.data, off1_at5.ctypes.data,
                                              vox5.ctypes.data, bucket_start.ctypes.data, b0min, b1min, NB0, NB1,
                                              o5.shape[0], Nx5, Ny5, xx5.ctypes.data,
                                              yy5.ctypes.data, axis.ctypes.data, off1.ctypes.data, Ny1, Nx1, Nz))
        # coordinates of the 1 cm volume (SRP_Prunning.py:158-160), for the leaf centres of asw_subdivide
        xx1 = np.ascontiguousarray(np.arange(r[0], r[1], 0.01))
        yy1 = np.ascontiguousarray(np.arange(r[2], r[3], 0.01))
        zz = np.ascontiguousarray(np.arange(r[4], r[5], 0.1))
        assert xx1.shape[0] == Nx1 and yy1.shape[0] == Ny1 and zz.shape[0] == Nz
        _lib.check(self.lib.asw_select_set_grid1(self._h, xx1.ctypes.data, yy1.ctypes.data, zz.ctypes.data))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.asw_select_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def select(self, srp_map, peaks, count):
        """-> (n_patches (B,), offsets (B, max_patches, D), widths (B, max_patches), peak ids (B, max_patches)), int32 CUDA."""
        _require_cuda(srp_map, "srp_map", torch.float32)
        _require_cuda(peaks, "peaks", torch.int32)
        _require_cuda(count, "count", torch.int32)
        B = srp_map.shape[0]
        dev = srp_map.device
        n = torch.empty((B,), device=dev, dtype=torch.int32)
        off = torch.zeros((B, self.max_patches, self.D), device=dev, dtype=torch.int32)
        wid = torch.zeros((B, self.max_patches), device=dev, dtype=torch.int32)
        pk = torch.full((B, self.max_patches), -1, device=dev, dtype=torch.int32)
        _lib.check(self.lib.asw_select_patches(self._h, _ptr(srp_map), _ptr(peaks), peaks.shape[1], _ptr(count), B,
                                               _ptr(n), _ptr(off), _ptr(wid), _ptr(pk), self.max_patches,
                                               _stream(dev)))
        return n, off, wid, pk


def subdivide(select_handle, centres, widths, upper_bound, max_leaves=128, member_cap=0):
    """Device search_area for n coarse patches.  centres (n, D) int32 CUDA, widths (n,) int32 CUDA,
    upper_bound (D,) float64 host.  Returns host numpy arrays
    (leaf_count (n,), leaf_off (n, L, D), leaf_w (n, L, D), leaf_npts (n, L), leaf_box (n, L, 2, D),
    root_after (n, 2, D), leaf_centre (n, L, 3)) and, with ``member_cap > 0``, a list of n int64 arrays: every
    candidate's member voxels (flat indices into the 1 cm volume) sorted ascending = the order of its area_points."""
    _require_cuda(centres, "centres", torch.int32)
    _require_cuda(widths, "widths", torch.int32)
    n, D = centres.shape
    dev = centres.device
    ub = np.ascontiguousarray(upper_bound, dtype=np.float64)
    if ub.shape != (D,):
        raise _lib.AswError(f"upper_bound must have {D} entries")
    cnt = torch.zeros((n,), device=dev, dtype=torch.int32)
    off = torch.zeros((n, max_leaves, D), device=dev, dtype=torch.int32)
    wid = torch.zeros((n, max_leaves, D), device=dev, dtype=torch.int32)
    npts = torch.zeros((n, max_leaves), device=dev, dtype=torch.int32)
    box = torch.zeros((n, max_leaves, 2, D), device=dev, dtype=torch.float64)
    centre = torch.zeros((n, max_leaves, 3), device=dev, dtype=torch.float64)
    root = torch.zeros((n, 2, D), device=dev, dtype=torch.int32)
    status = torch.zeros((n,), device=dev, dtype=torch.int32)
    members = rcount = None
    if member_cap > 0:
        members = torch.full((n, int(member_cap)), 2 ** 31 - 1, device=dev, dtype=torch.int32)
        rcount = torch.zeros((n,), device=dev, dtype=torch.int32)
    if n:
        _lib.check(select_handle.lib.asw_subdivide(select_handle._h, _ptr(centres), _ptr(widths), n, ub.ctypes.data,
                                                   int(max_leaves), _ptr(cnt), _ptr(off), _ptr(wid), _ptr(npts),
                                                   _ptr(box), _ptr(centre), _ptr(root), _ptr(status),
                                                   _ptr(members) if members is not None else None, int(member_cap),
                                                   _ptr(rcount) if rcount is not None else None, _stream(dev)))
    st = status.cpu().numpy()
    cn = cnt.cpu().numpy()
    if (st != 0).any() or (cn > max_leaves).any():
        raise _lib.AswCapacityError(f"asw_subdivide: capacity exceeded (status codes {sorted(set(st.tolist()))}: 1 member list, "
                                    f"2 nodes per level, 3 leaves; most leaves {int(cn.max())} of {max_leaves})")
    out = (cn, off.cpu().numpy(), wid.cpu().numpy(), npts.cpu().numpy(), box.cpu().numpy(), root.cpu().numpy(),
           centre.cpu().numpy())
    if members is None:
        return out
    rc = rcount.cpu().numpy()
    if (rc > member_cap).any():
        raise _lib.AswError(f"asw_subdivide: member list capacity {member_cap} exceeded ({int(rc.max())} voxels)")
    width = int(rc.max(initial=0))
    srt = torch.sort(members[:, :max(width, 1)], dim=1).values.cpu().numpy() if n else np.zeros((0, 1), dtype=np.int32)
    return out + ([srt[i, :rc[i]].astype(np.int64) for i in range(n)],)


def subdivide_device(select_handle, centres, widths, upper_bound, max_leaves=128):
    """``subdivide`` without the host round trip: device tensors (leaf_count (n,), leaf_off (n, L, D), leaf_w (n, L, D),
    root_after (n, 2, D), status (n,)).  Candidates with ``widths <= 0`` are empty slots (leaf_count 0).  The caller
    checks ``status`` / ``leaf_count <= max_leaves`` when it next synchronises."""
    _require_cuda(centres, "centres", torch.int32)
    _require_cuda(widths, "widths", torch.int32)
    n, D = centres.shape
    dev = centres.device
    ub = np.ascontiguousarray(upper_bound, dtype=np.float64)
    if ub.shape != (D,):
        raise _lib.AswError(f"upper_bound must have {D} entries")
    cnt = torch.zeros((n,), device=dev, dtype=torch.int32)
    off = torch.empty((n, max_leaves, D), device=dev, dtype=torch.int32)
    wid = torch.empty((n, max_leaves, D), device=dev, dtype=torch.int32)
    npts = torch.empty((n, max_leaves), device=dev, dtype=torch.int32)
    box = torch.empty((n, max_leaves, 2, D), device=dev, dtype=torch.float64)
    root = torch.zeros((n, 2, D), device=dev, dtype=torch.int32)
    status = torch.zeros((n,), device=dev, dtype=torch.int32)
    if n:
        _lib.check(select_handle.lib.asw_subdivide(select_handle._h, _ptr(centres), _ptr(widths), n, ub.ctypes.data,
                                                   int(max_leaves), _ptr(cnt), _ptr(off), _ptr(wid), _ptr(npts),
                                                   _ptr(box), None, _ptr(root), _ptr(status), None, 0, None,
                                                   _stream(dev)))
    return cnt, off, wid, root, status


def build_fine_table(leaf_count, leaf_off, root_after, widths, owner, capacity):
    """asw_build_fine_table: -> (shifts (capacity, D + 1), mix_index (capacity,), cand_index (capacity,),
    cand_start (n + 1,), n_total (1,)), all int32 CUDA, no host synchronisation."""
    for t, name in ((leaf_count, "leaf_count"), (leaf_off, "leaf_off"), (root_after, "root_after"), (widths, "widths")):
        _require_cuda(t, name, torch.int32)
    if owner is not None:
        _require_cuda(owner, "owner", torch.int32)
    n, L, D = leaf_off.shape
    dev = leaf_off.device
    shifts = torch.zeros((capacity, D + 1), device=dev, dtype=torch.int32)
    mix_index = torch.zeros((capacity,), device=dev, dtype=torch.int32)
    cand_index = torch.zeros((capacity,), device=dev, dtype=torch.int32)
    cand_start = torch.zeros((n + 1,), device=dev, dtype=torch.int32)
    n_total = torch.zeros((1,), device=dev, dtype=torch.int32)
    if n:
        with torch.cuda.device(dev):
            _lib.check(_lib.load().asw_build_fine_table(_ptr(leaf_count), _ptr(leaf_off), _ptr(root_after), _ptr(widths),
                                                        _ptr(owner) if owner is not None else None, n, L, D,
                                                        _ptr(shifts), _ptr(mix_index), _ptr(cand_index),
                                                        _ptr(cand_start), _ptr(n_total), int(capacity), _stream(dev)))
    return shifts, mix_index, cand_index, cand_start, n_total


def build_shift_table(n_patches, offsets, capacity, shifts=None, mix_index=None, n_total=None):
    """Per-mixture patch lists -> dense (capacity, D + 1) int32 shift table, (capacity,) mixture index and the
    device-resident total, without leaving the device."""
    _require_cuda(n_patches, "n_patches", torch.int32)
    _require_cuda(offsets, "offsets", torch.int32)
    B, max_patches, D = offsets.shape
    dev = offsets.device
    if shifts is None:
        shifts = torch.zeros((capacity, D + 1), device=dev, dtype=torch.int32)
    if mix_index is None:
        mix_index = torch.zeros((capacity,), device=dev, dtype=torch.int32)
    if n_total is None:
        n_total = torch.zeros((1,), device=dev, dtype=torch.int32)
    with torch.cuda.device(offsets.device):     # handle-less entry points launch on the current device
        _lib.check(_lib.load().asw_build_shift_table(_ptr(n_patches), _ptr(offsets), B, max_patches, D, _ptr(shifts),
                                                     _ptr(mix_index), _ptr(n_total), int(capacity), _stream(dev)))
    return shifts, mix_index, n_total


def shift_stack_counted(mix, shifts, mix_index, n_total, n_base, N, out):
    """shift_stack for rows [n_base, n_base + N) of a device-built table; rows >= n_total[0] are skipped."""
    if mix.dim() == 2:
        mix = mix.unsqueeze(0)
    _require_cuda(mix, "mix", torch.float32)
    _require_cuda(shifts, "shifts", torch.int32)
    _require_cuda(mix_index, "mix_index", torch.int32)
    _require_cuda(n_total, "n_total", torch.int32)
    _require_cuda(out, "out", torch.float32)
    B, M, T = mix.shape
    if shifts.shape[1] != M or out.numel() < N * M * T:
        raise _lib.AswError("shift table / output shape mismatch")
    if n_base < 0 or n_base + N > shifts.shape[0] or n_base + N > mix_index.shape[0]:
        raise _lib.AswError(f"rows [{n_base}, {n_base + N}) exceed the shift table's capacity {shifts.shape[0]}")
    with torch.cuda.device(mix.device):     # handle-less entry points launch on the current device
        _lib.check(_lib.load().asw_shift_stack_counted(_ptr(mix), _ptr(shifts), _ptr(mix_index), _ptr(n_total), int(n_base),
                                                       int(N), B, M, T, _ptr(out), _stream(mix.device)))
    return out


def map_topk(srp_map, K, idx_offset=0):
    """(B, G) float32 CUDA -> values (B, K) float32, indices (B, K) int32 (descending, ties to lower index)."""
    if srp_map.dim() == 1:
        srp_map = srp_map.unsqueeze(0)
    _require_cuda(srp_map, "srp_map", torch.float32)
    B, G = srp_map.shape
    val = torch.empty((B, K), device=srp_map.device, dtype=torch.float32)
    idx = torch.empty((B, K), device=srp_map.device, dtype=torch.int32)
    with torch.cuda.device(srp_map.device):     # handle-less entry points launch on the current device
        _lib.check(_lib.load().asw_map_topk(_ptr(srp_map), B, G, int(K), int(idx_offset), _ptr(val), _ptr(idx),
                                            _stream(srp_map.device)))
    return val, idx


def shift_stack(mix, shifts, mix_index=None, out=None):
    """out[n, c, t] = mix[mix_index[n], c, (t + shifts[n, c]) mod T]  (network.py:12-25, 75-83).

    mix (B, M, T) or (M, T) float32 CUDA; shifts (N, M) int32 CUDA; mix_index (N,) int32 CUDA or None."""
    if mix.dim() == 2:
        mix = mix.unsqueeze(0)
    _require_cuda(mix, "mix", torch.float32)
    _require_cuda(shifts, "shifts", torch.int32)
    B, M, T = mix.shape
    N = shifts.shape[0]
    if shifts.shape != (N, M):
        raise _lib.AswError(f"shifts must be (N, {M})")
    if mix_index is not None:
        _require_cuda(mix_index, "mix_index", torch.int32)
    if out is None:
        out = torch.empty((N, M, T), device=mix.device, dtype=torch.float32)
    else:
        _require_cuda(out, "out", torch.float32)
        if out.numel() < N * M * T:
            raise _lib.AswError("out is too small")
    if N == 0:
        return out[:0]
    with torch.cuda.device(mix.device):     # handle-less entry points launch on the current device
        _lib.check(_lib.load().asw_shift_stack(_ptr(mix), _ptr(shifts), _ptr(mix_index) if mix_index is not None else None,
                                               N, B, M, T, _ptr(out), _stream(mix.device)))
    return out[:N]


class CorrTables:
    """Per-mixture correlation tables (asw_corr_t) that let ``shift_stack_norm`` find a patch's mean / std from
    M + M + P look-ups instead of a pass over its samples (see include/asw.h).  ``max_lag``: largest
    |r_c' - r_c| the tables cover; patches beyond it fall back to the exact pass on the device."""

    MIN_T = 4096

    def __init__(self, num_mic, device=None, max_lag=512):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.AswError("no CUDA device: the correlation tables are CUDA-only (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.M, self.max_lag = int(num_mic), int(max_lag)
        self._h = ctypes.c_void_p()
        _lib.check(self.lib.asw_corr_create(ctypes.byref(self._h), self.device.index or 0, self.M, self.max_lag))
        self.table_len = int(self.lib.asw_corr_table_len(self._h))

    @staticmethod
    def lag_for_geometry(mic_positions, fs, C=343.0, width=8):
        """Smallest table range that covers every patch of an array: the largest pair TDoA the geometry can produce
        plus the hypercube width, rounded up to a multiple of 32 samples (the block length 2048 - 2 max_lag grows as
        the range shrinks: 141 blocks of a 3 s mixture at 512, 103 at 320)."""
        mic = np.asarray(mic_positions, dtype=np.float64)
        d = np.sqrt(((mic[:, None, :] - mic[None, :, :]) ** 2).sum(-1)).max()
        need = int(np.ceil(d / C * fs)) + int(width) + 2
        return int(min(512, max(32, -(-need // 32) * 32)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.asw_corr_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compute(self, mix, out=None):
        """mix (B, M, T) or (M, T) float32 CUDA -> tables (B, table_len) float64 CUDA."""
        if mix.dim() == 2:
            mix = mix.unsqueeze(0)
        _require_cuda(mix, "mix", torch.float32)
        B, M, T = mix.shape
        if M != self.M:
            raise _lib.AswError(f"mix has {M} channels, handle was built for {self.M}")
        if out is None:
            out = torch.empty((B, self.table_len), device=mix.device, dtype=torch.float64)
        else:
            _require_cuda(out, "out", torch.float64)
            if out.numel() < B * self.table_len:
                raise _lib.AswError("out is too small")
        _lib.check(self.lib.asw_corr_tables(self._h, _ptr(mix), B, T, _ptr(out), _stream(mix.device)))
        return out

    def pair_table(self, tables, b, i, j):
        """R_ij(l), l = -max_lag..max_lag, of mixture b (i < j) as a (2 max_lag + 1,) view."""
        p = i * self.M - i * (i + 1) // 2 + (j - i - 1)
        n = 2 * self.max_lag + 1
        return tables[b, 2 * self.M + p * n: 2 * self.M + (p + 1) * n]


def shift_stack_norm(mix, shifts, mix_index=None, out=None, tables=None, max_lag=0, n_total=None, n_base=0, N=None,
                     grouped=False):
    """shift_stack fused with normalize_input (SpeakerLocalization/network.py:28-40).
    Returns (data_norm (N, M, T), means (N, 1, 1), stds (N, 1, 1)).
    ``tables`` (B, table_len) float64 from ``CorrTables.compute`` (+ its ``max_lag``): the statistics come from the
    per-mixture tables (asw_shift_stack_norm_tab) instead of a pass over every patch.
    ``n_total`` (1,) int32 CUDA: rows [n_base, n_base + N) of a device-built table, rows >= n_total[0] skipped.
    ``grouped``: the rows are grouped by mixture (``mix_index`` non-decreasing, as the device-built tables are) and no
    ``tables`` are given: one pass over each mixture's audio serves all of its patches (asw_shift_stack_norm_grouped,
    exact integer statistics) -- the choice for a few dozen patches per mixture."""
    if mix.dim() == 2:
        mix = mix.unsqueeze(0)
    _require_cuda(mix, "mix", torch.float32)
    _require_cuda(shifts, "shifts", torch.int32)
    B, M, T = mix.shape
    if shifts.dim() != 2 or shifts.shape[1] != M:
        raise _lib.AswError(f"shifts must be (N, {M})")
    if N is None:
        N = shifts.shape[0] - n_base
    if n_base < 0 or N < 0 or n_base + N > shifts.shape[0]:
        raise _lib.AswError(f"rows [{n_base}, {n_base + N}) exceed the shift table's {shifts.shape[0]} rows")
    if mix_index is not None:
        _require_cuda(mix_index, "mix_index", torch.int32)
        if mix_index.shape[0] < n_base + N:
            raise _lib.AswError("mix_index is shorter than the shift table")
    if n_total is not None:
        _require_cuda(n_total, "n_total", torch.int32)
    if out is None:
        out = torch.empty((N, M, T), device=mix.device, dtype=torch.float32)
    else:
        _require_cuda(out, "out", torch.float32)
        if out.numel() < N * M * T:
            raise _lib.AswError("out is too small")
    means = torch.empty((N,), device=mix.device, dtype=torch.float32)
    stds = torch.empty((N,), device=mix.device, dtype=torch.float32)
    work = torch.empty((N, 2), device=mix.device, dtype=torch.float64)
    if N == 0:
        return out[:0], means.view(0, 1, 1), stds.view(0, 1, 1)
    mip = _ptr(mix_index) if mix_index is not None else None
    with torch.cuda.device(mix.device):     # handle-less entry points launch on the current device
        if grouped and tables is None:
            if mix_index is None:
                mix_index = torch.zeros((n_base + N,), device=mix.device, dtype=torch.int32)
                mip = _ptr(mix_index)
            ranges = torch.empty((B + 1,), device=mix.device, dtype=torch.int32)
            _lib.check(_lib.load().asw_shift_stack_norm_grouped(
                _ptr(mix), _ptr(shifts), mip, N, B, M, T, _ptr(out), _ptr(means), _ptr(stds), _ptr(work), _ptr(ranges),
                _ptr(n_total) if n_total is not None else None, int(n_base), _stream(mix.device)))
        elif tables is not None or n_total is not None or n_base:
            tl = 0
            if tables is not None:
                _require_cuda(tables, "tables", torch.float64)
                if tables.dim() != 2 or tables.shape[0] != B:
                    raise _lib.AswError("tables must be (B, table_len)")
                tl = int(tables.shape[1])
            _lib.check(_lib.load().asw_shift_stack_norm_tab(
                _ptr(mix), _ptr(shifts), mip, N, B, M, T, _ptr(tables) if tables is not None else None, tl, int(max_lag),
                _ptr(out), _ptr(means), _ptr(stds), _ptr(work), _ptr(n_total) if n_total is not None else None,
                int(n_base), _stream(mix.device)))
        else:
            _lib.check(_lib.load().asw_shift_stack_norm(_ptr(mix), _ptr(shifts), mip, N, B, M, T, _ptr(out),
                                                        _ptr(means), _ptr(stds), _ptr(work), _stream(mix.device)))
    return out[:N], means.view(N, 1, 1), stds.view(N, 1, 1)


def pcm16_to_f32(pcm, out=None):
    """int16 CUDA tensor -> float32 = pcm / 32768 (what soundfile/librosa produce for PCM_16 files)."""
    _require_cuda(pcm, "pcm", torch.int16)
    if out is None:
        out = torch.empty(pcm.shape, device=pcm.device, dtype=torch.float32)
    else:
        _require_cuda(out, "out", torch.float32)
        if out.numel() != pcm.numel():
            raise _lib.AswError("out must have as many elements as pcm")
    if pcm.numel():
        with torch.cuda.device(pcm.device):     # handle-less entry points launch on the current device
            _lib.check(_lib.load().asw_pcm16_to_f32(_ptr(pcm), _ptr(out), pcm.numel(), _stream(pcm.device)))
    return out


def patch_powers(x, window=12000, demean=True):
    """Per-row statistics of the separator's outputs (local_utils_3d.py:342-349, Mic_Array.py:288-296):
    ``x`` (N, T) float32 CUDA, de-meaned in place when ``demean`` ->
    (mean (N,), power = sum((x-mean)^2) (N,), max_avg_power (N,), start of that box (N,) int32)."""
    _require_cuda(x, "x", torch.float32)
    if x.dim() != 2:
        raise _lib.AswError("x must be (N, T)")
    N, T = x.shape
    mean = torch.empty((N,), device=x.device, dtype=torch.float32)
    power = torch.empty((N,), device=x.device, dtype=torch.float32)
    maxavg = torch.empty((N,), device=x.device, dtype=torch.float32)
    arg = torch.empty((N,), device=x.device, dtype=torch.int32)
    if N:
        with torch.cuda.device(x.device):     # handle-less entry points launch on the current device
            _lib.check(_lib.load().asw_patch_powers(_ptr(x), N, T, int(window), int(bool(demean)), _ptr(mean), _ptr(power),
                                                    _ptr(maxavg), _ptr(arg), _stream(x.device)))
    return mean, power, maxavg, arg


def offsets_to_shifts(offsets):
    """Patch.sample_offset list -> (N, M) int32 read offsets: r[0] = 0, r[c] = round_half_even(float32(off[c-1]))
    (sep/training/JointModel/network.py:81-82)."""
    off = np.asarray(offsets)
    if off.ndim == 1:
        off = off[None]
    v = np.concatenate([np.zeros((off.shape[0], 1)), off], axis=1).astype(np.float32)
    return np.rint(v).astype(np.int32)
