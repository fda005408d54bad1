"""Host-side mirror of the reference's ``SRP_PHAT`` (sep/Traditional_SP/SRP_Prunning.py:101-643)
for the hot path: same constructor arguments, attributes and method names, so the code that drives
it (``Mic_Array``) reads like the reference.  What changed underneath:

* the hypercube table build (``Map_3D_TDoA`` / ``search_cluster``, :277-344) is vectorised numpy +
  a native breadth-first walk (``asw_geometry_cluster``) that reproduces the reference's cluster and
  member order;
* the (G, F, P) float64 steering table (:221-243) is never built -- the device handle keeps the
  fractional pair lags instead (``native.NativeSRP``);
* ``SRP_Map_WINDOW_new`` (:384-433) is one call into libasw.so (CUDA, no CPU fallback);
* ``fill_powermap_torch`` / ``find_valid_peak_new`` (:347-357, :500-544) are vectorised.
"""
import ctypes
import os
import pickle

import numpy as np
import torch

from . import _lib, native
from .constants import ERR_TOLERANCE, KEEPOUT, PHAT_TOL, SRP_THRESHOLD_RATIO
from .patch import Patch


class Grid_cluster(object):
    """One hypercube: voxels sharing a quantised TDoA vector (SRP_Prunning.py:68-95)."""

    def __init__(self, sample_offset, pos, idx):
        self.sample_offset = sample_offset
        self.grids = pos
        self.index = idx

    def cluster_size(self):
        return len(self.grids)

    def center_pos(self):
        return np.mean(self.grids, axis=0)

    def equal(self, offset2):
        return np.array_equal(self.sample_offset, offset2)

    def dump(self):
        return [self.sample_offset, self.grids, self.index]


def hyperbola_offset(offset, pos, sample_offsets, width, first=None):
    """Points whose TDoA vector lies inside the closed box (SRP_Prunning.py:19-28).

    Same comparisons as the reference, evaluated progressively: dimension 0 on every voxel (``first``:
    an optional contiguous copy of ``offset[..., 0]``), the remaining dimensions only on the survivors.
    ``np.nonzero`` enumerates in C order, so the returned points are in the reference's order.  This
    scan is 0.2 s of the reference's 0.23 s pruning time."""
    D = offset.shape[-1]
    f = offset[..., 0] if first is None else first
    idx = np.nonzero((f >= sample_offsets[0] - width / 2) & (f <= sample_offsets[0] + width / 2))
    for i in range(1, D):
        if idx[0].size == 0:
            break
        v = offset[idx + (i,)]
        keep = (v >= sample_offsets[i] - width / 2) & (v <= sample_offsets[i] + width / 2)
        idx = tuple(a[keep] for a in idx)
    return pos[idx]


def hyperbola_area_sample(sample_list, sample_offsets, width):
    """Same box test on a list of TDoA vectors -> 0/1 (SRP_Prunning.py:30-39)."""
    lo = sample_offsets - width / 2
    hi = sample_offsets + width / 2
    return np.all((sample_list >= lo) & (sample_list <= hi), axis=1).astype(int)


def hyperbola_area_init(Axis_range, sample_offsets, width, Pos5, Offset5, Pos1, Offset1, first5=None, first1=None):
    """5 cm probe, then the 1 cm voxels inside the box, as (3, n) (SRP_Prunning.py:41-61)."""
    pts = hyperbola_offset(Offset5, Pos5, sample_offsets, width, first5)
    if pts.shape[0] == 0:
        return None
    x_min = max([Axis_range[0][0], pts[:, 0].min() - 0.05])
    x_max = min([Axis_range[0][1], pts[:, 0].max() + 0.05])
    ix0 = int(np.floor((x_min - Axis_range[0][0]) / 0.01))
    ix1 = int(np.ceil((x_max - Axis_range[0][0]) / 0.01))
    y_min = max([Axis_range[1][0], pts[:, 1].min() - 0.05])
    y_max = min([Axis_range[1][1], pts[:, 1].max() + 0.05])
    iy0 = int(np.floor((y_min - Axis_range[1][0]) / 0.01))
    iy1 = int(np.ceil((y_max - Axis_range[1][0]) / 0.01))
    pts = hyperbola_offset(Offset1[iy0:iy1, ix0:ix1, :, :], Pos1[iy0:iy1, ix0:ix1, :, :], sample_offsets, width,
                           None if first1 is None else first1[iy0:iy1, ix0:ix1, :])
    return pts.T


class SRP_PHAT(object):
    def __init__(self, mic_pos, freq_bins, Range_spk, C=343, FS=16000, n_fft=1024, grid_size=0.06,
                 grid_size_z=0.1, sample_resolution=4, threshold=0.03, WIDTH=8, device=None, cached=False,
                 cached_name=None, oversample=0, build_native=True):
        """``build_native=False`` builds only the host-side geometry (used by CPU tests of the host logic
        and by tools that inspect the hypercube table); scoring then raises instead of falling back."""
        if build_native:
            self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        else:
            self.device = torch.device("cpu")
        self.C = C
        self.FS = FS
        self.freq_bins = np.asarray(freq_bins)
        self.n_fft = n_fft
        self.mic_pos = np.asarray(mic_pos, dtype=np.float64)
        self.num_mic = self.mic_pos.shape[0]
        self.mic_center = self.mic_pos.mean(0)
        self.sample_resolution = sample_resolution
        self.WIDTH = WIDTH
        self.threshold = threshold
        self.Range_spk = Range_spk
        r = Range_spk
        self.x_grids = np.arange(r[0], r[1], grid_size)
        self.y_grids = np.arange(r[2], r[3], grid_size)
        self.z_grids = np.arange(r[4], r[5], grid_size_z)
        self.Lx, self.Ly, self.Lz = self.x_grids.shape[0], self.y_grids.shape[0], self.z_grids.shape[0]
        # :140-144, element by element: np.linalg.norm's dot-product rounding differs from a
        # vectorised sqrt(dx*dx + dy*dy) by 1 ulp, and dis_matrix feeds threshold comparisons
        self.dis_matrix = np.zeros((self.Lx, self.Ly))
        c2 = self.mic_center[:2]
        for ix in range(self.Lx):
            for iy in range(self.Ly):
                self.dis_matrix[ix][iy] = np.linalg.norm(np.array([self.x_grids[ix], self.y_grids[iy]]) - c2) + 1e-8
        self.Axis_range = [[r[0], r[1]], [r[2], r[3]], [r[4], r[5]]]
        self._fine = None
        b = self.mic_pos
        self.array_border = [b[:, 0].min() - KEEPOUT, b[:, 1].min() - KEEPOUT,
                             b[:, 0].max() + KEEPOUT, b[:, 1].max() + KEEPOUT]

        loaded = False
        if cached and cached_name is not None:
            fname = os.path.join(cached_name, "init_cached.pkl")
            if os.path.exists(fname):
                with open(fname, "rb") as fh:
                    data = pickle.load(fh)          # same keys as the reference's cache (:183-217)
                self.POWER_MAP = data["POWER_MAP"]
                self.POWER_INDEX = data["POWER_INDEX"]
                self.grids = data["grids"]
                self.clusters = [Grid_cluster(*a) for a in data["cluster"]]
                self._finish_tables()
                loaded = True
        if not loaded:
            self.Map_3D_TDoA()
            if cached and cached_name is not None:
                data = {"POWER_MAP": self.POWER_MAP, "POWER_INDEX": self.POWER_INDEX, "grids": self.grids,
                        "cluster": [c.dump() for c in self.clusters]}
                with open(os.path.join(cached_name, "init_cached.pkl"), "wb") as fh:
                    pickle.dump(data, fh)

        # device handle: fractional pair lags replace mode_mat_flat_real/imag (:221-243)
        self.native = None
        self.native_peaks = None
        self._native_select = None
        if build_native:
            self.native_peaks = native.NativePeaks(self.POWER_INDEX, self._member, self.dis_matrix,
                                                   self.grids.shape[0], self.threshold, SRP_THRESHOLD_RATIO,
                                                   device=self.device)
            lag = native.pair_lags(self.grids, self.mic_pos, FS, C)
            self.native = native.NativeSRP(lag, self.num_mic, device=self.device, bin0=int(self.freq_bins[0]),
                                           bin1=int(self.freq_bins[-1]) + 1, tol=PHAT_TOL, oversample=oversample)
        self.SRP_map = torch.zeros(self.grids.shape[0], device=self.device)
        self.peak_high_prio = []
        self.peak_low_prio = []
        self.MAX_POWER = -100
        self.Min_POWER = 0.0

    # ---- geometry ----------------------------------------------------------------------------
    def set_stft_pad_tail(self, enabled):
        """Frame convention of the STFT inside ``SRP_Map_WINDOW_new`` (:406, pyroomacoustics' ``analysis``, not
        vendored with the reference): False (default, assumption A1) drops the samples that do not fill a frame,
        True keeps them in one more zero-padded frame.  See ``asw_srp_set_frame_mode`` in include/asw.h."""
        if self.native is None:
            raise native._lib.AswError("no device handle (build_native=False)")
        self.native.set_pad_tail(enabled)

    def check_valid(self, idx):
        if idx[0] < 0 or idx[0] >= self.Lx or idx[1] < 0 or idx[1] >= self.Ly or idx[2] < 0 or idx[2] >= self.Lz:
            return False
        x, y = self.x_grids[idx[0]], self.y_grids[idx[1]]
        b = self.array_border
        return not (x > b[0] and y > b[1] and x < b[2] and y < b[3])

    def calculate_offset_pair(self, pos):
        d0 = np.linalg.norm(pos - self.mic_pos[0])
        return np.array([(np.linalg.norm(pos - self.mic_pos[i]) - d0) / self.C * self.FS
                         for i in range(1, self.num_mic)])

    def _fine_volumes(self):
        """Pos_5/Offset_5/Pos_1/Offset_1 (:148-170), built on first use (they are only needed by pruning)."""
        if self._fine is None:
            out = []
            r = self.Range_spk
            for step in (0.05, 0.01):
                xx = np.arange(r[0], r[1], step)
                yy = np.arange(r[2], r[3], step)
                zz = np.arange(r[4], r[5], 0.1)
                X, Y, Z = np.meshgrid(xx, yy, zz)
                pos = np.stack((X, Y, Z), axis=3)
                d0 = np.linalg.norm(pos - self.mic_pos[0, :], axis=3) / self.C * self.FS
                offs = np.stack([np.linalg.norm(pos - self.mic_pos[i, :], axis=3) / self.C * self.FS - d0
                                 for i in range(1, self.num_mic)], axis=3)
                out += [pos, offs]
            self._fine = out
        return self._fine

    def _first5(self):
        if getattr(self, "_first5_cache", None) is None:
            self._first5_cache = np.ascontiguousarray(self.Offset_5[..., 0])
        return self._first5_cache

    def _first1(self):
        if getattr(self, "_first1_cache", None) is None:
            self._first1_cache = np.ascontiguousarray(self.Offset_1[..., 0])
        return self._first1_cache

    Pos_5 = property(lambda self: self._fine_volumes()[0])
    Offset_5 = property(lambda self: self._fine_volumes()[1])
    Pos_1 = property(lambda self: self._fine_volumes()[2])
    Offset_1 = property(lambda self: self._fine_volumes()[3])

    def Map_3D_TDoA(self):
        Lx, Ly, Lz, M = self.Lx, self.Ly, self.Lz, self.num_mic
        X, Y, Z = np.meshgrid(self.x_grids, self.y_grids, self.z_grids, indexing="ij")
        b = self.array_border
        inside = (X > b[0]) & (Y > b[1]) & (X < b[2]) & (Y < b[3])
        valid = ~inside
        pos = np.stack((X, Y, Z), axis=3)

        def dist(m):
            d = pos - self.mic_pos[m]
            return np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2])

        d0 = dist(0)
        off = np.stack([(dist(i) - d0) / self.C * self.FS for i in range(1, M)], axis=3)
        q = np.round(off / self.sample_resolution).astype(np.int64) * self.sample_resolution
        q[~valid] = 0
        self.sample_tree = np.concatenate([valid[..., None].astype(np.int64), q], axis=3)

        lib = _lib.load()
        nvalid = int(valid.sum())
        label = np.empty((Lx, Ly, Lz), dtype=np.int32)
        order = np.empty(max(nvalid, 1), dtype=np.int32)
        start = np.empty(nvalid + 1, dtype=np.int32)
        ncl = ctypes.c_int32()
        qc = np.ascontiguousarray(q)
        vc = np.ascontiguousarray(valid.astype(np.uint8))
        _lib.check(lib.asw_geometry_cluster(qc.ctypes.data, vc.ctypes.data, Lx, Ly, Lz, M - 1, label.ctypes.data,
                                            order.ctypes.data, start.ctypes.data, ctypes.byref(ncl)))
        G = ncl.value
        ix, iy, iz = np.unravel_index(order[:nvalid], (Lx, Ly, Lz))
        allpos = np.stack([self.x_grids[ix], self.y_grids[iy], self.z_grids[iz]], axis=1)
        allidx = np.stack([ix, iy, iz], axis=1)
        self.clusters = []
        grids = np.empty((G, 3))
        for g in range(G):
            s, e = start[g], start[g + 1]
            mem_pos = allpos[s:e]
            self.clusters.append(Grid_cluster(q[ix[s], iy[s], iz[s]].copy(), mem_pos.tolist(), allidx[s:e].tolist()))
            grids[g] = np.mean(mem_pos, axis=0)
        self.grids = grids
        self.POWER_MAP = np.zeros((Lx, Ly, Lz))
        self.POWER_INDEX = np.where(label >= 0, label, 0).astype(int)
        self._member = label >= 0
        self.SRP_times = G

    def _finish_tables(self):
        """Member mask for the vectorised fill, rebuilt from the clusters of a loaded cache."""
        self._member = np.zeros(self.POWER_MAP.shape, dtype=bool)
        for c in self.clusters:
            idx = np.asarray(c.index)
            self._member[idx[:, 0], idx[:, 1], idx[:, 2]] = True

    # ---- scoring -----------------------------------------------------------------------------
    def reset(self):
        self.peak_high_prio = []
        self.peak_low_prio = []
        self.SRP_map = torch.zeros(self.grids.shape[0], device=self.device)

    def SRP_Map_WINDOW_new(self, signal, window=36000, tol=1e-8):
        """SRP_Prunning.py:384-434.  ``signal`` (M, T): numpy array or torch tensor (host or device).
        Side effects as in the reference: SRP_map (Tensor (G,)), MAX_POWER, Min_POWER, POWER_MAP."""
        if isinstance(signal, np.ndarray):
            signal = torch.from_numpy(np.ascontiguousarray(signal, dtype=np.float32))
        assert signal.shape[0] == self.num_mic
        if self.native is None:
            raise _lib.AswError("this SRP_PHAT was built with build_native=False: scoring needs the CUDA handle "
                                "(there is no CPU fallback)")
        sig = signal.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
        res = self.native.score(sig, window)[0]
        self.SRP_map = torch.maximum(self.SRP_map, res)
        # MAX_POWER / Min_POWER / POWER_MAP (:347-357, :432-433) are host-side views of the map: materialised on first
        # access.  Apply_SRP_PHAT prunes on the device and never reads them, so the drop-in call no longer pays a
        # blocking copy of the map plus the 1e5-voxel scatter per mixture.
        self._host_stale = True

    def load_map(self, srp_map):
        """Install an externally computed map (G,) as if SRP_Map_WINDOW_new had produced it
        (used by the batched front end, which scores many mixtures in one launch, and by tests)."""
        self.SRP_map = torch.from_numpy(np.ascontiguousarray(np.asarray(srp_map)))
        self._host_stale = True

    def _sync_host(self):
        """Bring the host-side views of the current map up to date (no-op when they are)."""
        if getattr(self, "_host_stale", False):
            self._host_stale = False
            m = self.SRP_map.detach().cpu().numpy()
            self._map_host_v = m
            self._max_power = float(m.max())
            self._min_power = float(m.min())
            self._power_map[self._member] = m[self.POWER_INDEX[self._member]]       # fill_powermap (:347-357) as one gather

    def fill_powermap_torch(self):
        """:347-357: POWER_MAP from the current map (kept for callers of the reference's method; the properties below
        do the same on demand)."""
        self._host_stale = True
        self._sync_host()

    @property
    def _map_host(self):
        self._sync_host()
        v = getattr(self, "_map_host_v", None)
        return v if v is not None else self.SRP_map.detach().cpu().numpy()

    @property
    def MAX_POWER(self):
        self._sync_host()
        return self._max_power

    @MAX_POWER.setter
    def MAX_POWER(self, v):
        self._max_power = v

    @property
    def Min_POWER(self):
        self._sync_host()
        return self._min_power

    @Min_POWER.setter
    def Min_POWER(self, v):
        self._min_power = v

    @property
    def POWER_MAP(self):
        self._sync_host()
        return self._power_map

    @POWER_MAP.setter
    def POWER_MAP(self, v):
        self._power_map = v

    # ---- pruning -----------------------------------------------------------------------------
    def find_valid_peak_new(self, rato=SRP_THRESHOLD_RATIO):
        """:500-544.  On the device when this object owns a CUDA handle (the map is already there);
        the numpy version below serves host-only geometry objects (``build_native=False``)."""
        if self.native_peaks is not None and rato == SRP_THRESHOLD_RATIO and self.SRP_map.is_cuda:
            peaks, count, _ = self.native_peaks.find(self.SRP_map.to(torch.float32))
            n = int(count[0])
            if n > self.native_peaks.max_peaks:
                raise _lib.AswError(f"{n} peak clusters exceed the device list of {self.native_peaks.max_peaks}")
            return [int(i) for i in peaks[0, :n].cpu().numpy()]
        return self._find_valid_peak_host(rato)

    def _find_valid_peak_host(self, rato=SRP_THRESHOLD_RATIO):
        thr = self.threshold[0] * self.MAX_POWER
        if thr < self.threshold[1]:
            thr = self.threshold[1]
        elif thr > self.threshold[2]:
            thr = self.threshold[2]
        thr2 = thr * rato
        pm = self.POWER_MAP
        NX, NY, NZ = pm.shape
        core = pm[2:-2, 2:-2, 1:-1]
        t1 = (thr * (0.9 + 1 / self.dis_matrix))[2:-2, 2:-2, None]
        t2 = (thr2 * (1 + 1 / self.dis_matrix))[2:-2, 2:-2, None]
        # >= every neighbour with dx, dy in [-2, 2], dz in {-1, 0} (sic, :523): running max of the window
        nb = np.maximum(pm[:, :, 1:-1], pm[:, :, 0:-2])
        mx = nb[2:-2, 2:-2]
        for dx in range(-2, 3):
            for dy in range(-2, 3):
                if dx == 0 and dy == 0:
                    continue
                mx = np.maximum(mx, nb[2 + dx:NX - 2 + dx, 2 + dy:NY - 2 + dy])
        cond2 = (core >= mx) & (core > t1) & (core <= t2)
        cond1 = core > t2
        vox = np.transpose(np.nonzero(cond2 | cond1))
        ids = self.POWER_INDEX[vox[:, 0] + 2, vox[:, 1] + 2, vox[:, 2] + 1]
        _, first = np.unique(ids, return_index=True)
        return [int(i) for i in ids[np.sort(first)]]

    @property
    def native_select(self):
        """Device handle of the greedy selection; built on first use (it uploads the 1 cm TDoA volume)."""
        if self._native_select is None and self.native is not None:
            self._native_select = native.NativeSelect(
                np.array([c.sample_offset for c in self.clusters], dtype=np.int32), self.Offset_5, self.Offset_1,
                self.Range_spk, self.Axis_range, self.WIDTH, device=self.device)
        return self._native_select

    def _area_builder(self, centres, width):
        return lambda: hyperbola_area_init(self.Axis_range, centres, width, self.Pos_5, self.Offset_5, self.Pos_1,
                                           self.Offset_1, self._first5(), self._first1())

    def patches_from_device(self, n, offsets, widths, peak_ids):
        """Patch objects for one mixture from asw_select_patches' outputs (host arrays); ``area_points`` is
        built on first access by the same code path the host selection uses."""
        n = int(n)
        D = self.num_mic - 1
        cent = np.asarray(offsets[:n]).astype(np.int64)                       # one conversion for the whole list;
        wl = np.repeat(np.asarray(widths[:n]).astype(np.int64)[:, None], D, axis=1)   # every patch gets its own rows
        pos = self.grids[np.asarray(peak_ids[:n]).astype(np.int64)]
        wtol = [int(w) + ERR_TOLERANCE for w in widths[:n]]
        return [Patch(cent[q], wl[q], None, pos[q], area_fn=self._area_builder(cent[q], wtol[q])) for q in range(n)]

    def local_source_adaptive_device(self):
        """:547-643 entirely on the device for this object's current map -> list[Patch]."""
        m = self.SRP_map.to(torch.float32).unsqueeze(0).contiguous()
        peaks, count, _ = self.native_peaks.find(m)
        n, off, wid, pk = self.native_select.select(m, peaks, count)
        # one device-to-host copy (and one synchronisation) for the counts and the three lists
        P, D = off.shape[1], off.shape[2]
        host = torch.cat([count[:1], n[:1], wid[0], pk[0], off[0].reshape(-1)]).cpu().numpy()
        if int(host[0]) > self.native_peaks.max_peaks:
            raise _lib.AswError(f"{int(host[0])} peak clusters exceed the device list of {self.native_peaks.max_peaks}")
        n_h = int(host[1])
        if n_h > self.native_select.max_patches:
            raise _lib.AswError(f"{n_h} patches exceed the device list of {self.native_select.max_patches}")
        return self.patches_from_device(n_h, host[2 + 2 * P:].reshape(P, D), host[2:2 + P], host[2 + P:2 + 2 * P])

    def local_source_adaptive(self, peak_index=None, peak_values=None):
        """:547-643 -> list[Patch].  ``peak_index`` / ``peak_values`` may be supplied by the batched front
        end (device peak picking for many mixtures at once); by default they come from this object's map."""
        if peak_index is None:
            peak_index = self.find_valid_peak_new()
        if len(peak_index) == 0:
            self.peak_candidate = np.zeros((0, 3))
            return []
        if peak_values is None:
            peaks = self._map_host[peak_index]
        else:
            peaks = np.asarray(peak_values)
        peaks_pos = self.grids[peak_index]
        self.peaks, self.peaks_pos = peaks, peaks_pos
        peaks_sample = np.array([self.clusters[i].sample_offset for i in peak_index])
        order = np.argsort(-1 * peaks)
        visited = np.zeros_like(peaks)
        peak_candidate, patch_candidate = [], []
        D = self.num_mic - 1
        W = self.WIDTH
        for pid in order:
            if visited[pid] >= 1:
                continue
            candidate = peaks_pos[pid, :]
            centre = peaks_sample[pid]
            peak_candidate.append(candidate)
            occupy = np.ones((D, W))
            for p in patch_candidate:
                delta = p.sample_offset - centre
                d1 = int(round((delta - p.width_list / 2 - W / 2).max()))
                d2 = int(round((delta + p.width_list / 2 + W / 2).min()))
                if d1 >= 0 or d2 <= 0:
                    continue
                if W + d1 < 0:
                    occupy[:, :] = 0
                else:
                    occupy[:, W + d1:] = 0
            widths, centres, dead = [], [], False
            for i in range(D):
                idx = np.where(occupy[i])[0]
                if idx.shape[0] == 0:
                    dead = True
                    break
                widths.append(idx.shape[0])
                centres.append(int(round(centre[i] + (idx[0] + idx[-1] - W + 1) / 2)))
            if dead:
                continue
            visited += hyperbola_area_sample(peaks_sample, centre, W + ERR_TOLERANCE)
            widths = np.array(widths)
            centres = np.array(centres)
            area = hyperbola_area_init(self.Axis_range, centres, widths[0] + ERR_TOLERANCE, self.Pos_5,
                                       self.Offset_5, self.Pos_1, self.Offset_1, self._first5(), self._first1())
            if area is None or area.shape[-1] == 0:
                continue
            patch_candidate.append(Patch(centres, widths, area, candidate))
        self.peak_candidate = np.array(peak_candidate)
        return patch_candidate
