"""Oracle: pruning of the SRP map into hypercube patches -- numpy restatement.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  * ``Patch`` ............................ sep/Traditional_SP/Patch_3D.py:3-93
  * box membership helpers ................ sep/Traditional_SP/SRP_Prunning.py:19-61
  * ``fill_powermap_torch`` ............... SRP_Prunning.py:347-357
  * ``find_valid_peak_new`` ............... SRP_Prunning.py:500-544
  * ``local_source_adaptive`` ............. SRP_Prunning.py:547-643
Quirks kept on purpose: dz in {-1, 0} only (B1); float->int64 truncation in
``Patch.check_out`` (B4); ``elif delta1 < 0`` shadows the ``delta2`` branch.
"""
import numpy as np

ERR_TOLERANCE = 0.2          # SRP_Prunning.py:17


class Patch:
    """Hypercube record: centre TDoAs, per-dimension widths, member points."""

    def __init__(self, sample_offset, width_list, area_points, peak_pos=None):
        self.sample_offset = sample_offset
        self.width_list = np.copy(width_list)
        self.area_points = area_points
        self.num_pair = sample_offset.shape[0]
        self.peak_pos = peak_pos

    def area_size(self):
        if self.area_points is None or self.area_points.shape[1] == 0:
            return 0
        return self.area_points.shape[1]

    def center_pos(self):
        if self.peak_pos is not None:
            return self.peak_pos
        if self.area_points is None or self.area_points.shape[1] == 0:
            return None
        return np.mean(self.area_points, axis=1)

    def _bounds(self, i):
        return (self.sample_offset[i] - self.width_list[i] / 2 - 1e-3,
                self.sample_offset[i] + self.width_list[i] / 2 + 1e-3)

    def hyperbola_sample(self, offset):                       # Patch_3D.py:40-47
        z = 1
        for i in range(offset.shape[0]):
            lo, hi = self._bounds(i)
            z = z & (offset[i, :] >= lo) & (offset[i, :] <= hi)
        return z.astype(int)

    def hyperbola_general_area(self, X, Y, Z, mic, c, fs):    # Patch_3D.py:28-38
        z = 1
        d0 = (((X - mic[0, 0]) ** 2 + (Y - mic[0, 1]) ** 2 + (Z - mic[0, 2]) ** 2) ** 0.5) / c * fs
        for i in range(mic.shape[0] - 1):
            f = (((X - mic[i + 1, 0]) ** 2 + (Y - mic[i + 1, 1]) ** 2 + (Z - mic[i + 1, 2]) ** 2) ** 0.5) / c * fs - d0
            lo, hi = self._bounds(i)
            z = z & (f >= lo) & (f <= hi)
        return z.astype(int)

    def check_out(self, upper_bound_pairwise):                # Patch_3D.py:69-87
        for i in range(self.num_pair):
            ub = upper_bound_pairwise[i]
            while True:
                if abs(self.sample_offset[i]) <= ub or self.width_list[i] <= 4:
                    break
                res = self.width_list[i]
                if self.sample_offset[i] > ub:
                    self.sample_offset[i] = self.sample_offset[i] - res / 4
                elif self.sample_offset[i] < -ub:
                    self.sample_offset[i] = self.sample_offset[i] + res / 4
                self.width_list[i] = res / 2


def box_points(offset, pos, centre, width):
    """SRP_Prunning.py:19-28 -- points whose TDoA vector lies in the closed box."""
    z = 1
    for i in range(offset.shape[-1]):
        z = z & (offset[..., i] >= centre[i] - width / 2) & (offset[..., i] <= centre[i] + width / 2)
    return pos[z == 1]


def box_mask_rows(rows, centre, width):
    """SRP_Prunning.py:30-39 -- same test on an (n, D) list of TDoA vectors."""
    z = 1
    for i in range(centre.shape[-1]):
        z = z & (rows[:, i] >= centre[i] - width / 2) & (rows[:, i] <= centre[i] + width / 2)
    return z.astype(int)


def area_init(axis_range, centre, width, Pos5, Offset5, Pos1, Offset1):
    """SRP_Prunning.py:41-61 -- 5 cm probe, then 1 cm voxels inside the box -> (3, n)."""
    pts = box_points(Offset5, Pos5, centre, width)
    if pts.shape[0] == 0:
        return None
    x0 = max(axis_range[0][0], pts[:, 0].min() - 0.05)
    x1 = min(axis_range[0][1], pts[:, 0].max() + 0.05)
    ix0 = int(np.floor((x0 - axis_range[0][0]) / 0.01))
    ix1 = int(np.ceil((x1 - axis_range[0][0]) / 0.01))
    y0 = max(axis_range[1][0], pts[:, 1].min() - 0.05)
    y1 = min(axis_range[1][1], pts[:, 1].max() + 0.05)
    iy0 = int(np.floor((y0 - axis_range[1][0]) / 0.01))
    iy1 = int(np.ceil((y1 - axis_range[1][0]) / 0.01))
    pts = box_points(Offset1[iy0:iy1, ix0:ix1], Pos1[iy0:iy1, ix0:ix1], centre, width)
    return pts.T


def fill_powermap(srp_map, clusters, shape, power_map=None, power_index=None):
    """SRP_Prunning.py:347-357."""
    if power_map is None:
        power_map = np.zeros(shape)
    if power_index is None:
        power_index = np.zeros(shape, dtype=int)
    for i, c in enumerate(clusters):
        for ix, iy, iz in c[2]:
            power_map[ix, iy, iz] = srp_map[i]
            power_index[ix, iy, iz] = i
    return power_map, power_index


def adaptive_thresholds(max_power, threshold=(0.15, 0.015, 0.05), ratio=4):
    """SRP_Prunning.py:501-504."""
    t = threshold[0] * max_power
    if t < threshold[1]:
        t = threshold[1]
    elif t > threshold[2]:
        t = threshold[2]
    return t, t * ratio


def find_valid_peaks(power_map, power_index, dis_matrix, max_power, n_grids,
                     threshold=(0.15, 0.015, 0.05), ratio=4):
    """SRP_Prunning.py:500-544 -> cluster ids, first-seen order."""
    t1, t2 = adaptive_thresholds(max_power, threshold, ratio)
    NX, NY, NZ = power_map.shape
    core = power_map[2:-2, 2:-2, 1:-1]
    th1 = np.repeat((t1 * (0.9 + 1 / dis_matrix))[2:-2, 2:-2, None], NZ - 2, axis=2)
    th2 = np.repeat((t2 * (1 + 1 / dis_matrix))[2:-2, 2:-2, None], NZ - 2, axis=2)
    is_max = np.ones_like(core, dtype=bool)
    for dx in range(-2, 3):
        for dy in range(-2, 3):
            for dz in range(-1, 1):           # sic: dz in {-1, 0}
                if dx == 0 and dy == 0 and dz == 0:
                    continue
                is_max &= core >= power_map[2 + dx:NX - 2 + dx, 2 + dy:NY - 2 + dy, 1 + dz:NZ - 1 + dz]
    cond2 = is_max & (core > th1) & (core <= th2)
    cond1 = core > th2
    vox = np.transpose(np.nonzero(cond2 | cond1))
    seen = np.zeros(n_grids, dtype=bool)
    ids = []
    for a, b, c in vox:
        g = power_index[a + 2, b + 2, c + 1]
        if seen[g]:
            continue
        seen[g] = True
        ids.append(int(g))
    return ids


def local_source_adaptive(srp_map, peak_index, grids, cluster_offsets, num_mic, geom, WIDTH=8):
    """SRP_Prunning.py:547-643 -> list[Patch].

    ``cluster_offsets[i]`` is cluster i's quantised TDoA vector;
    ``geom`` provides Axis_range, Pos_5, Offset_5, Pos_1, Offset_1."""
    peaks = np.asarray(srp_map)[peak_index]
    peaks_pos = grids[peak_index]
    peaks_sample = np.array([cluster_offsets[i] for i in peak_index])
    order = np.argsort(-1 * peaks)
    visited = np.zeros_like(peaks)
    patches = []
    D = num_mic - 1
    for pid in order:
        if visited[pid] >= 1:
            continue
        cand = peaks_pos[pid, :]
        centre = peaks_sample[pid]
        W = WIDTH
        occupy = np.ones((D, W))
        for p in patches:
            delta = p.sample_offset - centre
            lo1 = delta - p.width_list / 2
            hi1 = delta + p.width_list / 2
            d1 = int(round((lo1 - W / 2).max()))
            d2 = int(round((hi1 + W / 2).min()))
            if d1 >= 0 or d2 <= 0:
                continue
            elif d1 < 0:
                if W + d1 < 0:
                    occupy[:, :] = 0
                else:
                    occupy[:, W + d1:] = 0
            elif d2 > 0:      # unreachable, kept for fidelity
                if d2 > W:
                    occupy[:, :] = 0
                else:
                    occupy[:, 0:d2] = 0
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
        visited += box_mask_rows(peaks_sample, centre, W + ERR_TOLERANCE)
        widths = np.array(widths)
        centres = np.array(centres)
        area = area_init(geom.Axis_range, centres, widths[0] + ERR_TOLERANCE,
                         geom.Pos_5, geom.Offset_5, geom.Pos_1, geom.Offset_1)
        if area is None or area.shape[-1] == 0:
            continue
        patches.append(Patch(centres, widths, area, cand))
    return patches
