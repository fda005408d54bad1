"""Oracle: post-network scoring and hypercube subdivision -- numpy restatement.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows sep/helpers/local_utils_3d.py:
  * ``max_avg_power`` .................. :13-17
  * ``search_area`` ..................... :212-246
  * ``binary_area_divide_width`` ........ :248-335
  * ``binary_search_baseline`` .......... :339-388
and the patch-list assembly of ``Spotform_Small_Patch_Parallel``
(sep/Mic_Array.py:226-263).  Constants from sep/helpers/constants.py.
"""
import numpy as np
from scipy.ndimage import uniform_filter1d

from .prune_oracle import Patch

SPEED_OF_SOUND = 343.0
FS = 48000
MIN_AREA = 400
MIN_WIDTH = 3
MAX_BIG_PATCH = 30
MIN_WIDTH_REQUIRED = 2
SPOT_POWER_THRESHOLD1 = 0.008


def max_avg_power(x, window_size=12000):
    e = uniform_filter1d(x ** 2, size=window_size, mode="constant", origin=-window_size // 2)
    e = np.sqrt(np.abs(e))
    y = np.argmax(e)
    return e.max(), np.pad(x, (0, window_size))[y:y + window_size]


def point_tdoas(points, mic):
    """:220-224 -- TDoA (samples) of every area point to every mic vs mic 0."""
    d0 = (((points[0] - mic[0, 0]) ** 2 + (points[1] - mic[0, 1]) ** 2 + (points[2] - mic[0, 2]) ** 2) ** 0.5) / SPEED_OF_SOUND * FS
    rows = []
    for i in range(mic.shape[0] - 1):
        di = (((points[0] - mic[i + 1, 0]) ** 2 + (points[1] - mic[i + 1, 1]) ** 2 + (points[2] - mic[i + 1, 2]) ** 2) ** 0.5) / SPEED_OF_SOUND * FS
        rows.append(di - d0)
    return np.array(rows)


def divide(patch, samples0, upper_bound_pairwise):
    """binary_area_divide_width (:248-335)."""
    if upper_bound_pairwise is not None:
        patch.check_out(upper_bound_pairwise)
    area = patch.area_points
    cand = patch.sample_offset
    widths = patch.width_list
    npts = patch.area_size()
    D = cand.shape[0]
    if (np.amax(widths) / 2 <= MIN_WIDTH_REQUIRED) and npts <= MIN_AREA:
        return False, patch, samples0
    best_diff, best_patch, best_sample, keep8 = 2500000, None, None, False
    two_p = []
    for i in range(D):
        if widths[i] / 2 < MIN_WIDTH:
            continue
        two_p, two_s = [], []
        c0 = np.copy(cand)
        c0[i] -= widths[i] / 4          # float into int64: truncation (quirk B4)
        c1 = np.copy(cand)
        c1[i] += widths[i] / 4
        hw = np.copy(widths)
        hw[i] /= 2
        sizes = []
        for c in (c0, c1):
            p = Patch(c, hw, None)
            m = p.hyperbola_sample(samples0) == 1
            sz = np.sum(m)
            sizes.append(sz)
            if sz == 0:
                p.area_points = None
            else:
                p.area_points = area[:, m]
                two_p.append(p)
                two_s.append(samples0[:, m])
        diff = abs(sizes[0] - sizes[1])
        if hw[i] > MIN_WIDTH_REQUIRED:
            if not keep8:
                best_diff, best_patch, best_sample, keep8 = diff, two_p, two_s, True
            elif diff < best_diff:
                best_diff, best_patch, best_sample = diff, two_p, two_s
        else:
            if not keep8 and diff < best_diff:
                best_diff, best_patch, best_sample = diff, two_p, two_s
    if best_patch is None or len(two_p) == 0:
        return False, patch, samples0
    return True, best_patch, best_sample


def search_area(patch_list, mic, upper_bound_pairwise):
    """:212-246."""
    done = []
    samples_lists = [point_tdoas(patch_list[0].area_points, mic)]
    while True:
        nxt_p, nxt_s = [], []
        for i, patch in enumerate(patch_list):
            go, p, s = divide(patch, samples_lists[i], upper_bound_pairwise)
            if go:
                nxt_p.extend(p)
                nxt_s.extend(s)
            else:
                done.append(p)
        if len(nxt_p) == 0:
            break
        patch_list, samples_lists = nxt_p, nxt_s
    return done


def upper_bounds(mic):
    """sep/Mic_Array.py:113-115."""
    return np.array([(np.linalg.norm(mic[i] - mic[0]) + 0.08) / SPEED_OF_SOUND * FS
                     for i in range(1, mic.shape[0])])


def small_patch_list(candidates, mic):
    """sep/Mic_Array.py:226-263 -- fine patches + one centre width-2 patch per candidate."""
    ub = upper_bounds(mic)
    w2 = [2 for _ in range(mic.shape[0] - 1)]
    total, index = [], [0]
    for c in candidates:
        fine = search_area([c], mic, ub)
        centre = Patch(c.sample_offset, w2, None, c.peak_pos)
        if centre.center_pos() is not None:
            fine.append(centre)
        total.extend(fine)
        index.append(len(total))
    return total, index


def big_patch_select(sep_data, patch_list, mic):
    """binary_search_baseline after the network (:345-388) -> (kept, powers_with_dis, thr*1.2)."""
    pw_win, pw_dis = [], []
    for i in range(sep_data.shape[0]):
        sep_data[i, :] = sep_data[i, :] - np.mean(sep_data[i, :])
        p, _ = max_avg_power(sep_data[i, :])
        pw_win.append(p)
        cp = patch_list[i].center_pos()
        d = np.linalg.norm(cp - mic[0]) if cp.shape[0] == 3 else 4
        pw_dis.append(p * (d + 1))
    order = np.argsort(-1 * np.array(pw_win))
    thr = SPOT_POWER_THRESHOLD1
    keep = []
    for i in order:
        if pw_dis[i] < thr:
            continue
        if len(keep) >= MAX_BIG_PATCH:
            break
        keep.append(patch_list[i])
    return keep, pw_dis, thr * 1.2
