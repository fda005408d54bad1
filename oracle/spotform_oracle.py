"""Oracle: everything Spotform_Small_Patch_Parallel does after the network -- numpy restatement.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows sep/Mic_Array.py:
  * ``weight_mean_pos`` ............................. :32-47
  * ``find_merge_center`` ........................... :50-81   (the widening loop breaks after factor 0)
  * ``Spotform_Small_Patch_Parallel`` ............... :225-395 (patch list :244-262 via subdivide_oracle; power gates
                                                       :289-345, SI-SDR clustering :347-369, merged centres :376-392)
and ``si_sdr`` of sep/helpers/eval_utils.py:11-39.  Constants from sep/helpers/constants.py.
Pinned by tests/golden/*_spotform.npz, outputs of the reference's own method run with a deterministic stand-in
separator (oracle/make_golden.py).
"""
import math

import numpy as np

from . import subdivide_oracle
from .prune_oracle import Patch

SPEED_OF_SOUND = 343.0
FS = 48000
SPOT_POWER_THRESHOLD2 = 0.01
USE_RELATIVE_SPOT_POWER = False
MIN_ERR = 1e-8


def si_sdr(estimated_signal, reference_signals):
    """eval_utils.py:11-39 with scaling=True."""
    rss = np.dot(reference_signals, reference_signals)
    a = np.dot(reference_signals, estimated_signal) / rss
    e_true = a * reference_signals
    e_res = estimated_signal - e_true
    return 10 * math.log10((e_true ** 2).sum() / ((e_res ** 2).sum() + MIN_ERR))


def weight_mean_pos(patch_list, powers, id_lists):
    """:32-47 -- power-weighted mean over the members within 75 % of the cluster head's power."""
    total_pos = np.zeros((3,))
    total_power = 0
    max_power = powers[id_lists[0]]
    total_offsets = np.zeros(patch_list[0].sample_offset.shape)
    for _id in id_lists:
        if powers[_id] < max_power * 0.75:
            continue
        total_pos += powers[_id] * patch_list[_id].center_pos()
        total_offsets += powers[_id] * patch_list[_id].sample_offset
        total_power += powers[_id]
    return total_pos / total_power, total_offsets / total_power


def general_area(patch, points, mic):
    """Patch.hyperbola_general_area (Patch_3D.py:28-38) on (3, n) points."""
    d0 = (((points[0] - mic[0, 0]) ** 2 + (points[1] - mic[0, 1]) ** 2 + (points[2] - mic[0, 2]) ** 2) ** 0.5) / SPEED_OF_SOUND * FS
    z = np.ones(points.shape[1], dtype=bool)
    for i in range(mic.shape[0] - 1):
        di = (((points[0] - mic[i + 1, 0]) ** 2 + (points[1] - mic[i + 1, 1]) ** 2 + (points[2] - mic[i + 1, 2]) ** 2) ** 0.5) / SPEED_OF_SOUND * FS
        lo = patch.sample_offset[i] - patch.width_list[i] / 2 - 1e-3
        hi = patch.sample_offset[i] + patch.width_list[i] / 2 + 1e-3
        z &= ((di - d0) >= lo) & ((di - d0) <= hi)
    return z


def find_merge_center(merged_offsets, init_area, mic, big_patch_center):
    """:50-81.  ``for factor in range(4): ... break`` only ever evaluates factor 0, i.e. the same width 3 again."""
    D = mic.shape[0] - 1
    pc = Patch(merged_offsets, [3 for _ in range(D)], None)
    area = general_area(pc, init_area, mic)
    if area.sum() == 0:
        pc.width_list = [3 for _ in range(D)]
        area = general_area(pc, init_area, mic)
        if area.sum() > 0:
            pc.area_points = init_area[:, area]
        else:
            pc.peak_pos = big_patch_center
    else:
        pc.area_points = init_area[:, area]
    return pc


def small_patch_parallel(mix, candidates, spot_model, mic, min_trigger_power=0.5, relative_threshold=None):
    """:225-395 -> list of (patch_center, audio, power, tag, {"audio_offset", "localization_offset"}, big_label=-1).
    ``candidates``: oracle Patch objects with ``area_points`` (mutated by check_out like the reference's)."""
    thr_new = SPOT_POWER_THRESHOLD2
    if USE_RELATIVE_SPOT_POWER:
        thr_new = min([SPOT_POWER_THRESHOLD2, relative_threshold])
    ub = subdivide_oracle.upper_bounds(mic)
    w2 = [2 for _ in range(mic.shape[0] - 1)]
    total, index, init_areas, centres = [], [0], [], []
    for c in candidates:                                                   # :244-262
        fine = subdivide_oracle.search_area([c], mic, ub)
        init_areas.append(c.area_points)
        centre_patch = Patch(c.sample_offset, w2, None, c.peak_pos)
        centres.append(centre_patch.center_pos())
        if centres[-1] is not None:
            fine.append(centre_patch)
        total.extend(fine)
        index.append(len(total))
    sep_total = spot_model.shift_and_sep(mix, total, Strict=1)             # :263
    out = []
    for i in range(len(index) - 1):
        sep = sep_total[index[i]:index[i + 1]]
        patches = total[index[i]:index[i + 1]]
        powers, powers2 = [], []
        for j in range(len(patches)):                                      # :289-296
            sep[j, :] = sep[j, :] - np.mean(sep[j, :])
            powers.append(np.sum(sep[j, :] ** 2))
            powers2.append(subdivide_oracle.max_avg_power(sep[j, :])[0])
        cpos = candidates[i].center_pos()
        d = np.linalg.norm(cpos - mic[0]) if cpos.shape[0] == 3 else 4       # :334-337
        if np.amax(powers2) < thr_new / (1 + d):                            # :338-342
            continue
        order = np.argsort(-1 * np.array(powers))                           # :345
        clusters = {}
        min_trigger2 = min_trigger_power / (3 * 48000) * sep.shape[1]       # :349
        for _id in order:                                                   # :350-366
            unique = True
            d = np.linalg.norm(patches[_id].center_pos() - mic[0])
            if powers2[_id] < thr_new / (1 + d) or powers[_id] < min_trigger2:
                continue
            for cluster_id in clusters:
                head = clusters[cluster_id][0]
                if si_sdr(sep[_id, :], sep[head]) > -4:
                    clusters[head].append(_id)
                    unique = False
                    break
            if unique:
                clusters[_id] = [_id]
        for cluster_id in clusters:                                         # :375-392
            _, offsets = weight_mean_pos(patches, powers, clusters[cluster_id])
            pc = find_merge_center(offsets, init_areas[i], mic, centres[i])
            out.append((pc, sep[cluster_id, :], powers[cluster_id], str(i) + "_" + str(cluster_id),
                        {"audio_offset": patches[cluster_id].sample_offset, "localization_offset": offsets}, -1))
    return out


# ---- Clustering_new (sep/Mic_Array.py:399-500) with split_wav / split_wise_sisdr (sep/helpers/eval_utils.py:43-82) -------
def check_sisnr_win(values, t1=-2, t2=-7):
    """sep/Mic_Array.py:18-28."""
    return any(v > t1 for v in values) and not any(v < t2 for v in values)


def split_wav(wav, librosa, top_db=18):
    """eval_utils.py:43-70 around the two librosa calls (``librosa``: the module to use -- the real one, or
    oracle/fake_librosa.py in tests)."""
    power = librosa.feature.rms(y=wav, frame_length=1024, hop_length=256)
    if np.amax(power) < 0.04:
        intervals = librosa.effects.split(wav, top_db=top_db, ref=0.04, frame_length=1024, hop_length=256)
    else:
        intervals = librosa.effects.split(wav, top_db=top_db, frame_length=1024, hop_length=256)
    segs = []
    for a, b in intervals:
        n = b - a
        if n < 1000:
            continue
        if n > 4000:
            k = n // 4000
            for i in range(k):
                segs.append([a + i * 4000, b if i >= k - 1 else a + (i + 1) * 4000])
        else:
            segs.append([a, b])
    return segs


def clustering_new(output_pair, librosa, spot_times=0):
    """sep/Mic_Array.py:399-500 without ground truth -> (audio_final, patch_final, spot_times, [])."""
    cands = sorted(output_pair, key=lambda x: -x[2])
    clusters = {}
    for i in range(len(cands)):
        unique = True
        centre1, audio1 = cands[i][0].center_pos(), cands[i][1]
        segs = split_wav(audio1, librosa)
        if len(segs) == 0:
            continue
        per_head = []
        for cid in clusters:
            head = clusters[cid][0]
            audio2, centre2 = cands[head][1], cands[head][0].center_pos()
            sim = si_sdr(audio1, audio2)
            win = [si_sdr(audio1[a:b], audio2[a:b]) for a, b in segs]
            per_head.append(win)
            if sim > -1 or check_sisnr_win(win) or np.linalg.norm(centre1[:2] - centre2[:2]) < 0.45:
                clusters[head].append(i)
                unique = False
                break
        if per_head and check_sisnr_win(np.amax(np.array(per_head), axis=0), -1, -5):
            unique = False
        if unique:
            clusters[i] = [i]
    heads = [clusters[c][0] for c in clusters]
    return [cands[h][1] for h in heads], [cands[h] for h in heads], spot_times, []
