"""Host glue of the localisation stages: post-network power scoring and hypercube subdivision.

Mirror of the reference's sep/helpers/local_utils_3d.py (``max_avg_power`` :13-17, ``search_area``
:212-246, ``binary_area_divide_width`` :248-335, ``binary_search_baseline`` :339-388) and of
``si_sdr`` (sep/helpers/eval_utils.py:11-39).  Same names, arguments and quirks; plotting dropped.
"""
import math

import numpy as np
from scipy.ndimage import uniform_filter1d

from .constants import (FS, MAX_BIG_PATCH, MIN_AREA, MIN_WIDTH, MIN_WIDTH_REQUIRED, SPEED_OF_SOUND,
                        SPOT_POWER_THRESHOLD1, USE_RELATIVE_SPOT_POWER)
from .patch import Patch

MIN_ERR = 1e-8


def si_sdr(estimated_signal, reference_signals, scaling=True):
    """Scale-invariant SDR in dB (eval_utils.py:11-39)."""
    Rss = np.dot(reference_signals, reference_signals)
    a = np.dot(reference_signals, estimated_signal) / Rss if scaling else 1
    e_true = a * reference_signals
    e_res = estimated_signal - e_true
    return 10 * math.log10((e_true ** 2).sum() / ((e_res ** 2).sum() + MIN_ERR))


def split_wav(wav, top_db=18):
    """Non-silent segments of an output, 1000 .. 4000+ samples each (eval_utils.py:43-70).  Like the reference this
    needs ``librosa`` (``feature.rms`` and ``effects.split``); it is imported here, on first use, because nothing else
    on the path does."""
    import librosa
    MIN_SEG, MAX_SEG = 1000, 4000
    power_list = librosa.feature.rms(y=wav, frame_length=1024, hop_length=256)
    max_ref = np.amax(power_list)
    split_threshold = 0.04
    if max_ref < split_threshold:
        intervals = librosa.effects.split(wav, top_db=top_db, ref=split_threshold, frame_length=1024, hop_length=256)
    else:
        intervals = librosa.effects.split(wav, top_db=top_db, frame_length=1024, hop_length=256)
    finetune_seg = []
    for indexes in intervals:
        interval_len = indexes[1] - indexes[0]
        if interval_len < MIN_SEG:
            continue
        elif interval_len > MAX_SEG:
            num_seg = interval_len // MAX_SEG
            for i in range(num_seg):
                if i >= num_seg - 1:
                    finetune_seg.append([indexes[0] + i * MAX_SEG, indexes[1]])
                else:
                    finetune_seg.append([indexes[0] + i * MAX_SEG, indexes[0] + (i + 1) * MAX_SEG])
        else:
            finetune_seg.append([indexes[0], indexes[1]])
    return finetune_seg


def split_wise_sisdr(estimated_signal, reference_signals, seg_index):
    """SI-SDR of every segment (eval_utils.py:73-82)."""
    assert len(seg_index) > 0
    return [si_sdr(estimated_signal[a:b], reference_signals[a:b]) for a, b in seg_index]


def max_avg_power(x, window_size=12000):
    """Largest RMS over any ``window_size`` box and the samples of that box (:13-17)."""
    e = uniform_filter1d(x ** 2, size=window_size, mode="constant", origin=-window_size // 2)
    e = np.sqrt(np.abs(e))
    y = np.argmax(e)
    return e.max(), np.pad(x, (0, window_size))[y:y + window_size]


def _tdoa_rows(points, mic_positions, fs=FS):
    """TDoA (samples at ``fs``) of every point to mics 1.. vs mic 0 (:220-224; the reference hard-codes 48 kHz)."""
    d0 = (((points[0, :] - mic_positions[0, 0]) ** 2 + (points[1, :] - mic_positions[0, 1]) ** 2
           + (points[2, :] - mic_positions[0, 2]) ** 2) ** 0.5) / SPEED_OF_SOUND * fs
    rows = []
    for i in range(mic_positions.shape[0] - 1):
        di = (((points[0, :] - mic_positions[i + 1, 0]) ** 2 + (points[1, :] - mic_positions[i + 1, 1]) ** 2
               + (points[2, :] - mic_positions[i + 1, 2]) ** 2) ** 0.5) / SPEED_OF_SOUND * fs
        rows.append(di - d0)
    return np.array(rows)


def search_area(patch_list, mic_positions, upper_bound_pairwise, fs=FS):
    """Recursively halve a coarse hypercube until every dimension is fine enough (:212-246).  ``fs``: the sampling
    rate the patches' offsets are expressed in (the reference's constant 48 kHz by default)."""
    finished = []
    samples_lists = [_tdoa_rows(patch_list[0].area_points, mic_positions, fs)]
    while True:
        next_patches, next_samples = [], []
        for i, patch in enumerate(patch_list):
            go, nxt, smp = binary_area_divide_width(patch, samples_lists[i], mic_positions, upper_bound_pairwise)
            if go:
                next_patches.extend(nxt)
                next_samples.extend(smp)
            else:
                finished.append(nxt)
        if len(next_patches) == 0:
            break
        patch_list, samples_lists = next_patches, next_samples
    return finished


def binary_area_divide_width(patch, samples0, mic_positions, upper_bound_pairwise):
    """Split along the dimension that balances the member points best (:248-335)."""
    if upper_bound_pairwise is not None:
        patch.check_out(upper_bound_pairwise)
    area = patch.area_points
    candidates = patch.sample_offset
    widths = patch.width_list
    num_points = patch.area_size()
    num_pair = candidates.shape[0]
    if (np.amax(widths) / 2 <= MIN_WIDTH_REQUIRED) and num_points <= MIN_AREA:
        return False, patch, samples0

    min_difference, min_patch, min_sample = 2500000, None, None
    remain_width_8 = False
    two_patches = []
    for i in range(num_pair):
        if widths[i] / 2 < MIN_WIDTH:
            continue
        two_patches, two_samples, sizes = [], [], []
        half_width = np.copy(widths)
        half_width[i] /= 2                      # float into int64: truncates (reference quirk)
        for sign in (-1, 1):
            centre = np.copy(candidates)
            if sign < 0:
                centre[i] -= widths[i] / 4
            else:
                centre[i] += widths[i] / 4
            half = Patch(centre, half_width, None)
            inside = half.hyperbola_sample(samples0) == 1
            size = np.sum(inside)
            sizes.append(size)
            if size == 0:
                half.area_points = None
            else:
                half.area_points = area[:, inside]
                two_patches.append(half)
                two_samples.append(samples0[:, inside])
        diff = abs(sizes[0] - sizes[1])
        if half_width[i] > MIN_WIDTH_REQUIRED:
            if not remain_width_8:
                min_difference, min_patch, min_sample, remain_width_8 = diff, two_patches, two_samples, True
            elif diff < min_difference:
                min_difference, min_patch, min_sample = diff, two_patches, two_samples
        elif not remain_width_8 and diff < min_difference:
            min_difference, min_patch, min_sample = diff, two_patches, two_samples
    if min_patch is None or len(two_patches) == 0:
        return False, patch, samples0
    return True, min_patch, min_sample


def binary_search_baseline(mix_data, spot_model, patch_list, mic_positions):
    """Coarse stage: run the spot model on every coarse hypercube, keep the energetic ones (:339-388)."""
    device_powers = getattr(spot_model, "shift_and_sep_powers", None)
    if device_powers is not None:           # de-mean + max_avg_power of every row in one libasw.so launch
        sep_data, _, win = device_powers(mix_data, patch_list, Strict=0)
    else:                                   # a foreign spot model: the reference's per-row host loop
        sep_data = spot_model.shift_and_sep(mix_data, patch_list, Strict=0)
        win = None
    powers_win, powers_with_dis = [], []
    for i in range(sep_data.shape[0]):
        if win is None:
            sep_data[i, :] = sep_data[i, :] - np.mean(sep_data[i, :])
            p, _ = max_avg_power(sep_data[i, :])
        else:
            p = win[i]
        powers_win.append(p)
        centre = patch_list[i].center_pos()
        d = np.linalg.norm(centre - mic_positions[0]) if centre.shape[0] == 3 else 4
        powers_with_dis.append(p * (d + 1))
    sort_idx = np.argsort(-1 * np.array(powers_win))
    if USE_RELATIVE_SPOT_POWER:
        relative_threshold = min([0.4 * max(powers_with_dis), SPOT_POWER_THRESHOLD1])
    else:
        relative_threshold = SPOT_POWER_THRESHOLD1
    valid_patch = []
    for i in sort_idx:
        if powers_with_dis[i] < relative_threshold:
            continue
        if len(valid_patch) >= MAX_BIG_PATCH:
            break
        valid_patch.append(patch_list[i])
    return valid_patch, powers_with_dis, relative_threshold * 1.2
