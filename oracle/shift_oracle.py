"""Oracle: per-hypercube circular channel shift, stacking and input normalisation.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  * ``roll_by_gather`` ............ sep/training/JointModel/network.py:12-25
  * the shift loop ................ sep/training/JointModel/network.py:75-83
  * window-condition embedding .... sep/training/JointModel/network.py:62-73
  * ``normalize_input`` ........... sep/training/SpeakerLocalization/network.py:28-40
  * ``unnormalize_input`` ......... sep/training/SpeakerLocalization/network.py:42-47
"""
import numpy as np


def shift_indices(sample_offset):
    """network.py:81-82: ``-round(float32([0, *offset]))`` as int64; the gather
    then reads ``mix[c, (t - shifts[c]) mod T]``, so the read offset is
    ``r[c] = +round_half_even(float32(offset))`` with ``r[0] = 0``."""
    v = np.asarray([0, *sample_offset], dtype=np.float32)
    return np.rint(v).astype(np.int64)          # rint == round-half-to-even == torch.round


def roll_by_gather(mat, shifts):
    """network.py:12-25 with dim=1.  ``shifts`` (M,) int64 as passed by the caller
    (i.e. NEGATED read offsets): out[c, t] = mat[c, (t - shifts[c]) mod T]."""
    M, T = mat.shape
    idx = (np.arange(T)[None, :] - np.asarray(shifts).reshape(M, 1)) % T     # python-style mod, in [0, T)
    return np.take_along_axis(mat, idx, axis=1)


def shift_stack(mix, offsets_list):
    """network.py:75-83 for one batch -> (N, M, T) float32."""
    mix = np.asarray(mix, dtype=np.float32)
    out = np.empty((len(offsets_list),) + mix.shape, dtype=np.float32)
    for n, off in enumerate(offsets_list):
        out[n] = roll_by_gather(mix, -shift_indices(off))
    return out


def window_condition(n, strict):
    """network.py:62-73: Strict==1 -> [1, 0]; else [0, 1]."""
    e = np.zeros((n, 2), dtype=np.float32)
    e[:, 0 if strict == 1 else 1] = 1
    return e


def normalize_input(data):
    """SpeakerLocalization/network.py:28-40 in float32 (unbiased std, like torch)."""
    data = np.asarray(data, dtype=np.float32)
    q = np.rint(data * np.float32(2 ** 15)) / np.float32(2 ** 15)
    ref = q.mean(1, dtype=np.float32)
    means = ref.mean(1, dtype=np.float64).astype(np.float32)[:, None, None]
    stds = ref.std(1, ddof=1, dtype=np.float64).astype(np.float32)[:, None, None]
    return (q - means) / stds, means, stds


def unnormalize_input(data, means, stds):
    return data * stds + means
