"""Seeded synthetic scenes for tests and benchmarks.

The reference has no sample data; its scene recipe lives in
datasets/generate_dataset.py (desk-edge array :341-376, desk placement :378-420,
speaker region / ROI :512-555, speaker spacing :560-580).  This module follows
that recipe's geometry with ``numpy.random.default_rng(seed)`` and replaces the
room simulation + VCTK speech (both unavailable) by free-field propagation of
band-limited noise bursts: integer-sample delay ``round(d / C * fs)``, 1/d
gain, white sensor noise.
"""
from dataclasses import dataclass

import numpy as np

from .constants import FS, SPEED_OF_SOUND

DESK_LENGTH = (1.2, 2.0)
DESK_WIDTH = (0.6, 1.2)
THETA_MAX_DEV = np.deg2rad(6)
EXPAND_MAX_DEV = 0.08
MIC_HEIGHT = 0.02
WALL_KEEPOUT = 0.5
SPK_RANGE_W = 3.0
SPK_RANGE_H = 4.5
MIN_SPEAKER_DIST = 0.51
SPK_Z = (0.1, 0.8)
ROOM = 7.0


@dataclass
class Scene:
    mic_positions: np.ndarray      # (M, 3) float64
    roi: list                      # [x0, x1, y0, y1, z0, z1]
    fs: int


def desk_array(n_mics, rng, fs=FS):
    """Desk-edge array against the left wall of a ROOM x ROOM room (generate_dataset.py:341-399)."""
    L = rng.uniform(*DESK_LENGTH)
    W = rng.uniform(*DESK_WIDTH)
    mid = np.arctan(L / 2 / W)
    ang = np.linspace(0, np.pi, n_mics - 1) - np.pi / 2
    rel = np.zeros((n_mics, 2))
    for i in range(n_mics - 1):
        a = ang[i] + rng.uniform(-THETA_MAX_DEV, THETA_MAX_DEV)
        if -mid < a < mid:
            r = W / np.cos(a)
        else:
            r = L / 2 / np.sin(abs(a))
        r -= 0.04
        rel[i + 1] = [r * np.cos(a) + rng.uniform(-EXPAND_MAX_DEV, EXPAND_MAX_DEV),
                      r * np.sin(a) + rng.uniform(-EXPAND_MAX_DEV, EXPAND_MAX_DEV)]
    cx = 0.1 + rng.uniform(0.0, 0.35)
    cy = rng.uniform(0.1 + 1.8, ROOM - 0.1 - 1.8)
    xy = rel + np.array([cx, cy])
    mic = np.concatenate([xy, np.full((n_mics, 1), MIC_HEIGHT)], axis=1)
    # speaker region and ROI (generate_dataset.py:533-555, wall 0) + eval's z bump (eval_model.py:84)
    x0 = max(cx + 0.25, 0.0 + WALL_KEEPOUT)
    x1 = min(cx + SPK_RANGE_H, ROOM - WALL_KEEPOUT)
    y0 = max(cy - SPK_RANGE_W, 0.0 + WALL_KEEPOUT)
    y1 = min(cy + SPK_RANGE_W, ROOM - WALL_KEEPOUT)
    roi = [x0 - 0.1, x1 + 0.1, y0 - 0.1, y1 + 0.1, 0.0, 0.9 + 0.02]
    return Scene(mic, [float(v) for v in roi], fs)


def table_array(n_mics, rng, fs=FS, size=(1.2, 2.0)):
    """Config C5: mic 0 at the table edge + (n_mics - 1) mics uniform on the table."""
    cx, cy = 0.3, ROOM / 2
    xy = np.zeros((n_mics, 2))
    xy[0] = [cx, cy]
    xy[1:, 0] = cx + rng.uniform(0, size[0], n_mics - 1)
    xy[1:, 1] = cy + rng.uniform(-size[1] / 2, size[1] / 2, n_mics - 1)
    mic = np.concatenate([xy, np.full((n_mics, 1), MIC_HEIGHT)], axis=1)
    x1 = min(cx + SPK_RANGE_H, ROOM - WALL_KEEPOUT)
    roi = [cx + size[0] + 0.15, x1 + 0.1, WALL_KEEPOUT - 0.1, ROOM - WALL_KEEPOUT + 0.1, 0.0, 0.92]
    return Scene(mic, [float(v) for v in roi], fs)


def small_scene(n_mics=4, seed=0, fs=FS):
    """A deliberately tiny geometry (few hundred hypercubes) for fast tests."""
    rng = np.random.default_rng(seed)
    xy = np.zeros((n_mics, 2))
    xy[1:] = rng.uniform(-0.25, 0.25, (n_mics - 1, 2))
    mic = np.concatenate([xy, np.full((n_mics, 1), MIC_HEIGHT)], axis=1)
    roi = [0.6, 1.6, -0.5, 0.5, 0.0, 0.5]
    return Scene(mic, roi, fs)


def speaker_positions(scene, n_spk, rng):
    r = scene.roi
    mic = scene.mic_positions
    bx0, by0 = mic[:, 0].min() - 0.25, mic[:, 1].min() - 0.25
    bx1, by1 = mic[:, 0].max() + 0.25, mic[:, 1].max() + 0.25
    out = []
    tries = 0
    while len(out) < n_spk:
        tries += 1
        if tries > 20000:       # a region too small for n_spk sources MIN_SPEAKER_DIST apart: start over
            out, tries = [], 0
        p = np.array([rng.uniform(r[0] + 0.1, r[1] - 0.1), rng.uniform(r[2] + 0.1, r[3] - 0.1),
                      rng.uniform(max(r[4], SPK_Z[0]), min(r[5], SPK_Z[1]))])
        if bx0 < p[0] < bx1 and by0 < p[1] < by1:
            continue
        if any(np.linalg.norm(p - q) < MIN_SPEAKER_DIST for q in out):
            continue
        out.append(p)
    return np.array(out)


def burst_source(T, rng, fs=FS, sigma=0.1):
    """Band-limited Gaussian burst train: 9-tap Hann-smoothed white noise, gated on/off."""
    w = rng.standard_normal(T + 8)
    h = np.hanning(9)
    h /= np.sqrt((h ** 2).sum())
    s = np.convolve(w, h, mode="valid") * sigma
    env = np.zeros(T)
    t = int(rng.uniform(0, 0.2) * fs)
    while t < T:
        on = int(rng.uniform(0.25, 0.8) * fs)
        env[t:t + on] = rng.uniform(0.5, 1.0)
        t += on + int(rng.uniform(0.05, 0.4) * fs)
    return s * env


def mixture(scene, n_spk, T, seed, noise=1e-3, return_sources=False):
    """One (M, T) float32 mixture of ``n_spk`` sources in ``scene``."""
    rng = np.random.default_rng(seed)
    mic = scene.mic_positions
    M = mic.shape[0]
    spk = speaker_positions(scene, n_spk, rng)
    pad = int(np.ceil(12.0 / SPEED_OF_SOUND * scene.fs)) + 8
    mix = rng.standard_normal((M, T)) * noise
    for p in spk:
        s = burst_source(T + pad, rng, scene.fs)
        for m in range(M):
            d = np.linalg.norm(p - mic[m])
            k = int(round(d / SPEED_OF_SOUND * scene.fs))
            mix[m] += s[pad - k: pad - k + T] / max(d, 0.1)
    mix = mix.astype(np.float32)
    if return_sources:
        return mix, spk
    return mix


def mixtures(scene, n_spk, T, seeds, noise=1e-3):
    return np.stack([mixture(scene, n_spk, T, s, noise) for s in seeds])


def true_offsets(scene, spk):
    """TDoA (samples) of each speaker to mics 1..M-1 vs mic 0 (generate_dataset.py:505-511)."""
    mic = scene.mic_positions
    d = np.linalg.norm(spk[:, None, :] - mic[None, :, :], axis=2)
    return (d[:, 1:] - d[:, :1]) / SPEED_OF_SOUND * scene.fs
