"""Oracle: SRP-PHAT scoring of TDoA hypercubes -- staged numpy restatement.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows sep/Traditional_SP/SRP_Prunning.py:
  * steering table (``generate_mod_vector`` + pair products) .. :221-243, 368-381
  * analysis-window framing ................................... :393-403
  * STFT (assumption A1, oracle/pra_stft.py) .................. :404-409
  * per-channel PHAT .......................................... :414-416
  * cross-spectra, upper triangle, row-major pair order ....... :418-426
  * steering contraction, /F/P ................................ :428-429
  * max over windows (map starts at zeros) .................... :253, 430-433
and sep/Mic_Array.py:160-163 for the analysis-window length rule.

Dtypes follow the reference: STFT complex64, PHAT/cross-spectra complex64,
steering table float64, map float64.  The (G,F,P) table is evaluated in
G-chunks (identical arithmetic per element).
"""
import numpy as np

from . import pra_stft


def pair_list(M):
    """Row-major upper triangle i<j -- the order ``mask_triu`` selects (:223-229)."""
    return [(i, j) for i in range(M) for j in range(i + 1, M)]


def window_length(T):
    """sep/Mic_Array.py:160-163."""
    return 36000 if T >= 72000 else 24000


def window_starts(T, window):
    """:393-403  ->  list of window start samples."""
    step = window // 2
    out = []
    for j in range(T // step - 1):
        if j * step + window > T:
            break
        out.append(j * step)
    return out


def steering_dist(grids, mic_pos):
    """:369-376  distance used for steering; mic height is IGNORED (quirk R11)."""
    px, py, pz = grids[None, :, 0], grids[None, :, 1], grids[None, :, 2]
    mx, my = mic_pos[:, None, 0], mic_pos[:, None, 1]
    return np.sqrt((px - mx) ** 2 + (py - my) ** 2 + pz ** 2)          # (M, G)


def mode_vector(grids, mic_pos, freq_bins, fs, nfft, C=343.0):
    """:368-381  mode[f, m, g] = exp(1j * omega_f * dist[m, g] / C)."""
    dist = steering_dist(grids, mic_pos) / C
    omega = 2 * np.pi * fs * np.asarray(freq_bins) / nfft
    return np.exp(1j * omega[:, None, None] * dist[None])


def steering_table_chunk(grids, mic_pos, freq_bins, fs, nfft, C=343.0):
    """:227-229 for a chunk of grids -> (g, F, P) complex128."""
    mode = mode_vector(grids, mic_pos, freq_bins, fs, nfft, C)        # (F, M, g)
    mid = np.moveaxis(mode, 2, 0)                                      # (g, F, M)
    prs = pair_list(mic_pos.shape[0])
    i = np.array([p[0] for p in prs])
    j = np.array([p[1] for p in prs])
    return mid[:, :, i] * np.conj(mid[:, :, j])


def stft_window(seg, nfft, hop, pad_tail=False):
    """:404-409  (M, nfft//2+1, n_frames) complex64."""
    return np.array([pra_stft.analysis(x, nfft, hop, pad_tail=pad_tail).T for x in seg])


def phat(X, tol=1e-8):
    """:414-416 in complex64/float32 like torch."""
    a = np.abs(X).astype(np.float32)
    a[a < tol] = tol
    return (X / a).astype(np.complex64)


def cross_spectra(pX, freq_bins):
    """:418-426  -> CC_flat (F, P) complex64."""
    M = pX.shape[0]
    n = pX.shape[2]
    prs = pair_list(M)
    out = np.empty((len(freq_bins), len(prs)), dtype=np.complex64)
    for a, k in enumerate(freq_bins):
        A = pX[:, k, :]
        cc = (A @ np.conj(A).T) / np.float32(n)
        out[a] = [cc[i, j] for (i, j) in prs]
    return out


def contract(CC_flat, grids, mic_pos, freq_bins, fs, nfft, C=343.0, chunk=2048):
    """:428-429  map_w[g] = sum_{f,p}(Re CC Re tab - Im CC Im tab)/F/P  (float64)."""
    G = grids.shape[0]
    F, P = CC_flat.shape
    cr = CC_flat.real.astype(np.float64)
    ci = CC_flat.imag.astype(np.float64)
    out = np.empty(G)
    for s in range(0, G, chunk):
        tab = steering_table_chunk(grids[s:s + chunk], mic_pos, freq_bins, fs, nfft, C)
        r = cr[None] * tab.real - ci[None] * tab.imag
        out[s:s + chunk] = r.sum((1, 2)) / F / P
    return out


def score(signal, grids, mic_pos, freq_bins, fs, nfft, window=None, tol=1e-8,
          C=343.0, stages=False, pad_tail=False):
    """SRP_Map_WINDOW_torch (:387-433).  Returns the float64 map (G,) and, with
    ``stages=True``, the per-window intermediates."""
    signal = np.asarray(signal)
    M, T = signal.shape
    if window is None:
        window = window_length(T)
    srp = np.zeros(grids.shape[0])
    st = {"starts": [], "CC": [], "map_w": []}
    for s0 in window_starts(T, window):
        seg = signal[:, s0:s0 + window]
        X = stft_window(seg, nfft, nfft // 4, pad_tail)
        pX = phat(X, tol)
        CC = cross_spectra(pX, freq_bins)
        mw = contract(CC, grids, mic_pos, freq_bins, fs, nfft, C)
        srp = np.maximum(srp, mw)
        if stages:
            st["starts"].append(s0)
            st["CC"].append(CC)
            st["map_w"].append(mw)
    if stages:
        return srp, st
    return srp


def pair_lags(grids, mic_pos, fs, C=343.0):
    """Fractional pair lags (samples): tau[g, p] = fs (d_i(g) - d_j(g)) / C.
    Not in the reference as such -- it is the phase slope of the table (:379,
    :228): tab[g,f,p] = exp(2*pi*i * k_f * tau[g,p] / nfft)."""
    d = steering_dist(grids, mic_pos)                                   # (M, G)
    prs = pair_list(mic_pos.shape[0])
    return np.stack([fs * (d[i] - d[j]) / C for (i, j) in prs], axis=1)  # (G, P)
