"""Assumption A1 -- restatement of ``pyroomacoustics.transform.stft.analysis``.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference calls ``pra.transform.stft.analysis(x, nfft, nfft // 4).T`` at
sep/Traditional_SP/SRP_Prunning.py:406 with pyroomacoustics==0.5.0
(requirements.txt:10).  That package is neither vendored under /root/reference
nor installable here, so its published behaviour is restated:

* ``win=None``  -> rectangular analysis window (no taper, no scaling);
* no zero padding, no centring;
* ``n = (len(x) - L) // hop + 1`` frames, frame ``i`` = ``x[i*hop : i*hop+L]``;
* one-sided real FFT of length ``L`` -> ``L//2 + 1`` bins, returned ``(n, L//2+1)``;
* for float32 input the result is ``complex64`` (numpy 1.23, pinned by the
  reference, evaluates the FFT in double precision and pyroomacoustics stores
  it into a complex64 buffer).

**PARITY UNPINNED** for this function: no reference test, fixture or golden
vector pins this boundary.  The only internal evidence is that the caller
indexes the transposed result as ``(bin, frame)`` (SRP_Prunning.py:422) and
uses ``shape[2]`` as the frame count (:419).  Everything depending on A1 lives
in this one function so it can be re-pinned in one place.
"""
import numpy as np


def analysis(x, L, hop, win=None, zp_back=0, zp_front=0, pad_tail=False):
    """``pad_tail=False`` is assumption A1 as stated above.  ``pad_tail=True`` is the alternative reading of the
    same boundary -- a non-streaming STFT that keeps a ragged tail: ``ceil((len - L) / hop) + 1`` frames, the input
    extended with zeros to fill the last one (one zero-padded frame when ``len < L``).  libasw.so implements both
    (``asw_srp_set_frame_mode``); they coincide when ``(len - L) % hop == 0``."""
    if win is not None or zp_back or zp_front:
        raise NotImplementedError("only the call shape used by the reference is restated")
    x = np.asarray(x)
    if x.ndim != 1:
        raise NotImplementedError("mono input only (the reference loops over channels)")
    ctype = np.complex64 if x.dtype == np.float32 else np.complex128
    if pad_tail and x.shape[0] > 0:
        n = 1 if x.shape[0] < L else -(-(x.shape[0] - L) // hop) + 1
        extra = (n - 1) * hop + L - x.shape[0]
        if extra > 0:
            x = np.concatenate([x, np.zeros(extra, dtype=x.dtype)])
    else:
        n = (x.shape[0] - L) // hop + 1
    if n <= 0:
        return np.zeros((0, L // 2 + 1), dtype=ctype)
    idx = np.arange(L)[None, :] + hop * np.arange(n)[:, None]
    frames = x[idx].astype(np.float64)
    X = np.fft.rfft(frames, n=L, axis=1)
    return X.astype(ctype)
