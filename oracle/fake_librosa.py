"""A deterministic stand-in for the two librosa calls of ``split_wav`` (sep/helpers/eval_utils.py:43-70).

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``librosa`` is a third-party dependency of the reference that is absent from this image (like pyroomacoustics): the
reference's ``Clustering_new`` (sep/Mic_Array.py:399-500) segments every output with ``librosa.feature.rms`` and
``librosa.effects.split``.  The package's mirror imports the real librosa, exactly as the reference does.  To pin
everything ELSE of ``Clustering_new`` -- the SI-SDR non-maximum suppression, the window-wise checks, the distance
rule, the returned tuple -- the golden fixtures are produced by running the unmodified reference with THIS module
installed as its ``librosa`` (oracle/make_golden.py), and the tests install the same module for the mirror and the
oracle.  It is NOT a restatement of librosa: frames are not centred or padded, dB are taken against the maximum.
"""
import types

import numpy as np


def _rms(y, frame_length=2048, hop_length=512):
    y = np.asarray(y, dtype=np.float64)
    n = max(1, 1 + (len(y) - frame_length) // hop_length) if len(y) >= frame_length else 1
    out = np.zeros((1, n))
    for i in range(n):
        seg = y[i * hop_length:i * hop_length + frame_length]
        out[0, i] = np.sqrt(np.mean(seg ** 2)) if len(seg) else 0.0
    return out


def _split(y, top_db=60, ref=np.max, frame_length=2048, hop_length=512):
    r = _rms(y, frame_length, hop_length)[0]
    ref_value = ref(r) if callable(ref) else float(ref)
    db = 20.0 * np.log10(np.maximum(r, 1e-10) / max(ref_value, 1e-10))
    on = db > -top_db
    edges = np.flatnonzero(np.diff(np.concatenate([[0], on.astype(int), [0]])))
    out = [[int(s * hop_length), int(min(len(y), e * hop_length))] for s, e in zip(edges[0::2], edges[1::2])]
    return np.array(out, dtype=np.int64).reshape(-1, 2)


feature = types.SimpleNamespace(rms=lambda y=None, frame_length=2048, hop_length=512, **kw: _rms(y, frame_length, hop_length))
effects = types.SimpleNamespace(split=_split)
