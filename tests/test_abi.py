"""CPU: libasw.so loads without a GPU and exports every symbol include/asw.h declares; argument
validation and the host-only entry points work; compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from acousticswarms_speech_b200 import _lib
from oracle import geometry_oracle, srp_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "asw.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asw_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/asw.h but not exported"
    assert sorted(_lib.SYMBOLS) == syms
    assert lib.asw_version() >= 100


def test_window_and_frame_helpers_match_reference_rules():
    lib = _lib.load()
    for T in (0, 1, 12000, 20000, 24000, 30000, 36000, 71999, 72000, 132300, 144000, 480000):
        for win in (24000, 36000):
            assert lib.asw_srp_num_windows(T, win) == len(srp_oracle.window_starts(T, win))
    assert lib.asw_srp_num_frames(36000, 2048, 512) == 67
    assert lib.asw_srp_num_frames(24000, 2048, 512) == 43
    assert lib.asw_srp_num_frames(1000, 2048, 512) == 0
    # the two readings of assumption A1 (include/asw.h, oracle/pra_stft.py)
    from oracle import pra_stft
    for win in (36000, 24000, 2048, 2049, 2048 + 512, 12288, 1000):
        x = np.zeros(win, dtype=np.float32)
        assert lib.asw_srp_num_frames_mode(win, 2048, 512, 0) == pra_stft.analysis(x, 2048, 512).shape[0]
        assert lib.asw_srp_num_frames_mode(win, 2048, 512, 1) == pra_stft.analysis(x, 2048, 512, pad_tail=True).shape[0]
    assert lib.asw_srp_num_frames_mode(36000, 2048, 512, 1) == 68
    assert lib.asw_srp_num_frames_mode(24000, 2048, 512, 1) == 44


def test_argument_validation_reports_errors():
    lib = _lib.load()
    h = ctypes.c_void_p()
    lag = np.zeros((4, 1))
    rc = lib.asw_srp_create(ctypes.byref(h), 0, 2, 4, None, 2048, 512, 2, 200, ctypes.c_float(1e-8), 0)
    assert rc == -1 and b"null" in lib.asw_last_error()
    rc = lib.asw_srp_create(ctypes.byref(h), 0, 1, 4, lag.ctypes.data, 2048, 512, 2, 200, ctypes.c_float(1e-8), 0)
    assert rc == -1
    rc = lib.asw_srp_create(ctypes.byref(h), 0, 2, 4, lag.ctypes.data, 1024, 256, 2, 200, ctypes.c_float(1e-8), 0)
    assert rc == -1 and b"nfft" in lib.asw_last_error()
    rc = lib.asw_srp_create(ctypes.byref(h), 0, 2, 4, lag.ctypes.data, 2048, 512, 2, 200, ctypes.c_float(1e-8), 3)
    assert rc == -1 and b"oversample" in lib.asw_last_error()
    assert lib.asw_shift_stack(None, None, None, 1, 1, 2, 16, None, None) == -1
    assert lib.asw_map_topk(None, 1, 1, 1, 0, None, None, None) == -1
    with pytest.raises(_lib.AswError):
        _lib.check(-1)


def test_no_cpu_fallback():
    """Without a CUDA device the scoring handle cannot be created and the wrappers refuse host tensors."""
    import torch
    from acousticswarms_speech_b200 import native
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less container")
    with pytest.raises(_lib.AswError):
        native.NativeSRP(np.zeros((4, 1)), 2)
    with pytest.raises(_lib.AswError):
        native.shift_stack(torch.zeros((1, 2, 16)), torch.zeros((1, 2), dtype=torch.int32))
    with pytest.raises(_lib.AswError):
        native.map_topk(torch.zeros((1, 8)), 2)
    lib = _lib.load()
    h = ctypes.c_void_p()
    lag = np.zeros((4, 1))
    rc = lib.asw_srp_create(ctypes.byref(h), 0, 2, 4, lag.ctypes.data, 2048, 512, 2, 200, ctypes.c_float(1e-8), 0)
    assert rc == -2 and b"cuda" in lib.asw_last_error().lower()


def test_native_cluster_walk_matches_oracle_bfs():
    """asw_geometry_cluster (host C++) vs the literal restatement of search_cluster."""
    from acousticswarms_speech_b200 import synth
    scene = synth.small_scene(n_mics=5, seed=4)
    scene.roi = [0.4, 1.4, -0.6, 0.6, 0.0, 0.5]
    geo = geometry_oracle.GeometryOracle(scene.mic_positions, scene.roi, build_fine=False)
    tree = geo.sample_tree_full
    valid = np.ascontiguousarray(tree[..., 0].astype(np.uint8))
    off = np.ascontiguousarray(tree[..., 1:].astype(np.int64))
    Lx, Ly, Lz, D = off.shape
    n = int(valid.sum())
    label = np.empty((Lx, Ly, Lz), dtype=np.int32)
    order = np.empty(n, dtype=np.int32)
    start = np.empty(n + 1, dtype=np.int32)
    ncl = ctypes.c_int32()
    lib = _lib.load()
    _lib.check(lib.asw_geometry_cluster(off.ctypes.data, valid.ctypes.data, Lx, Ly, Lz, D, label.ctypes.data,
                                        order.ctypes.data, start.ctypes.data, ctypes.byref(ncl)))
    assert ncl.value == len(geo.clusters)
    for g, c in enumerate(geo.clusters):
        mem = [list(np.unravel_index(i, (Lx, Ly, Lz))) for i in order[start[g]:start[g + 1]]]
        assert [[int(a) for a in m] for m in mem] == c[2]
    assert np.array_equal(np.where(label >= 0, label, 0), geo.POWER_INDEX)
