"""Checks shared by the CPU (oracle) and GPU (device path) tests against tests/golden/*_spotform.npz."""
import numpy as np


def pcm_content(x):
    """16-bit PCM content carried as float32 (the fixtures' mixtures: normalize_input's re-quantisation is the identity)."""
    return (np.clip(np.rint(x * 32768.0), -32768, 32767) / 32768.0).astype(np.float32)


def check_spotform_pairs(pairs, g, power_rtol=1e-6, offset_atol=1e-9, centre_atol=1e-9):
    """The list a Spotform_Small_Patch_Parallel implementation returns (sep/Mic_Array.py:376-392) vs what the
    reference's own method returned for the fixture: tags (candidate _ cluster head), separation and localisation
    offsets, powers, merged centres, the audio rows (decimated) and the labels."""
    assert [t for _, _, _, t, _, _ in pairs] == [str(t) for t in g["tags"]]
    D = g["audio_offsets"].shape[1]
    assert np.array_equal(np.array([o["audio_offset"] for *_, o, _ in pairs]).reshape(-1, D), g["audio_offsets"])
    loc = np.array([o["localization_offset"] for *_, o, _ in pairs], dtype=np.float64).reshape(-1, D)
    assert np.abs(loc - g["localization_offsets"]).max() <= offset_atol, np.abs(loc - g["localization_offsets"]).max()
    pw = np.array([p for _, _, p, _, _, _ in pairs], dtype=np.float64)
    assert np.abs(pw - g["powers"]).max() <= power_rtol * g["powers"].max()
    assert [int(l) for *_, l in pairs] == list(g["labels"])
    for k, (pc, audio, *_rest) in enumerate(pairs):
        assert (pc.peak_pos is not None) == bool(g["centre_is_peak"][k])
        assert pc.area_size() == int(g["centre_area_sizes"][k])
        c = pc.center_pos()
        assert (c is None) == bool(np.isnan(g["centres"][k]).all())
        if c is not None:
            assert np.abs(np.asarray(c) - g["centres"][k]).max() <= centre_atol
        dec = np.asarray(audio)[::97]
        assert np.abs(dec - g["audio_dec"][k]).max() <= power_rtol * np.abs(g["audio_dec"][k]).max() + 1e-12
