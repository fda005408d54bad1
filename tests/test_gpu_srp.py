"""GPU parity: SRP-PHAT scoring stages vs the oracle restatement of
SRP_PHAT.SRP_Map_WINDOW_torch (sep/Traditional_SP/SRP_Prunning.py:387-433).
Floating point: 1e-4 relative in fp32 (north_star), measured normwise."""
import numpy as np
import pytest
import torch

from acousticswarms_speech_b200 import synth
from acousticswarms_speech_b200.constants import freq_bins, n_fft
from oracle import geometry_oracle, srp_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def small():
    scene = synth.small_scene(n_mics=4, seed=2)
    geo = geometry_oracle.GeometryOracle(scene.mic_positions, scene.roi, build_fine=False)
    return scene, geo


def _native(scene, grids, **kw):
    from acousticswarms_speech_b200 import native
    lag = native.pair_lags(grids, scene.mic_positions, scene.fs, 343.0)
    return native.NativeSRP(lag, scene.mic_positions.shape[0], **kw)


@pytest.mark.parametrize("T", [48000, 96000])
def test_stage_parity_small(cuda_device, small, T):
    scene, geo = small
    mix = synth.mixture(scene, 2, T, seed=7)
    srp = _native(scene, geo.grids)
    win = srp_oracle.window_length(T)
    got = srp.score(torch.from_numpy(mix).cuda(), win)
    torch.cuda.synchronize()
    want, st = srp_oracle.score(mix, geo.grids, scene.mic_positions, freq_bins, scene.fs, n_fft, stages=True)
    # cross-spectra (:418-426)
    cc = srp.read_cc().cpu().numpy()[0]
    ref_cc = np.stack(st["CC"])
    assert cc.shape == ref_cc.shape
    assert np.abs(cc - ref_cc).max() <= TOL * np.abs(ref_cc).max()
    # GCC lag tables against a direct float64 evaluation of the same band-limited transform
    lo, n, off, tl, U = srp.gcc_layout()
    gcc = srp.read_gcc().cpu().numpy()[0]
    Nw = ref_cc.shape[0]
    F, P = ref_cc.shape[1:]
    worst = 0.0
    scale = 0.0
    for p in range(P):
        lags = lo[p] + np.arange(n[p]) / U
        ph = np.exp(2j * np.pi * np.outer(lags, freq_bins) / n_fft)
        npad = (n[p] + 3) // 4 * 4
        for w in range(Nw):
            ref = (ph @ ref_cc[w, :, p].astype(np.complex128)).real / (F * P)
            seg = gcc[Nw * off[p] + w * npad: Nw * off[p] + w * npad + n[p]]
            worst = max(worst, np.abs(seg - ref).max())
            scale = max(scale, np.abs(ref).max())
    assert worst <= TOL * scale * P      # P pair terms add into one map value
    # the map (:428-430)
    g = got.cpu().numpy()[0]
    assert np.abs(g - want).max() <= TOL * want.max()
    assert (g >= 0).all()


def test_batch_matches_single(cuda_device, small):
    scene, geo = small
    T = 72000
    mixes = synth.mixtures(scene, 2, T, seeds=[1, 2, 3])
    srp = _native(scene, geo.grids)
    win = srp_oracle.window_length(T)
    both = srp.score(torch.from_numpy(mixes).cuda(), win).cpu().numpy()
    for b in range(3):
        one = srp.score(torch.from_numpy(mixes[b]).cuda(), win).cpu().numpy()[0]
        assert np.array_equal(one, both[b])        # deterministic: no atomics on the scoring path
        want = srp_oracle.score(mixes[b], geo.grids, scene.mic_positions, freq_bins, scene.fs, n_fft)
        assert np.abs(both[b] - want).max() <= TOL * want.max()


def test_short_signal_gives_zero_map(cuda_device, small):
    scene, geo = small
    srp = _native(scene, geo.grids)
    mix = synth.mixture(scene, 1, 30000, seed=1)       # T//step - 1 == 1 but one 24000 window fits
    got = srp.score(torch.from_numpy(mix).cuda(), 24000).cpu().numpy()[0]
    want = srp_oracle.score(mix, geo.grids, scene.mic_positions, freq_bins, scene.fs, n_fft, window=24000)
    assert np.abs(got - want).max() <= TOL * max(want.max(), 1e-12)
    mix = mix[:, :20000]                               # no window fits: map stays zero (:253)
    got = srp.score(torch.from_numpy(mix).cuda(), 24000).cpu().numpy()[0]
    assert (got == 0).all()


@pytest.mark.parametrize("n_mics", [3, 7, 8])
@pytest.mark.parametrize("win", [2048, 2600, 3100, 6200])
def test_few_frames_per_window(cuda_device, n_mics, win):
    """Windows of 1, 2, 3 and 9 STFT frames: fewer (frame, mic) FFTs than the eight warps of the round-robin kernel, a
    last round with idle warps, one partial frame group; M = 8 takes the warp-per-mic kernel."""
    scene = synth.small_scene(n_mics=n_mics, seed=3)
    geo = geometry_oracle.GeometryOracle(scene.mic_positions, scene.roi, build_fine=False)
    srp = _native(scene, geo.grids)
    T = 4 * win
    mix = synth.mixture(scene, 2, T, seed=5)
    assert srp.num_frames(win) == (win - n_fft) // (n_fft // 4) + 1
    got = srp.score(torch.from_numpy(mix).cuda(), win).cpu().numpy()[0]
    want = srp_oracle.score(mix, geo.grids, scene.mic_positions, freq_bins, scene.fs, n_fft, window=win)
    assert want.max() > 0
    assert np.abs(got - want).max() <= TOL * want.max()


def test_silence_and_phat_floor(cuda_device, small):
    scene, geo = small
    srp = _native(scene, geo.grids)
    mix = np.zeros((4, 48000), dtype=np.float32)       # |X| < tol everywhere -> pX = 0 -> map = 0
    got = srp.score(torch.from_numpy(mix).cuda(), 24000).cpu().numpy()[0]
    assert (got == 0).all()


@pytest.mark.parametrize("U", [2, 8])
def test_oversampling_variants(cuda_device, small, U):
    scene, geo = small
    T = 48000
    mix = synth.mixture(scene, 2, T, seed=9)
    srp = _native(scene, geo.grids, oversample=U)
    got = srp.score(torch.from_numpy(mix).cuda(), 24000).cpu().numpy()[0]
    want = srp_oracle.score(mix, geo.grids, scene.mic_positions, freq_bins, scene.fs, n_fft)
    assert np.abs(got - want).max() <= TOL * want.max()


def test_generic_mic_count(cuda_device):
    """M > 8 takes the generic cross-spectrum path and multi-group table staging."""
    rng = np.random.default_rng(4)
    scene = synth.table_array(10, rng)
    scene.roi = [2.0, 2.6, 3.2, 3.8, 0.0, 0.4]
    geo = geometry_oracle.GeometryOracle(scene.mic_positions, scene.roi, build_fine=False)
    mix = synth.mixture(scene, 2, 48000, seed=3)
    srp = _native(scene, geo.grids)
    got = srp.score(torch.from_numpy(mix).cuda(), 24000).cpu().numpy()[0]
    want = srp_oracle.score(mix, geo.grids, scene.mic_positions, freq_bins, scene.fs, n_fft)
    assert np.abs(got - want).max() <= TOL * want.max()


def test_max_mic_count(cuda_device):
    """32 mics = 496 pairs: the split STFT in four chunks of eight mics, 20 pair tiles, gather stages of more than 32
    pairs (the producer warp's second lane round) and the full stage-descriptor table."""
    scene = synth.table_array(32, np.random.default_rng(9))
    scene.roi = [2.0, 2.6, 3.2, 3.8, 0.0, 0.4]                 # where the speakers stand
    geo = geometry_oracle.GeometryOracle(scene.mic_positions, [2.1, 2.4, 3.3, 3.6, 0.0, 0.2], build_fine=False)
    mix = synth.mixture(scene, 2, 48000, seed=3)
    srp = _native(scene, geo.grids)
    got = srp.score(torch.from_numpy(mix).cuda(), 24000).cpu().numpy()[0]
    want = srp_oracle.score(mix, geo.grids, scene.mic_positions, freq_bins, scene.fs, n_fft)
    assert np.abs(got - want).max() <= TOL * want.max()


@pytest.mark.parametrize("n_mics,T", [(4, 48000), (4, 96000), (4, 50003), (10, 48000)])
def test_tail_padded_frame_mode(cuda_device, small, n_mics, T):
    """ASW_FRAMES_PAD_TAIL (the other reading of assumption A1): ceil((win - nfft) / hop) + 1 frames, the ragged
    last frame zero padded -- warp kernel (M = 4, also with rows that are not 16-byte aligned) and generic
    kernel (M = 10) against the oracle with the same convention; and it is a different map from the default."""
    if n_mics == 4:
        scene, geo = small
    else:
        scene = synth.table_array(10, np.random.default_rng(4))
        scene.roi = [2.0, 2.6, 3.2, 3.8, 0.0, 0.4]
        geo = geometry_oracle.GeometryOracle(scene.mic_positions, scene.roi, build_fine=False)
    mix = synth.mixture(scene, 2, T, seed=3)
    win = srp_oracle.window_length(T)
    srp = _native(scene, geo.grids, pad_tail=True)
    assert srp.num_frames(win) == -(-(win - n_fft) // (n_fft // 4)) + 1 == srp_oracle.stft_window(
        mix[:1, :win], n_fft, n_fft // 4, pad_tail=True).shape[2]
    x = torch.from_numpy(mix).cuda()
    got = srp.score(x, win).cpu().numpy()[0]
    want, st = srp_oracle.score(mix, geo.grids, scene.mic_positions, freq_bins, scene.fs, n_fft, stages=True,
                                pad_tail=True)
    cc = srp.read_cc().cpu().numpy()[0]
    ref_cc = np.stack(st["CC"])
    assert np.abs(cc - ref_cc).max() <= TOL * np.abs(ref_cc).max()
    assert np.abs(got - want).max() <= TOL * want.max()
    srp.set_pad_tail(False)
    base = srp.score(x, win).cpu().numpy()[0]
    want0 = srp_oracle.score(mix, geo.grids, scene.mic_positions, freq_bins, scene.fs, n_fft)
    assert np.abs(base - want0).max() <= TOL * want0.max()
    assert np.abs(base - got).max() > 10 * TOL * want.max()
    # a window that the frames tile exactly: both conventions are the same computation
    w2 = n_fft + 20 * (n_fft // 4)
    a = srp.score(x, w2)
    srp.set_pad_tail(True)
    assert srp.num_frames(w2) == 21
    assert torch.equal(a, srp.score(x, w2))


def test_split_stft_path(cuda_device, small):
    """ASW_STFT_SPLIT (spectra through global memory + pair-product kernel): bit-identical to the fused register
    kernel for M = 4 (same products, same adds, same frame order), and within tolerance of the oracle and of the
    generic kernel for M = 10 where AUTO selects it."""
    scene, geo = small
    x = torch.from_numpy(synth.mixtures(scene, 2, 96000, seeds=[5, 6])).cuda()
    srp = _native(scene, geo.grids)
    fused = srp.score(x, 36000).clone()
    cc_fused = srp.read_cc().clone()
    srp.set_stft_path("split")
    split = srp.score(x, 36000).clone()
    assert torch.equal(srp.read_cc(), cc_fused) and torch.equal(split, fused)
    srp.set_pad_tail(True)                                  # the ragged last frame through the split path too
    a = srp.score(x, 36000).clone()
    srp.set_stft_path("auto")
    assert torch.equal(a, srp.score(x, 36000))

    scene = synth.table_array(10, np.random.default_rng(4))
    scene.roi = [2.0, 2.6, 3.2, 3.8, 0.0, 0.4]
    geo = geometry_oracle.GeometryOracle(scene.mic_positions, scene.roi, build_fine=False)
    mix = synth.mixture(scene, 2, 48000, seed=3)
    srp = _native(scene, geo.grids)                         # AUTO -> split for 10 mics
    got = srp.score(torch.from_numpy(mix).cuda(), 24000).cpu().numpy()[0]
    want = srp_oracle.score(mix, geo.grids, scene.mic_positions, freq_bins, scene.fs, n_fft)
    assert np.abs(got - want).max() <= TOL * want.max()
    srp.set_stft_path("generic")
    gen = srp.score(torch.from_numpy(mix).cuda(), 24000).cpu().numpy()[0]
    assert np.abs(gen - want).max() <= TOL * want.max()
    assert np.abs(gen - got).max() <= 1e-5 * want.max()


def test_topk(cuda_device):
    from acousticswarms_speech_b200 import native
    rng = np.random.default_rng(0)
    B, G = 5, 20315
    m = rng.random((B, G)).astype(np.float32)
    m[m < 0.27] = 0.0                      # ~27 % exact zeros like a real map
    m[0, 100] = m[0, 7] = m[0].max()       # exact ties -> lower index first
    for K in (1, 30, 128, 1000):
        val, idx = native.map_topk(torch.from_numpy(m).cuda(), K, idx_offset=1000)
        val, idx = val.cpu().numpy(), idx.cpu().numpy() - 1000
        for b in range(B):
            order = np.lexsort((np.arange(G), -m[b]))[:K]
            assert np.array_equal(idx[b], order)
            assert np.array_equal(val[b], m[b][order])
    # K > G pads with (-inf, -1)
    val, idx = native.map_topk(torch.from_numpy(m[:, :10].copy()).cuda(), 16)
    assert (idx.cpu().numpy()[:, 10:] == -1).all() and np.isneginf(val.cpu().numpy()[:, 10:]).all()


@pytest.mark.parametrize("T", [48001, 48002, 50003])
def test_misaligned_rows(cuda_device, small, T):
    """T not a multiple of 4: odd channels start at 4- or 8-byte alignment (cp.async 4/8-byte prefetch path)."""
    scene, geo = small
    mix = synth.mixture(scene, 2, T, seed=13)
    srp = _native(scene, geo.grids)
    got = srp.score(torch.from_numpy(mix).cuda(), 24000).cpu().numpy()[0]
    want = srp_oracle.score(mix, geo.grids, scene.mic_positions, freq_bins, scene.fs, n_fft)
    assert np.abs(got - want).max() <= TOL * want.max()


def test_fs_44100(cuda_device):
    """BASELINE.json quotes 44.1 kHz; the reference hard-codes 48 kHz.  fs is a parameter here."""
    scene = synth.small_scene(n_mics=4, seed=6, fs=44100)
    geo = geometry_oracle.GeometryOracle(scene.mic_positions, scene.roi, FS=44100, build_fine=False)
    T = 132300
    mix = synth.mixture(scene, 2, T, seed=3)
    srp = _native(scene, geo.grids)
    got = srp.score(torch.from_numpy(mix).cuda(), srp_oracle.window_length(T)).cpu().numpy()[0]
    want = srp_oracle.score(mix, geo.grids, scene.mic_positions, freq_bins, 44100, n_fft)
    assert len(srp_oracle.window_starts(T, 36000)) == 6
    assert np.abs(got - want).max() <= TOL * want.max()


def test_kernel_variants_bit_identical(cuda_device, tmp_path):
    """The scoring kernels have measured alternatives kept for A/B runs: the fused STFT with a block barrier per frame
    (ASW_STFT=classic) vs the round-robin / mbarrier kernel, and the gather whose warp 0 stages and gathers
    (ASW_GATHER=bulk) or the cp.async ring (legacy) vs the warp-specialised kernel.  Every variant does the same
    arithmetic in the same order: maps and cross-spectra must be equal bit for bit (3 .. 8 mics, ragged tail, batches)."""
    import os
    import subprocess
    import sys
    from _variant_worker import variant_maps
    mine = variant_maps()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for env_add in ({"ASW_STFT": "classic", "ASW_GATHER": "bulk"}, {"ASW_STFT": "rr", "ASW_GATHER": "legacy"}):
        path = str(tmp_path / ("variant_" + env_add["ASW_STFT"] + ".npz"))
        env = dict(os.environ, **env_add)
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        r = subprocess.run([sys.executable, os.path.join(root, "tests", "_variant_worker.py"), path], env=env, cwd=root,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        other = np.load(path)
        assert set(other.files) == set(mine)
        for k in mine:
            assert np.array_equal(mine[k], other[k]), (env_add, k)
