"""CPU: the product's host-side mirrors (geometry, pruning, subdivision, shift bookkeeping, sharding)
against the oracle and the reference's golden vectors.  No CUDA compute is called."""
import copy
import hashlib
import os

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from acousticswarms_speech_b200 import constants, dist, local_utils, native, synth
from acousticswarms_speech_b200.mic_array import find_merge_center, weight_mean_pos
from acousticswarms_speech_b200.patch import Patch
from acousticswarms_speech_b200.srp_phat import SRP_PHAT
from oracle import geometry_oracle, prune_oracle, shift_oracle, srp_oracle, subdivide_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def host_node(mic, roi, fs=48000):
    return SRP_PHAT(mic, constants.freq_bins, roi, FS=fs, n_fft=constants.n_fft, grid_size=0.05,
                    threshold=list(constants.SRP_THRESHOLDS), WIDTH=constants.INIT_WIDTH, build_native=False)


class DelayAndSum:
    def shift_and_sep(self, mix, patch_list, Strict=0, save_input=False):
        out = shift_oracle.shift_stack(np.asarray(mix), [p.sample_offset for p in patch_list]).mean(1)
        return (out * (1.0 if Strict == 0 else 0.5)).astype(np.float32)


@pytest.mark.parametrize("name", ["small_scene", "desk_scene"])
def test_host_geometry_and_pruning_match_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    node = host_node(g["mic_positions"], list(g["roi"]))
    assert np.array_equal(node.grids, g["grids"])
    assert np.array_equal(np.array([c.sample_offset for c in node.clusters]), g["cluster_offsets"])
    assert np.array_equal(np.array([c.cluster_size() for c in node.clusters]), g["cluster_sizes"])
    assert sha(node.POWER_INDEX.astype(np.int64)) == str(g["power_index_sha"])
    assert np.array_equal(node.dis_matrix, g["dis_matrix"])
    node.load_map(g["srp_map"])
    assert sha(node.POWER_MAP) == str(g["power_map_sha"])
    assert node.find_valid_peak_new() == [int(i) for i in g["peaks"]]
    patches = node.local_source_adaptive()
    assert np.array_equal(np.array([p.sample_offset for p in patches]).reshape(len(patches), -1), g["patch_offsets"])
    assert np.array_equal(np.array([p.width_list for p in patches]).reshape(len(patches), -1), g["patch_widths"])
    assert [sha(p.area_points) for p in patches] == [str(s) for s in g["patch_area_sha"]]
    # coarse selection + subdivision through the product's local_utils with the stand-in spot model
    if "mix" in g:
        mix = g["mix"]
    else:
        scene = synth.Scene(g["mic_positions"], list(g["roi"]), int(g["fs"]))
        mix = synth.mixture(scene, int(g["n_spk"]), int(g["T"]), int(g["seed"]))
        if sha(mix) != str(g["mix_sha"]):
            return
    kept, _, thr = local_utils.binary_search_baseline(mix, DelayAndSum(), patches, g["mic_positions"])
    assert np.array_equal(np.array([p.sample_offset for p in kept]).reshape(len(kept), -1), g["big_kept_offsets"])
    off, wid = [], []
    for c in copy.deepcopy(kept[:3]):
        fine = local_utils.search_area([c], g["mic_positions"], g["upper_bound_pairwise"])
        off += [p.sample_offset for p in fine]
        wid += [p.width_list for p in fine]
    D = g["mic_positions"].shape[0] - 1
    assert np.array_equal(np.array(off).reshape(-1, D), g["fine_offsets"])
    assert np.array_equal(np.array(wid).reshape(-1, D), g["fine_widths"])


def test_scoring_without_handle_raises():
    scene = synth.small_scene(4, 2)
    node = host_node(scene.mic_positions, scene.roi)
    from acousticswarms_speech_b200._lib import AswError
    with pytest.raises(AswError):
        node.SRP_Map_WINDOW_new(np.zeros((4, 48000), dtype=np.float32), window=24000)


def test_pair_lags_are_the_table_phase_slope():
    scene = synth.small_scene(4, 3)
    geo = geometry_oracle.GeometryOracle(scene.mic_positions, scene.roi, build_fine=False)
    lag = native.pair_lags(geo.grids, scene.mic_positions, 48000, 343.0)
    assert np.array_equal(lag, srp_oracle.pair_lags(geo.grids, scene.mic_positions, 48000))
    tab = srp_oracle.steering_table_chunk(geo.grids[:5], scene.mic_positions, constants.freq_bins, 48000, 2048)
    k = constants.freq_bins[None, :, None]
    assert np.abs(tab - np.exp(2j * np.pi * k * lag[:5, None, :] / 2048)).max() < 1e-11


@settings(max_examples=200, deadline=None)
@given(st.lists(st.floats(-5000, 5000, allow_nan=False, width=32), min_size=1, max_size=6))
def test_offsets_to_shifts_rounds_like_torch(offs):
    got = native.offsets_to_shifts(np.array(offs))[0]
    assert got[0] == 0
    assert np.array_equal(got, shift_oracle.shift_indices(offs).astype(np.int32))


def test_offsets_to_shifts_half_integers():
    assert native.offsets_to_shifts(np.array([[0.5, 1.5, 2.5, -0.5, -1.5, -2.5]])).tolist() == [[0, 0, 2, 2, 0, -2, -2]]


def test_patch_check_out_truncates_like_reference():
    for off, w, ub in [([30, -41, 7], [8, 8, 8], [20.0, 20.0, 20.0]), ([101, -3, 55], [8, 4, 8], [60.5, 10.0, 50.2])]:
        a = Patch(np.array(off), np.array(w), None)
        b = prune_oracle.Patch(np.array(off), np.array(w), None)
        a.check_out(np.array(ub))
        b.check_out(np.array(ub))
        assert np.array_equal(a.sample_offset, b.sample_offset) and np.array_equal(a.width_list, b.width_list)
        assert a.sample_offset.dtype == np.int64


def test_max_avg_power_and_si_sdr():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(30000).astype(np.float32)
    x[:10000] *= 0.01
    p, seg = local_utils.max_avg_power(x)
    q, seg2 = subdivide_oracle.max_avg_power(x)
    assert p == q and np.array_equal(seg, seg2)
    y = x + 0.1 * rng.standard_normal(30000).astype(np.float32)
    assert 15 < local_utils.si_sdr(y, x) < 25


def test_si_sdr_from_gram_matches_direct():
    """The Gram-matrix form used by the device fine stage (spot.si_sdr_from_gram) against eval_utils' definition,
    around the -4 dB clustering threshold and far from it; and Patch.center_pos with a device-supplied centre."""
    from acousticswarms_speech_b200.patch import Patch
    from acousticswarms_speech_b200.spot import si_sdr_from_gram
    rng = np.random.default_rng(5)
    ref = rng.standard_normal(144000).astype(np.float32)
    for noise in (0.01, 0.3, 1.0, 1.6, 3.0, 30.0):
        est = (0.7 * ref + noise * rng.standard_normal(ref.shape[0])).astype(np.float32)
        want = local_utils.si_sdr(est, ref)
        e, r = est.astype(np.float64), ref.astype(np.float64)
        got = si_sdr_from_gram(e @ e, r @ r, e @ r)
        assert abs(got - want) < 1e-3, (noise, got, want)
    p = Patch(np.zeros(3, dtype=np.int64), np.full(3, 4), None, None, area_fn=lambda: np.ones((3, 5)),
              centre=np.array([1.0, 2.0, 3.0]))
    assert np.array_equal(p.center_pos(), [1.0, 2.0, 3.0]) and p._area_points is None
    assert p.area_size() == 5 and np.array_equal(p.area_points_getter()(), np.ones((3, 5)))


def test_weight_mean_and_merge_center():
    mic = synth.small_scene(4, 2).mic_positions
    pts = np.stack(np.meshgrid(np.linspace(0.8, 1.2, 9), np.linspace(-0.2, 0.2, 9), [0.3], indexing="ij"), 0).reshape(3, -1)
    ps = [Patch(np.array([4, 0, -4]), [4, 4, 4], pts[:, :40]), Patch(np.array([8, 0, -4]), [4, 4, 4], pts[:, 40:])]
    pos, off = weight_mean_pos(ps, [1.0, 0.5], [0, 1])
    assert np.allclose(off, (np.array([4, 0, -4]) * 1.0) / 1.0)          # second is below 0.75 x max: ignored
    c = find_merge_center(off, pts, mic, np.array([1.0, 0.0, 0.3]))
    assert c.center_pos() is not None


def test_shard_ranges_cover_and_balance():
    for n in (0, 1, 7, 256, 20315):
        for world in (1, 2, 3, 4, 8):
            spans = [dist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_merge_topk_order_and_ties():
    import torch
    vals = torch.tensor([[0.5, 0.9, 0.9, -1.0, 0.1, 0.9]])
    idxs = torch.tensor([[7, 30, 2, -1, 5, 11]], dtype=torch.int32)
    v, i = dist.merge_topk(vals, idxs, 4)
    assert i.tolist() == [[2, 11, 30, 7]] and v.tolist()[0][:3] == [pytest.approx(0.9)] * 3


def test_merge_topk_matches_a_naive_merge():
    """Random candidate lists with many ties, padding and negative values against a lexicographic sort."""
    import torch
    rng = np.random.default_rng(21)
    for _ in range(20):
        B, n, K = 3, 40, 9
        vals = rng.choice([-1.5, -0.25, 0.0, 0.25, 0.5, 2.0], size=(B, n)).astype(np.float32)
        idxs = np.stack([rng.permutation(200)[:n] for _ in range(B)]).astype(np.int32)
        idxs[rng.random((B, n)) < 0.2] = -1
        v, i = dist.merge_topk(torch.from_numpy(vals), torch.from_numpy(idxs), K)
        for b in range(B):
            cand = [(-float(vals[b, k]), int(idxs[b, k])) for k in range(n) if idxs[b, k] >= 0]
            cand.sort()
            want = cand[:K]
            got = [(-float(a), int(c)) for a, c in zip(v[b], i[b]) if c >= 0]
            assert got == want[:len(got)] and len(got) == min(K, len(cand))


def test_window_constants():
    assert constants.window_length(144000) == 36000 and constants.window_length(60000) == 24000
    assert constants.frames_per_window(36000) == 67 and constants.frames_per_window(24000) == 43
    assert constants.window_starts(144000, 36000) == srp_oracle.window_starts(144000, 36000)


def test_fine_stage_tdoa_rows_follow_the_array_rate():
    """The fine stage works in the rate the hypercube table was built for (Mic_Array.fs), not in the reference's
    constant 48 kHz: at 44.1 kHz the point TDoAs used by search_area / the leaf builders / find_merge_center must be
    the node's own Offset_1 volume, and the 48 kHz default must stay what the reference computes."""
    from acousticswarms_speech_b200.mic_array import find_merge_center
    scene = synth.small_scene(n_mics=4, seed=2)
    for fs in (44100, 48000):
        node = host_node(scene.mic_positions, scene.roi, fs=fs)
        pos1, off1 = node.Pos_1.reshape(-1, 3), node.Offset_1.reshape(-1, 3)
        pick = np.random.default_rng(fs).choice(pos1.shape[0], 500, replace=False)
        rows = local_utils._tdoa_rows(pos1[pick].T, scene.mic_positions, fs)
        assert np.array_equal(rows.T, off1[pick])
        if fs == 48000:
            assert np.array_equal(local_utils._tdoa_rows(pos1[pick].T, scene.mic_positions), rows)
            assert np.array_equal(subdivide_oracle.point_tdoas(pos1[pick].T, scene.mic_positions), rows)
        # a merged centre taken from real voxels of that rate is found again inside its own area
        centre = off1[pick[:40]].mean(0)
        area = pos1[pick].T
        pc = find_merge_center(np.round(off1[pick[0]]), area, scene.mic_positions, np.zeros(3), fs)
        inside = np.all(np.abs(off1[pick] - np.round(off1[pick[0]])) <= 1.5 + 1e-3, axis=1)
        assert pc.area_points is not None and pc.area_points.shape[1] == int(inside.sum()) >= 1


def test_corr_table_lag_range_from_geometry():
    """CorrTables.lag_for_geometry: the largest pair TDoA of the array plus the hypercube width, in 32-sample steps,
    capped at the kernel's 512; every coarse patch of a geometry must fit (|r_c' - r_c| <= lag for all pairs)."""
    scene = synth.desk_array(7, np.random.default_rng(1), 48000)
    L = native.CorrTables.lag_for_geometry(scene.mic_positions, 48000)
    d = np.sqrt(((scene.mic_positions[:, None] - scene.mic_positions[None]) ** 2).sum(-1)).max()
    assert L % 32 == 0 and 32 <= L <= 512 and L >= d / 343.0 * 48000 + 8
    assert native.CorrTables.lag_for_geometry(scene.mic_positions, 44100) <= L
    far = np.array([[0.0, 0.0, 0.0], [9.0, 0.0, 0.0]])
    assert native.CorrTables.lag_for_geometry(far, 48000) == 512            # beyond the table: those patches take the exact pass
    node = host_node(scene.mic_positions, scene.roi)
    offs = np.array([c.sample_offset for c in node.clusters])              # every hypercube centre the geometry can produce
    r = np.concatenate([np.zeros((offs.shape[0], 1)), offs], axis=1)
    assert np.abs(r[:, :, None] - r[:, None, :]).max() + 8 <= L
