"""Build-container only: the oracle restatements against the LIVE reference (/root/reference) on a
scene that is not one of the committed fixtures."""
import copy

import numpy as np
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.reference
if not ref_loader.available():
    pytest.skip("reference not present (GPU box)", allow_module_level=True)

from acousticswarms_speech_b200 import synth  # noqa: E402
from oracle import geometry_oracle, prune_oracle, shift_oracle, srp_oracle, subdivide_oracle  # noqa: E402


@pytest.fixture(scope="module")
def live():
    ns = ref_loader.load()
    scene = synth.small_scene(n_mics=5, seed=11)
    scene.roi = [0.5, 1.9, -0.7, 0.7, 0.0, 0.6]
    mix = synth.mixture(scene, 2, 72000, seed=5)
    MA = ns.Mic_Array.Mic_Array(scene.mic_positions, Spk_Range=scene.roi)
    patches, _ = MA.Apply_SRP_PHAT(torch.tensor(mix))
    return ns, scene, mix, MA, patches


def test_geometry_scoring_pruning(live):
    ns, scene, mix, MA, patches = live
    node = MA.SRP_node
    geo = geometry_oracle.GeometryOracle(scene.mic_positions, scene.roi)
    assert np.array_equal(geo.grids, node.grids)
    assert all(np.array_equal(a[0], b.sample_offset) and a[2] == b.index for a, b in zip(geo.clusters, node.clusters))
    assert np.array_equal(geo.Offset_1, node.Offset_1) and np.array_equal(geo.Offset_5, node.Offset_5)
    m = srp_oracle.score(mix, geo.grids, scene.mic_positions, np.arange(2, 200), 48000, 2048)
    ref = node.SRP_map.numpy()
    assert np.abs(m - ref).max() <= 1e-7 * ref.max()
    pm, pi = prune_oracle.fill_powermap(ref, geo.clusters, (geo.Lx, geo.Ly, geo.Lz))
    assert np.array_equal(pm, node.POWER_MAP) and np.array_equal(pi, node.POWER_INDEX)
    peaks = prune_oracle.find_valid_peaks(pm, pi, geo.dis_matrix, node.MAX_POWER, len(geo.clusters))
    assert peaks == [int(i) for i in node.find_valid_peak_new()]
    mine = prune_oracle.local_source_adaptive(ref, peaks, geo.grids, [c[0] for c in geo.clusters], 5, geo)
    assert len(mine) == len(patches)
    for a, b in zip(mine, patches):
        assert np.array_equal(a.sample_offset, b.sample_offset) and np.array_equal(a.width_list, b.width_list)
        assert np.array_equal(a.area_points, b.area_points)


def test_shift_and_subdivision(live):
    ns, scene, mix, MA, patches = live
    mt = torch.tensor(mix)
    for p in patches[:4]:
        sh = torch.round(-torch.Tensor([0, *p.sample_offset]).unsqueeze(1)).long()
        ref = ns.joint_network.roll_by_gather(mt, 1, sh).numpy()
        assert np.array_equal(shift_oracle.shift_stack(mix, [p.sample_offset])[0], ref)
    if patches:
        a = ns.local_utils_3d.search_area([copy.deepcopy(patches[0])], scene.mic_positions, MA.upper_bound_pairwise)
        mine = prune_oracle.Patch(patches[0].sample_offset.copy(), patches[0].width_list.copy(),
                                  patches[0].area_points, patches[0].peak_pos)
        b = subdivide_oracle.search_area([mine], scene.mic_positions, subdivide_oracle.upper_bounds(scene.mic_positions))
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert np.array_equal(x.sample_offset, y.sample_offset) and np.array_equal(x.width_list, y.width_list)
