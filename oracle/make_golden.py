"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) under the stubs
of oracle/ref_loader.py.  Build-container only.

    python -m oracle.make_golden

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference ships no golden vectors or
tests, so these fixtures -- outputs of the reference's own classes on seeded synthetic inputs -- are what
pins the oracle restatements (tests/test_oracle_golden.py) and, through them, the CUDA path.  The one
link the fixtures cannot pin is assumption A1 (pyroomacoustics' STFT, oracle/pra_stft.py): the reference
is run here with that same restatement in place of the absent package.
"""
import copy
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def patches_to_arrays(patches):
    return {
        "offsets": np.array([p.sample_offset for p in patches], dtype=np.int64).reshape(len(patches), -1),
        "widths": np.array([p.width_list for p in patches], dtype=np.int64).reshape(len(patches), -1),
        "area_sizes": np.array([p.area_size() for p in patches], dtype=np.int64),
        "area_sha": np.array([sha(p.area_points) if p.area_points is not None else "" for p in patches]),
        "peak_pos": np.array([p.peak_pos if p.peak_pos is not None else [np.nan] * 3 for p in patches]).reshape(
            len(patches), 3),
    }


class DelayAndSumSpot:
    """Deterministic stand-in for the spot network (weights are external to the reference repo):
    same interface as DataParallelSpotModel.shift_and_sep, output = mean over mics of the shifted mix."""

    def __init__(self, ns):
        self.ns = ns

    def shift_and_sep(self, input_channels, patch_list, Strict=0, save_input=False):
        out = np.zeros((len(patch_list), input_channels.shape[-1]), dtype=np.float32)
        for j, p in enumerate(patch_list):
            sh = -torch.Tensor([0, *p.sample_offset]).unsqueeze(1)
            sh = torch.round(sh).long()
            data = self.ns.joint_network.roll_by_gather(input_channels, 1, sh)
            out[j] = data.mean(0).numpy() * (1.0 if Strict == 0 else 0.5)
        return out


def scene_fixture(ns, name, scene, n_spk, T, seed, store_mix, store_fine=True):
    from acousticswarms_speech_b200 import synth
    mix = synth.mixture(scene, n_spk, T, seed)
    MA = ns.Mic_Array.Mic_Array(scene.mic_positions, Spk_Range=scene.roi)
    node = MA.SRP_node
    patches, _ = MA.Apply_SRP_PHAT(torch.tensor(mix))
    peaks = node.find_valid_peak_new()
    out = {
        "mic_positions": scene.mic_positions, "roi": np.array(scene.roi), "fs": scene.fs, "n_spk": n_spk, "T": T,
        "seed": seed, "mix_sha": sha(mix),
        "grids": node.grids, "cluster_offsets": np.array([c.sample_offset for c in node.clusters], dtype=np.int32),
        "cluster_sizes": np.array([c.cluster_size() for c in node.clusters], dtype=np.int32),
        "cluster_first_index": np.array([c.index[0] for c in node.clusters], dtype=np.int32),
        "power_index_sha": sha(node.POWER_INDEX.astype(np.int64)),
        "dis_matrix": node.dis_matrix,
        "srp_map": node.SRP_map.numpy().astype(np.float64), "max_power": node.MAX_POWER, "min_power": node.Min_POWER,
        "power_map_sha": sha(node.POWER_MAP),
        "peaks": np.array(peaks, dtype=np.int64),
        "upper_bound_pairwise": MA.upper_bound_pairwise,
    }
    for k, v in patches_to_arrays(patches).items():
        out["patch_" + k] = v
    if store_mix:
        out["mix"] = mix
    # coarse stage with the stand-in spot model, then the fine patch list of the first survivors
    spot = DelayAndSumSpot(ns)
    kept = MA.Spotform_Big_Patch(torch.tensor(mix), copy.deepcopy(patches), spot)
    out["big_kept_offsets"] = np.array([p.sample_offset for p in kept], dtype=np.int64).reshape(len(kept), -1)
    out["relative_threshold"] = MA.Relative_Threshold
    fine_off, fine_w, fine_idx = [], [], [0]
    for c in copy.deepcopy(kept[:3]):
        fine = ns.local_utils_3d.search_area([c], scene.mic_positions, MA.upper_bound_pairwise)
        fine_off += [p.sample_offset for p in fine]
        fine_w += [p.width_list for p in fine]
        fine_idx.append(len(fine_off))
    D = scene.mic_positions.shape[0] - 1
    out["fine_offsets"] = np.array(fine_off, dtype=np.int64).reshape(-1, D)
    out["fine_widths"] = np.array(fine_w, dtype=np.int64).reshape(-1, D)
    out["fine_index"] = np.array(fine_idx)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: G={node.grids.shape[0]} peaks={len(peaks)} patches={len(patches)} kept={len(kept)} "
          f"fine={len(fine_off)} max={node.MAX_POWER:.4f}")


def shift_fixture(ns):
    rng = np.random.default_rng(42)
    M, T = 5, 997
    mix = rng.standard_normal((M, T)).astype(np.float32)
    offs = np.array([[0, 0, 0, 0], [3, -7, 250, -250], [996, -996, 997, -997], [1000, -1994, 2.5, -3.5],
                     [0.5, 1.5, -0.5, -1.5]], dtype=np.float64)
    outs = []
    for o in offs:
        sh = -torch.Tensor([0, *o]).unsqueeze(1)
        sh = torch.round(sh).long()
        outs.append(ns.joint_network.roll_by_gather(torch.tensor(mix), 1, sh).numpy())
    data = torch.tensor(np.stack(outs))
    dn, mu, sd = ns.spot_network.normalize_input(data)
    np.savez_compressed(os.path.join(GOLDEN, "shift.npz"), mix=mix, offsets=offs, shifted=np.stack(outs),
                        norm=dn.numpy(), means=mu.numpy(), stds=sd.numpy())
    print("shift: ok")


AUDIO_DECIMATION = 97


def pcm_content(x):
    """16-bit PCM content carried as float32: normalize_input's re-quantisation is then the identity, so a device
    spot model whose network is the delay-and-sum stand-in reproduces ``DelayAndSumSpot`` up to float32 rounding."""
    return (np.clip(np.rint(x * 32768.0), -32768, 32767) / 32768.0).astype(np.float32)


def spotform_fixture(ns, name, scene, n_spk, T, seed, max_candidates):
    """The whole post-pruning chain of the reference with the stand-in separator: Spotform_Big_Patch ->
    Spotform_Small_Patch_Parallel (sep/Mic_Array.py:196-395); what the latter returns is the fixture."""
    from acousticswarms_speech_b200 import synth
    mix = pcm_content(synth.mixture(scene, n_spk, T, seed))
    MA = ns.Mic_Array.Mic_Array(scene.mic_positions, Spk_Range=scene.roi)
    patches, _ = MA.Apply_SRP_PHAT(torch.tensor(mix))
    spot = DelayAndSumSpot(ns)
    kept = MA.Spotform_Big_Patch(torch.tensor(mix), copy.deepcopy(patches), spot)
    cands = copy.deepcopy(kept[:max_candidates])
    pairs = MA.Spotform_Small_Patch_Parallel(torch.tensor(mix), cands, spot)
    D = scene.mic_positions.shape[0] - 1
    centres = [pc.center_pos() for pc, *_ in pairs]
    # Clustering_new (sep/Mic_Array.py:399-500) on those outputs.  Its split_wav calls librosa, which is absent here: the
    # reference runs with oracle/fake_librosa.py installed as eval_utils' librosa (everything else is the reference's).
    from oracle import fake_librosa
    sys.modules["sep.helpers.eval_utils"].librosa = fake_librosa
    audio_final, patch_final, spot_times, _ = MA.Clustering_new([(pc, a.copy(), p, t, o, l) for pc, a, p, t, o, l in pairs])
    out = {
        "final_tags": np.array([p[3] for p in patch_final]), "final_spot_times": spot_times,
        "final_audio_sha": np.array([sha(np.asarray(a, dtype=np.float32)) for a in audio_final]),
        "mic_positions": scene.mic_positions, "roi": np.array(scene.roi), "fs": scene.fs, "n_spk": n_spk, "T": T,
        "seed": seed, "mix_sha": sha(mix), "max_candidates": max_candidates,
        "patch_offsets": np.array([p.sample_offset for p in patches], dtype=np.int64).reshape(len(patches), D),
        "patch_widths": np.array([p.width_list for p in patches], dtype=np.int64).reshape(len(patches), D),
        "kept_offsets": np.array([p.sample_offset for p in kept], dtype=np.int64).reshape(len(kept), D),
        "relative_threshold": MA.Relative_Threshold,
        "spotforming_times": MA.spotforming_times,
        "cands_after_offsets": np.array([c.sample_offset for c in cands], dtype=np.int64).reshape(len(cands), D),
        "cands_after_widths": np.array([c.width_list for c in cands], dtype=np.int64).reshape(len(cands), D),
        "tags": np.array([t for _, _, _, t, _, _ in pairs]),
        "powers": np.array([p for _, _, p, _, _, _ in pairs], dtype=np.float64),
        "audio_offsets": np.array([o["audio_offset"] for *_, o, _ in pairs], dtype=np.int64).reshape(len(pairs), D),
        "localization_offsets": np.array([o["localization_offset"] for *_, o, _ in pairs], dtype=np.float64).reshape(len(pairs), D),
        "centres": np.array([c if c is not None else [np.nan] * 3 for c in centres], dtype=np.float64).reshape(len(pairs), 3),
        "centre_area_sizes": np.array([pc.area_size() for pc, *_ in pairs], dtype=np.int64),
        "centre_is_peak": np.array([pc.peak_pos is not None for pc, *_ in pairs]),
        "labels": np.array([l for *_, l in pairs], dtype=np.int64),
        "audio_dec": np.array([a[::AUDIO_DECIMATION] for _, a, *_ in pairs], dtype=np.float32).reshape(len(pairs), -1),
        "audio_sha": np.array([sha(np.asarray(a, dtype=np.float32)) for _, a, *_ in pairs]),
    }
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: patches={len(patches)} kept={len(kept)} candidates={len(cands)} fine={MA.spotforming_times} "
          f"outputs={len(pairs)} tags={list(out['tags'])} final={list(out['final_tags'])}")


def main():
    from acousticswarms_speech_b200 import synth
    from oracle import ref_loader
    os.makedirs(GOLDEN, exist_ok=True)
    ns = ref_loader.load()
    what = sys.argv[1:] or ["scenes", "spotform"]
    if "scenes" in what:
        shift_fixture(ns)
        scene_fixture(ns, "small_scene", synth.small_scene(n_mics=4, seed=2), 2, 36000, 7, store_mix=True)
        scene_fixture(ns, "desk_scene", synth.desk_array(7, np.random.default_rng(0)), 3, 144000, 0, store_mix=False)
    if "spotform" in what:      # python -m oracle.make_golden spotform   (leaves the fixtures above untouched)
        spotform_fixture(ns, "small_spotform", synth.small_scene(n_mics=4, seed=2), 2, 36000, 7, max_candidates=6)
        spotform_fixture(ns, "desk_spotform", synth.desk_array(7, np.random.default_rng(0)), 3, 144000, 0, max_candidates=8)


if __name__ == "__main__":
    main()
