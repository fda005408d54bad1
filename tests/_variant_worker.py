"""Child process of tests/test_gpu_srp.py::test_kernel_variants_bit_identical: the kernel-variant switches (ASW_STFT,
ASW_GATHER) are read once per process, so the maps of a fixed set of scenes are computed here under the parent's
environment and written to an .npz the parent compares with its own."""
import sys

import numpy as np
import torch

from acousticswarms_speech_b200 import native, synth
from oracle import geometry_oracle


def variant_maps():
    out = {}
    for n_mics, T, pad_tail, B in ((3, 48000, False, 2), (4, 50003, True, 1), (6, 96000, False, 3), (7, 96000, False, 5),
                                   (8, 48000, False, 2), (32, 48000, False, 2)):
        scene = synth.small_scene(n_mics=n_mics, seed=2)
        roi = scene.roi if n_mics <= 8 else [0.6, 0.9, -0.2, 0.1, 0.0, 0.2]     # 496 pairs: keep the grid small
        geo = geometry_oracle.GeometryOracle(scene.mic_positions, roi, build_fine=False)
        lag = native.pair_lags(geo.grids, scene.mic_positions, scene.fs, 343.0)
        srp = native.NativeSRP(lag, n_mics, pad_tail=pad_tail)
        mix = np.stack([synth.mixture(scene, 2, T, seed=11 + b) for b in range(B)])
        win = T // 2 if T % 2 == 0 else (T - 1) // 2
        got = srp.score(torch.from_numpy(mix).cuda(), win)
        torch.cuda.synchronize()
        out[f"map_{n_mics}_{T}"] = got.cpu().numpy()
        out[f"cc_{n_mics}_{T}"] = srp.read_cc().cpu().numpy()
    return out


if __name__ == "__main__":
    np.savez(sys.argv[1], **variant_maps())
