"""Probe (NOT product code): per-kernel time of one C5-scale mixture (16 mics, 10 s, dense grid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import constants, synth
from acousticswarms_speech_b200.srp_phat import SRP_PHAT

rng = np.random.default_rng(16)
scene = synth.table_array(16, rng)
t0 = time.perf_counter()
node = SRP_PHAT(scene.mic_positions, constants.freq_bins, scene.roi, FS=48000, n_fft=constants.n_fft,
                grid_size=0.025, grid_size_z=0.05, threshold=list(constants.SRP_THRESHOLDS), WIDTH=8)
print(f"setup {time.perf_counter()-t0:.1f} s, G = {node.grids.shape[0]}, P = 120")
T = 480000
for B in (1, 4):
    mix = torch.from_numpy(synth.mixtures(scene, 4, T, seeds=list(range(8, 8 + B)))).cuda()
    for _ in range(2):
        node.native.score(mix, 36000)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        node.native.score(mix, 36000)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"B={B}: score {ms:.2f} ms per call = {B * node.grids.shape[0] / ms * 1e3:.3e} hypercubes/s")
