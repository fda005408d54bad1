"""Probe (NOT product code): latency of the drop-in Mic_Array.Apply_SRP_PHAT on ONE mixture (BASELINE config C1:
7 mics, 3 s @ 48 kHz, 3 speakers), host tensor in, Patch list out -- the call the reference makes per sample."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import synth
from acousticswarms_speech_b200.mic_array import Mic_Array

scene = synth.desk_array(7, np.random.default_rng(0), 48000)
t = time.perf_counter(); ma = Mic_Array(scene.mic_positions, Spk_Range=scene.roi); t_setup = time.perf_counter() - t
mix = torch.from_numpy(synth.mixture(scene, 3, 144000, 0))
ma.Apply_SRP_PHAT(mix); torch.cuda.synchronize()      # first call builds the selection handle (uploads the 1 cm volume)
ts = []
for i in range(20):
    t = time.perf_counter(); patches, _ = ma.Apply_SRP_PHAT(mix); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
print(f"G={ma.SRP_node.grids.shape[0]} setup {t_setup:.2f} s; Apply_SRP_PHAT median {np.median(ts)*1e3:.2f} ms, min {min(ts)*1e3:.2f} ms, patches {len(patches)}")
t = time.perf_counter(); [p.area_points for p in patches]; print(f"lazy area_points for all {len(patches)} patches: {(time.perf_counter()-t)*1e3:.1f} ms")
