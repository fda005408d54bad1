"""Probe (NOT product code): error of the device SRP map against the reference's float64 golden map (desk scene)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from acousticswarms_speech_b200 import synth
from acousticswarms_speech_b200.mic_array import Mic_Array

g = np.load(os.path.join(ROOT, "tests", "golden", "desk_scene.npz"))
scene = synth.Scene(g["mic_positions"], list(g["roi"]), int(g["fs"]))
mix = synth.mixture(scene, int(g["n_spk"]), int(g["T"]), int(g["seed"]))
ma = Mic_Array(scene.mic_positions, Spk_Range=scene.roi)
ma.Apply_SRP_PHAT(torch.from_numpy(mix))
got, ref = ma.SRP_node.SRP_map.cpu().numpy(), g["srp_map"]
print(f"G={ref.shape[0]} max|err|/max(ref) = {np.abs(got-ref).max()/ref.max():.3e}; "
      f"normwise = {np.linalg.norm(got-ref)/np.linalg.norm(ref):.3e}; bar 1e-4")
