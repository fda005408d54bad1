"""Probe (NOT product code): cProfile of Spotform_Small_Patch_Parallel with a stand-in separator."""
import copy, cProfile, os, pstats, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import synth
from acousticswarms_speech_b200.mic_array import Mic_Array
from acousticswarms_speech_b200.spot import DataParallelSpotModel
from fine_stage_probe import Net  # noqa: E402  (runs the probe once as a warm-up)

scene = synth.desk_array(7, np.random.default_rng(1), 48000)
ma = Mic_Array(scene.mic_positions, Spk_Range=scene.roi)
spot = DataParallelSpotModel(Net(), batch_size=128)
mix = torch.from_numpy(synth.mixture(scene, 5, 144000, seed=100))
patches, _ = ma.Apply_SRP_PHAT(mix)
kept = ma.Spotform_Big_Patch(mix, copy.deepcopy(patches), spot)
ma.Spotform_Small_Patch_Parallel(mix, copy.deepcopy(kept), spot)
pr = cProfile.Profile()
cands = copy.deepcopy(kept)
pr.enable()
ma.Spotform_Small_Patch_Parallel(mix, cands, spot)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
