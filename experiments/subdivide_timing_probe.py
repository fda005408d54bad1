"""Probe (NOT product code): device search_area (asw_subdivide) vs the host mirror, 25 coarse candidates."""
import copy, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import local_utils, synth
from acousticswarms_speech_b200.mic_array import Mic_Array
from acousticswarms_speech_b200.pipeline import FrontEnd

scene = synth.desk_array(7, np.random.default_rng(1), 48000)
ma = Mic_Array(scene.mic_positions, Spk_Range=scene.roi)
fe = FrontEnd(ma.SRP_node)
mix = synth.mixtures(scene, 5, 144000, seeds=[100])
cands = fe.prune(fe.score(torch.from_numpy(mix).cuda())[0])[0][:25]
for c in cands:
    c.area_points
t = time.perf_counter(); host = [local_utils.search_area([c], scene.mic_positions, ma.upper_bound_pairwise) for c in copy.deepcopy(cands)]
t_host = time.perf_counter() - t
ma._search_area_device(copy.deepcopy(cands)); torch.cuda.synchronize()
t = time.perf_counter(); dev = ma._search_area_device(copy.deepcopy(cands)); torch.cuda.synchronize()
t_dev = time.perf_counter() - t
print(f"candidates {len(cands)} -> fine patches {sum(len(h) for h in host)}; host {t_host*1e3:.1f} ms, device (incl. D2H + Patch objects) {t_dev*1e3:.1f} ms")
