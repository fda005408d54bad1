"""Probe (NOT product code): where the host time of Spotform_Small_Patch_Parallel goes (25 coarse candidates of one
C2 mixture, stand-in separator)."""
import copy, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import synth
from acousticswarms_speech_b200.mic_array import Mic_Array
from acousticswarms_speech_b200.spot import DataParallelSpotModel


class Net(torch.nn.Module):
    def forward(self, x, cond):
        return x.mean(1, keepdim=True) * (cond[:, 1:2] + 2 * cond[:, 0:1]).unsqueeze(-1)


class HostPowersOnly:
    def __init__(self, spot):
        self.shift_and_sep = spot.shift_and_sep


scene = synth.desk_array(7, np.random.default_rng(1), 48000)
ma = Mic_Array(scene.mic_positions, Spk_Range=scene.roi)
spot = DataParallelSpotModel(Net(), batch_size=128)
mix = torch.from_numpy(synth.mixture(scene, 5, 144000, seed=100))
patches, _ = ma.Apply_SRP_PHAT(mix)
kept = ma.Spotform_Big_Patch(mix, copy.deepcopy(patches), spot)
print(f"coarse patches {len(patches)}, kept {len(kept)}")
for label, model in (("device powers + leaf centres", spot), ("host loops (reference style)", HostPowersOnly(spot))):
    for rep in range(2):
        cands = copy.deepcopy(kept)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        total, index, _, _ = ma.small_patch_list(copy.deepcopy(cands))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        out = ma.Spotform_Small_Patch_Parallel(mix, cands, model)
        torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{label}: fine patches {len(total)}; small_patch_list {1e3*(t1-t0):.1f} ms; whole Spotform_Small_Patch_Parallel "
          f"{1e3*(t2-t1):.1f} ms; outputs {len(out)}")
t0 = time.perf_counter(); ma.Spotform_Big_Patch(mix, copy.deepcopy(patches), spot); torch.cuda.synchronize()
print(f"Spotform_Big_Patch (device powers): {1e3*(time.perf_counter()-t0):.1f} ms")
t0 = time.perf_counter(); ma.Spotform_Big_Patch(mix, copy.deepcopy(patches), HostPowersOnly(spot)); torch.cuda.synchronize()
print(f"Spotform_Big_Patch (host loop):     {1e3*(time.perf_counter()-t0):.1f} ms")
