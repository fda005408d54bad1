"""Probe (NOT product code): asw_subdivide on the 10 candidates with the most member voxels of a 16-mixture group
(the kernel's tail) -- launched alone so that an ncu source view shows where a single large candidate spends its time."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import native, synth
from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft
from acousticswarms_speech_b200.pipeline import FrontEnd
from acousticswarms_speech_b200.srp_phat import SRP_PHAT

dev = torch.device("cuda", 0)
scene = synth.desk_array(7, np.random.default_rng(1), 48000)
node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=48000, n_fft=n_fft, grid_size=0.05,
                threshold=list(SRP_THRESHOLDS), WIDTH=8, device=dev)
fe = FrontEnd(node, dev)
x = torch.from_numpy(synth.mixtures(scene, 5, 144000, seeds=[100 + b for b in range(16)])).to(dev)
n_sel, off, wid, pk = fe.select(fe.score(x)[0])
B, P, D = off.shape
valid = torch.arange(P, device=dev, dtype=torch.int32)[None, :] < n_sel.clamp(max=P)[:, None]
w = (wid * valid).reshape(-1).to(torch.int32)
cands, cw = off.reshape(-1, D)[w > 0].contiguous(), w[w > 0].contiguous()
out = native.subdivide(node.native_select, cands, cw, fe.upper_bound_pairwise(), member_cap=1 << 17)
rc = np.array([len(m) for m in out[-1]])
idx = torch.from_numpy(np.ascontiguousarray(np.argsort(-rc)[:10])).to(dev)
c2, w2 = cands[idx].contiguous(), cw[idx].contiguous()
for _ in range(3):
    native.subdivide_device(node.native_select, c2, w2, fe.upper_bound_pairwise())
torch.cuda.synchronize()
print("members", sorted(rc)[-10:], "leaves", out[0][idx.cpu().numpy()])
