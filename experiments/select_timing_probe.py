"""Probe (NOT product code): where does asw_select_patches spend its time?"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import synth
from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft
from acousticswarms_speech_b200.pipeline import FrontEnd
from acousticswarms_speech_b200.srp_phat import SRP_PHAT

dev = torch.device("cuda", 0)
scene = synth.desk_array(7, np.random.default_rng(1), 48000)
node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=48000, n_fft=n_fft, grid_size=0.05,
                threshold=list(SRP_THRESHOLDS), WIDTH=8, device=dev)
fe = FrontEnd(node, dev)
B = 32
mix = torch.from_numpy(synth.mixtures(scene, 5, 144000, seeds=[100 + b for b in range(B)])).to(dev)
smap, _, _ = fe.score(mix)
peaks, count, _ = node.native_peaks.find(smap)
sel = node.native_select
print("peaks per mixture:", count.cpu().numpy().tolist())


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for k in (1, 5, 20, 60, 1000):
    c = count.clamp(max=k)
    n, off, wid, pk = sel.select(smap, peaks, c)
    print(f"k={k:5d}  select {t(lambda: sel.select(smap, peaks, c)):8.1f} us   patches/mixture mean {n.float().mean().item():.1f}")
one = smap[:1].contiguous(); p1 = peaks[:1].contiguous(); c1 = count[:1].contiguous()
print("B=1 select", t(lambda: sel.select(one, p1, c1)), "us")

