"""Probe (NOT product code): lag-table oversampling U = 2 against the default U = 4: time of the GCC
and gather stages at C2 size (64 mixtures) and the map difference, normwise and at the maximum."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import native, synth
from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft, window_length
from acousticswarms_speech_b200.srp_phat import SRP_PHAT

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
scene = synth.desk_array(7, np.random.default_rng(1), 48000)
T, B = 144000, 64
base = torch.from_numpy(synth.mixtures(scene, 5, T, seeds=list(range(200, 208)))).to(dev)
mix = torch.cat([torch.roll(base, shifts=i, dims=2) for i in range(B // 8)], 0).contiguous()
win = window_length(T)
maps = {}
for U in (8, 4, 2):
    node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=48000, n_fft=n_fft, grid_size=0.05,
                    threshold=list(SRP_THRESHOLDS), WIDTH=8, device=dev, oversample=U)
    h = node.native
    Nw = h.num_windows(T, win)
    tabs = h.gcc(mix, win)
    out = h.gather(tabs, Nw)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    best = [1e9, 1e9]
    for _ in range(5):
        torch.cuda.synchronize()
        ev[0].record(); h.gcc(mix, win, out=tabs); ev[1].record(); h.gather(tabs, Nw, out=out); ev[2].record()
        torch.cuda.synchronize()
        best = [min(best[0], ev[0].elapsed_time(ev[1])), min(best[1], ev[1].elapsed_time(ev[2]))]
    maps[U] = h.score(mix, win).double().cpu().numpy()
    print(f"U={U}: transform (STFT + GCC tables) {best[0]*1e3:.0f} us, gather {best[1]*1e3:.0f} us, table {h.table_len()} entries per window")
ref = maps[8]
for U in (4, 2):
    d = maps[U] - ref
    print(f"U={U} vs U=8: normwise {np.linalg.norm(d) / np.linalg.norm(ref):.2e}, max |d| / max map {np.abs(d).max() / ref.max():.2e}")
