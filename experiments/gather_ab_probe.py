"""Probe (NOT product code): time asw_srp_gather alone at C2 (64 mixtures and 1 mixture, 7 mics) and C5 (4 mixtures,
16 mics) and print a hash of the map, so that kernel variants can be compared across processes:
    for v in legacy bulk ws; do ASW_GATHER=$v python experiments/gather_ab_probe.py; done"""
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import synth
from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft, window_length
from acousticswarms_speech_b200.srp_phat import SRP_PHAT

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)


def run(tag, scene, T, B, nspk, **grid):
    node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=48000, n_fft=n_fft,
                    threshold=list(SRP_THRESHOLDS), WIDTH=8, device=dev, **grid)
    h = node.native
    M = scene.mic_positions.shape[0]
    nb = min(B, 4)
    base = torch.from_numpy(synth.mixtures(scene, nspk, T, seeds=list(range(200, 200 + nb)))).to(dev)
    mix = torch.cat([torch.roll(base, shifts=i, dims=2) for i in range((B + nb - 1) // nb)], 0)[:B].contiguous()
    win = window_length(T)
    Nw = h.num_windows(T, win)
    tabs = h.gcc(mix, win)
    out = h.gather(tabs, Nw)
    best = 1e9
    for _ in range(7):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); h.gather(tabs, Nw, out=out); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    sha = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
    print(f"{os.environ.get('ASW_GATHER', 'ws'):7s} {tag}: M={M} G={node.grids.shape[0]} B={B} Nw={Nw} gather {best * 1e3:.1f} us  map sha {sha}", flush=True)


run("C2 B=64", synth.desk_array(7, np.random.default_rng(1), 48000), 144000, 64, 5, grid_size=0.05)
run("C2 B=1 ", synth.desk_array(7, np.random.default_rng(1), 48000), 144000, 1, 5, grid_size=0.05)
run("C5 B=4 ", synth.table_array(16, np.random.default_rng(16)), 480000, 4, 8, grid_size=0.025, grid_size_z=0.05)
