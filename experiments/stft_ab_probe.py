"""Probe (NOT product code): time the transform stage (fused STFT + PHAT + pair products, then the GCC tables) and print
a hash of the lag tables, so that STFT kernel variants can be compared across processes:
    for v in classic rr; do ASW_STFT=$v python experiments/stft_ab_probe.py; done"""
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import synth
from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft, window_length
from acousticswarms_speech_b200.srp_phat import SRP_PHAT

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)


def run(tag, M, T, B, fs=48000):
    scene = synth.desk_array(M, np.random.default_rng(1), fs)
    node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=fs, n_fft=n_fft, grid_size=0.05,
                    threshold=list(SRP_THRESHOLDS), WIDTH=8, device=dev)
    h = node.native
    if os.environ.get('STFT_PATH'):
        h.set_stft_path(os.environ['STFT_PATH'])      # 'split': FFT + PHAT to global memory, then the pair-product kernel
    nb = min(B, 4)
    base = torch.from_numpy(synth.mixtures(scene, 3, T, seeds=list(range(200, 200 + nb)))).to(dev)
    mix = torch.cat([torch.roll(base, shifts=i, dims=2) for i in range((B + nb - 1) // nb)], 0)[:B].contiguous()
    win = window_length(T)
    tabs = h.gcc(mix, win)
    best = 1e9
    for _ in range(7):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); h.gcc(mix, win, out=tabs); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    sha = hashlib.sha256(tabs.cpu().numpy().tobytes()).hexdigest()[:16]
    print(f"{os.environ.get('ASW_STFT', 'rr'):7s} {os.environ.get('STFT_PATH', 'auto'):5s} {tag}: M={M} B={B} T={T} transform {best * 1e3:.1f} us  tables sha {sha}", flush=True)


run("C2 B=64", 7, 144000, 64)
run("C2 B=1 ", 7, 144000, 1)
for M in (2, 3, 4, 5, 6, 8):
    run(f"M={M} B=16", M, 96000, 16)
run("44.1k B=3", 7, 132300, 3, fs=44100)
