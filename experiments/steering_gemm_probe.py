"""Probe (NOT product code): is the frequency-domain steering contraction worth a hand-written tcgen05
kernel?  north_star: "Tensor cores are used only if a frequency-domain steering formulation is benchmarked
as a true dense contraction."  The reference's formulation (SRP_Prunning.py:428-429) is the GEMM
    map[g, (b,w)] = sum_{f,p} Re(tab[g,f,p]) Re(CC[b,w,f,p]) - Im(tab) Im(CC)      (G x 2FP) . (2FP x B*Nw)
This script times that GEMM with the LIBRARY (torch.matmul, bf16 / tf32 / bf16x3 split) against the
lag-table path (gcc_kernel + srp_gather_kernel) on the bench workload and measures its error, so the
decision is made on numbers.  Results are recorded in DESIGN.md section 9."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from acousticswarms_speech_b200 import native, synth  # noqa: E402
from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft  # noqa: E402
from acousticswarms_speech_b200.srp_phat import SRP_PHAT  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda", 0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    scene = synth.desk_array(7, np.random.default_rng(1), 48000)
    node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=48000, n_fft=n_fft, grid_size=0.05,
                    threshold=list(SRP_THRESHOLDS), WIDTH=8, device=dev)
    h = node.native
    G, F, P = h.G, h.F, h.P
    mix = torch.from_numpy(synth.mixtures(scene, 5, 144000, seeds=range(100, 100 + B))).to(dev)
    ref_map = h.score(mix, 36000).clone()
    cc = h.read_cc()                                             # (B, Nw, F, P) complex64
    Nw = cc.shape[1]
    lag = torch.from_numpy(native.pair_lags(node.grids, scene.mic_positions, 48000, 343.0)).to(dev)   # (G, P) f64
    k = torch.arange(2, 200, device=dev, dtype=torch.float64)
    ph = 2 * np.pi * lag[:, None, :] * k[None, :, None] / 2048.0                                        # (G, F, P)
    A = torch.cat([torch.cos(ph).reshape(G, F * P), -torch.sin(ph).reshape(G, F * P)], 1)              # (G, 2FP) f64
    Bm = torch.cat([cc.real.reshape(B * Nw, F * P), cc.imag.reshape(B * Nw, F * P)], 1).t().contiguous()  # (2FP, B*Nw)
    scale = 1.0 / (F * P)

    def finish(out):
        return torch.clamp(out.view(G, B, Nw).amax(2), min=0).t() * scale

    exact = finish(A @ Bm.double())
    res = {"B": B, "G": G, "K": 2 * F * P, "N": B * Nw,
           "lag_table_path_err": float((ref_map - exact).abs().max() / exact.max())}
    A32, B32 = A.float(), Bm.float()
    Ab, Bb = A32.bfloat16(), B32.bfloat16()
    Alo, Blo = (A32 - Ab.float()).bfloat16(), (B32 - Bb.float()).bfloat16()
    torch.backends.cuda.matmul.allow_tf32 = True
    variants = {
        "bf16": lambda: Ab @ Bb,
        "bf16x3": lambda: (Ab @ Bb).float() + (Ab @ Blo).float() + (Alo @ Bb).float(),
        "tf32": lambda: A32 @ B32,
    }
    for name, fn in variants.items():
        out = finish(fn().float())
        res[name + "_err"] = float((out - exact).abs().max() / exact.max())
        res[name + "_ms"] = timeit(fn)
    res["flops"] = 2.0 * G * 2 * F * P * B * Nw
    res["table_bytes_bf16"] = G * 2 * F * P * 2
    print(json.dumps(res))


if __name__ == "__main__":
    main()
