"""Time asw_patch_powers against the reference's per-row numpy loop (Mic_Array.py:288-296) on a
fine-stage sized batch: 604 rows x 144000 samples."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import native
from acousticswarms_speech_b200.local_utils import max_avg_power

N, T = 604, 144000
x = (0.05 * torch.randn((N, T), device="cuda")).contiguous()
for _ in range(3):
    native.patch_powers(x, demean=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    native.patch_powers(x, demean=False)
e1.record(); torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / 10
h = x[:64].cpu().numpy()
t0 = time.perf_counter()
for j in range(h.shape[0]):
    h[j] = h[j] - np.mean(h[j]); np.sum(h[j] ** 2); max_avg_power(h[j])
host_ms = (time.perf_counter() - t0) * 1e3 * N / h.shape[0]
print(f"patch_powers {N}x{T}: device {dev_ms:.3f} ms ({N*T*4*3/dev_ms/1e6:.0f} GB/s of 3 reads), host loop {host_ms:.0f} ms")
