"""Probe (NOT product code): throughput of the fused shift-stack + normalize_input against the plain shift-stack
(128 patches x 7 mics x 144000 samples, sources of 4 mixtures)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import native

N, B, M, T = 128, 4, 7, 144000
g = torch.Generator(device="cuda").manual_seed(0)
mix = (0.05 * torch.randn((B, M, T), device="cuda", generator=g)).contiguous()
shifts = torch.randint(-250, 251, (N, M), device="cuda", dtype=torch.int32, generator=g)
shifts[:, 0] = 0
mi = (torch.arange(N, device="cuda", dtype=torch.int32) * B // N).contiguous()
out = torch.empty((N, M, T), device="cuda")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


mb = N * M * T * 4 / 1e6
t_plain = timeit(lambda: native.shift_stack(mix, shifts, mi, out=out))
t_norm = timeit(lambda: native.shift_stack_norm(mix, shifts, mi, out=out))
ct = native.CorrTables(M, torch.device("cuda", 0), max_lag=512)
tab = ct.compute(mix)
t_tab = timeit(lambda: ct.compute(mix, out=tab))
t_fast = timeit(lambda: native.shift_stack_norm(mix, shifts, mi, out=out, tables=tab, max_lag=512))
mix1 = mix[:1].contiguous()
t_tab1 = timeit(lambda: ct.compute(mix1))
print(f"correlation tables: {t_tab:.1f} us for {B} mixtures ({t_tab / B:.1f} us each), {t_tab1:.1f} us for one; "
      f"table-driven fused normalize {t_fast:.1f} us ({mb / t_fast * 1e3:.0f} GB/s written)")
print(f"plain shift-stack {t_plain:.1f} us ({mb / t_plain * 1e3:.0f} GB/s written); fused with normalize_input {t_norm:.1f} us "
      f"({mb / t_norm * 1e3:.0f} GB/s written)")
