"""Probe (NOT product code): time asw_shift_stack (128 patches x 7 x 144000) for the library currently in place."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import native
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
mix = torch.randn((4, 7, 144000), generator=g).to(dev)
sh = torch.randint(-250, 250, (128, 7), generator=g, dtype=torch.int32); sh[:, 0] = 0
sh = sh.to(dev); mi = torch.randint(0, 4, (128,), generator=g, dtype=torch.int32).to(dev)
bufs = [torch.empty((128, 7, 144000), device=dev) for _ in range(2)]
for i in range(6): native.shift_stack(mix, sh, mi, out=bufs[i & 1])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(40): native.shift_stack(mix, sh, mi, out=bufs[i & 1])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 40
print(f"{sys.argv[1] if len(sys.argv) > 1 else ''} {ms*1e3:.1f} us  {128*7*144000*4/ms/1e6:.0f} GB/s")
