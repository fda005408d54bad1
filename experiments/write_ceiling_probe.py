"""Probe (NOT product code): pure-write and copy ceilings of this GPU for the shift-stack roofline.
Times a driver memset (write-only) and a device-to-device copy (read+write) on the shift-stack's own
buffer size (128 x 7 x 144000 float32 = 516 MB), alternating two buffers like the bench's ring."""
import json

import torch


def t(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


a = torch.empty((128, 7, 144000), device="cuda")
b = torch.empty_like(a)
nbytes = a.numel() * 4
k = [0]


def fill():
    (a if k[0] & 1 else b).zero_()
    k[0] += 1


ms_fill = t(fill)
ms_copy = t(lambda: b.copy_(a))
print(json.dumps({"bytes": nbytes, "memset_ms": ms_fill, "memset_GBps": nbytes / ms_fill / 1e6,
                  "copy_ms": ms_copy, "copy_GBps_read_plus_write": 2 * nbytes / ms_copy / 1e6}))
