"""Probe (NOT product code): VERDICT r1 item 5 -- the STFT stage as a dense contraction on tensor cores.

Identity: frame n of a window (2048 samples, hop 512) is four consecutive 512-sample blocks, so for bin k
    X_n[k] = sum_{b<4} (-i)^{k b} Y_{n+b}[k],      Y_j[k] = sum_{t<512} x[512 j + t] exp(-2 pi i k t / 2048),
i.e. 70 block transforms per window instead of 67 overlapping frames, each a (2 x 198) x 512 real matrix applied to a
512-sample block:  (396 x 512) x (512 x N), N = B * M * Nw * 70 = 109 760 blocks for the C2 sub-batch (44.5 GFLOP).
Inputs are 16-bit PCM: x = hi + lo with hi, lo exact in bf16; the DFT matrix needs its own hi + lo split to keep fp32
accuracy.  This script times the LIBRARY GEMMs (torch.matmul -> cuBLAS) that a hand-written tcgen05 kernel would have
to beat before its epilogue (frame combine, PHAT, pair products) even starts, and measures the error of each variant
against a float64 DFT.  The CUDA-core kernel it competes with: stft_cc_warp_kernel<7>, 249 us per sub-batch, all
stages fused, error ~1e-6."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

dev = torch.device("cuda", 0)
B, M, Nw, NBLK, K, F = 32, 7, 7, 70, 512, 198
N = B * M * Nw * NBLK
g = torch.Generator(device="cuda").manual_seed(0)
pcm = torch.randint(-3000, 3000, (N, K), device=dev, generator=g, dtype=torch.int32)         # speech-level 16-bit samples
x64 = pcm.double() / 32768.0
k = torch.arange(2, 200, device=dev, dtype=torch.float64)[:, None]
t = torch.arange(K, device=dev, dtype=torch.float64)[None, :]
ang = -2 * np.pi * k * t / 2048
W64 = torch.cat([torch.cos(ang), torch.sin(ang)])                      # (396, 512)
ref = x64[:4096] @ W64.t()                                              # float64 reference on a slice


def split_bf16(a64):
    hi = a64.to(torch.bfloat16)
    lo = (a64 - hi.double()).to(torch.bfloat16)
    return hi, lo


xh, xl = split_bf16(x64)
wh, wl = split_bf16(W64)
assert torch.equal((xh.double() + xl.double()), x64), "16-bit PCM is exact in two bf16 terms"


def timeit(fn, n=10):
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


out = torch.empty((N, 2 * F), device=dev, dtype=torch.float32)
variants = {
    "bf16 x bf16 (1 GEMM, matrix rounded to 8 bits)": lambda: torch.matmul(xh, wh.t(), out=None),
    "bf16 split, 3 GEMMs (hi*hi + lo*hi + hi*lo)": lambda: (xh @ wh.t()).float() + (xl @ wh.t()).float() + (xh @ wl.t()).float(),
    "tf32 (1 GEMM)": None,
    "fp32 (1 GEMM)": None,
}
x32, w32 = x64.float(), W64.float()
flop = 2.0 * N * K * 2 * F
for name, fn in variants.items():
    if name.startswith("tf32"):
        torch.backends.cuda.matmul.allow_tf32 = True
        fn = lambda: x32 @ w32.t()
    elif name.startswith("fp32"):
        torch.backends.cuda.matmul.allow_tf32 = False
        fn = lambda: x32 @ w32.t()
    us = timeit(fn)
    got = fn()[:4096].double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    ngemm = 3 if "3 GEMMs" in name else 1
    print(f"{name:50s} {us:8.1f} us  ({ngemm * flop / us / 1e6:7.1f} TFLOP/s)  max err / max |Y| = {err:.2e}")
print(f"N = {N} blocks, output of one GEMM = {N * 2 * F * 4 / 1e6:.0f} MB (fp32) that the epilogue must read back unless fused")
