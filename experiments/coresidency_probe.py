"""Probe (NOT product code): do the scoring kernels and the shift-stack actually share SMs?

One process per configuration (the ASW_* switches are read once).  Times, with CUDA events on the launching streams,
  A  the shift-stack alone (1152 patches, 4.6 GB written),
  B  the transform stage alone (asw_srp_gcc: STFT + PHAT + CC, then the GCC tables), C the gather alone,
  A||B, A||C  both launched back to back on two streams (shift-stack first, then the other; priorities from argv),
and prints the span of each kernel group and of the pair.  A||B == A + B means the CTAs never co-reside."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import native, synth
from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft, window_length
from acousticswarms_speech_b200.srp_phat import SRP_PHAT

stack_pri = int(sys.argv[1]) if len(sys.argv) > 1 else 0
other_pri = int(sys.argv[2]) if len(sys.argv) > 2 else -1
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
scene = synth.desk_array(7, np.random.default_rng(1), 48000)
node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=48000, n_fft=n_fft, grid_size=0.05,
                threshold=list(SRP_THRESHOLDS), WIDTH=8, device=dev)
T, B, N, M = 144000, 64, 1152, 7
base = torch.from_numpy(synth.mixtures(scene, 5, T, seeds=list(range(200, 208)))).to(dev)
mix = torch.cat([torch.roll(base, shifts=i, dims=2) for i in range(B // 8)], 0).contiguous()
mix2 = torch.roll(mix, shifts=3, dims=0).contiguous()
rng = np.random.default_rng(0)
shifts = torch.from_numpy(rng.integers(-300, 300, size=(N, M)).astype(np.int32)).to(dev)
shifts[:, 0] = 0
mix_index = torch.from_numpy(np.sort(rng.integers(0, B, size=N)).astype(np.int32)).to(dev)
out = torch.empty((N, M, T), device=dev)
win = window_length(T)
Nw = node.native.num_windows(T, win)
tabs = node.native.gcc(mix2, win)
smap = node.native.gather(tabs, Nw)
s_stack = torch.cuda.Stream(device=dev, priority=stack_pri)
s_other = torch.cuda.Stream(device=dev, priority=other_pri)


def ev():
    return torch.cuda.Event(enable_timing=True)


def run(do_stack, other):
    a0, a1, b0, b1 = ev(), ev(), ev(), ev()
    torch.cuda.synchronize()
    if do_stack:
        with torch.cuda.stream(s_stack):
            a0.record()
            native.shift_stack(mix, shifts, mix_index, out=out)
            a1.record()
    if other:
        with torch.cuda.stream(s_other):
            b0.record()
            other()
            b1.record()
    torch.cuda.synchronize()
    ta = a0.elapsed_time(a1) if do_stack else 0.0
    tb = b0.elapsed_time(b1) if other else 0.0
    span = max(a0.elapsed_time(a1), a0.elapsed_time(b1)) if (do_stack and other) else max(ta, tb)
    return ta, tb, span


def best(do_stack, other, reps=5):
    r = [run(do_stack, other) for _ in range(reps)]
    return min(r, key=lambda x: x[2])


gcc = lambda: node.native.gcc(mix2, win, out=tabs)
gat = lambda: node.native.gather(tabs, Nw, out=smap)
for _ in range(2):
    run(True, gcc); run(True, gat)
A = best(True, None)[0]
Bt = best(False, gcc)[1]
Ct = best(False, gat)[1]
ab = best(True, gcc)
ac = best(True, gat)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("ASW_"))
print(f"[{tag} pri stack={stack_pri} other={other_pri}] stack {A:.3f} | transform {Bt:.3f} gather {Ct:.3f} | "
      f"stack||transform: stack {ab[0]:.3f} transform {ab[1]:.3f} span {ab[2]:.3f} (sum {A + Bt:.3f}) | "
      f"stack||gather: stack {ac[0]:.3f} gather {ac[1]:.3f} span {ac[2]:.3f} (sum {A + Ct:.3f})")
