"""Probe (NOT product code): run-to-run determinism of the mbarrier-pipelined kernels (round-robin STFT, warp-specialised
gather) and of the grouped statistics at bench sizes -- a race in a ring / barrier protocol shows up as a differing hash."""
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import native, synth
from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft, window_length
from acousticswarms_speech_b200.srp_phat import SRP_PHAT

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 30


def sha(t):
    return hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest()[:16]


def score_case(tag, scene, T, B, nspk, **grid):
    node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=48000, n_fft=n_fft, threshold=list(SRP_THRESHOLDS), WIDTH=8,
                    device=dev, **grid)
    nb = min(B, 4)
    base = torch.from_numpy(synth.mixtures(scene, nspk, T, seeds=list(range(300, 300 + nb)))).to(dev)
    mix = torch.cat([torch.roll(base, shifts=11 * i, dims=2) for i in range((B + nb - 1) // nb)], 0)[:B].contiguous()
    win = window_length(T)
    side = torch.cuda.Stream()
    junk = torch.empty(64 << 20, device=dev)
    hs = set()
    for r in range(REPS):
        if r % 3 == 1:                       # perturb the timing: a copy kernel on another stream competes for the SMs
            with torch.cuda.stream(side):
                junk.add_(1.0)
        hs.add(sha(node.native.score(mix, win)))
    print(f"{tag}: {REPS} runs, {len(hs)} distinct map hash(es)", flush=True)
    assert len(hs) == 1
    return mix


mix = score_case("C2 B=64", synth.desk_array(7, np.random.default_rng(1), 48000), 144000, 64, 5, grid_size=0.05)
score_case("M=5 B=7", synth.desk_array(5, np.random.default_rng(2), 48000), 96000, 7, 3, grid_size=0.05)
score_case("C5 B=4", synth.table_array(16, np.random.default_rng(16)), 480000, 4, 8, grid_size=0.025, grid_size_z=0.05)
rng = np.random.default_rng(5)
N = 2200
mi = torch.from_numpy(np.sort(rng.integers(0, 64, size=N)).astype(np.int32)).to(dev)
sh = rng.integers(-200, 201, size=(N, 7)).astype(np.int32); sh[:, 0] = 0
shifts = torch.from_numpy(sh).to(dev)
hs = set()
buf = torch.empty((256, 7, 144000), device=dev)
for r in range(REPS):
    _, mu, sd = native.shift_stack_norm(mix, shifts, mi, out=buf, n_base=0, N=256, grouped=True)
    hs.add(sha(mu) + sha(sd))
print(f"grouped statistics: {REPS} runs, {len(hs)} distinct hash(es)")
assert len(hs) == 1
