"""Probe (NOT product code): where the front half of the C3 fine stage goes for one group of 16 mixtures --
score / select / asw_subdivide / fine table / correlation tables -- with CUDA events per component."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import native, synth
from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft
from acousticswarms_speech_b200.pipeline import FrontEnd
from acousticswarms_speech_b200.srp_phat import SRP_PHAT

dev = torch.device("cuda", 0)
GB = int(sys.argv[1]) if len(sys.argv) > 1 else 16
scene = synth.desk_array(7, np.random.default_rng(1), 48000)
node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=48000, n_fft=n_fft, grid_size=0.05,
                threshold=list(SRP_THRESHOLDS), WIDTH=8, device=dev)
fe = FrontEnd(node, dev)
x = torch.from_numpy(synth.mixtures(scene, 5, 144000, seeds=[100 + b for b in range(GB)])).to(dev)
corr = native.CorrTables(7, dev, max_lag=512)


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


for rep in range(3):
    t = [ev()]
    smap, _, _ = fe.score(x); t.append(ev())
    n_sel, off, wid, pk = fe.select(smap); t.append(ev())
    B, P, D = off.shape
    slot = torch.arange(P, device=dev, dtype=torch.int32)
    valid = slot[None, :] < n_sel.clamp(max=P)[:, None]
    w = (wid * valid).reshape(-1).to(torch.int32).contiguous()
    owner = torch.arange(B, device=dev, dtype=torch.int32)[:, None].expand(B, P).reshape(-1).contiguous()
    t.append(ev())
    cnt, loff, _, root, status = native.subdivide_device(node.native_select, off.reshape(-1, D).contiguous(), w, fe.upper_bound_pairwise())
    t.append(ev())
    tab = native.build_fine_table(cnt, loff, root, w, owner, GB * 1280); t.append(ev())
    tabs = corr.compute(x); t.append(ev())
    torch.cuda.synchronize()
    names = ["score", "select", "torch glue", "subdivide", "fine table", "corr tables"]
    print(f"rep {rep}: candidates {int((w > 0).sum())}, fine rows {int(tab[4])}: " +
          ", ".join(f"{n} {a.elapsed_time(b):.2f} ms" for n, a, b in zip(names, t[:-1], t[1:])))

# distribution of the per-candidate work: member voxels of the coarse patch (root) and leaves
cands = off.reshape(-1, D)[w > 0].contiguous()
cw = w[w > 0].contiguous()
out = native.subdivide(node.native_select, cands, cw, fe.upper_bound_pairwise(), member_cap=1 << 17)
cn = out[0]
rc = np.array([len(m) for m in out[-1]])
print(f"candidates {len(rc)}: root members mean {rc.mean():.0f} median {np.median(rc):.0f} max {rc.max()} "
      f"p90 {np.percentile(rc, 90):.0f}; leaves mean {cn.mean():.1f} max {cn.max()}")
order = np.argsort(-rc)
for name, idx in (("10 largest", order[:10]), ("all but the 10 largest", order[10:]), ("median 100", order[len(order) // 2 - 50: len(order) // 2 + 50])):
    ii = torch.from_numpy(np.ascontiguousarray(idx)).to(dev)
    c2, w2 = cands[ii].contiguous(), cw[ii].contiguous()
    native.subdivide_device(node.native_select, c2, w2, fe.upper_bound_pairwise())
    torch.cuda.synchronize()
    a = ev(); native.subdivide_device(node.native_select, c2, w2, fe.upper_bound_pairwise()); b = ev()
    torch.cuda.synchronize()
    print(f"subdivide of the {name} ({len(idx)} candidates, {rc[idx].sum()} members): {a.elapsed_time(b):.2f} ms")
