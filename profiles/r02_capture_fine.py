"""ncu target for the fine-stage kernels (round 2): correlation tables, table-driven fused shift-stack + normalize,
asw_subdivide and the fine-table kernels, on one group of 16 C2 mixtures.
    ncu --set full --clock-control none --import-source on -k regex:"xcorr|shift_ref_stats|shift_stack_vec|subdivide|fine_table" \
        -s 14 -c 14 -o gpurun_out/prof_r02_fine python profiles/r02_capture_fine.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import native, synth
from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft
from acousticswarms_speech_b200.pipeline import FrontEnd
from acousticswarms_speech_b200.srp_phat import SRP_PHAT

dev = torch.device("cuda", 0)
scene = synth.desk_array(7, np.random.default_rng(1), 48000)
node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=48000, n_fft=n_fft, grid_size=0.05,
                threshold=list(SRP_THRESHOLDS), WIDTH=8, device=dev)
fe = FrontEnd(node, dev, launch_batches=8, ring=16)
x = torch.from_numpy(synth.mixtures(scene, 5, 144000, seeds=[100 + b for b in range(16)])).to(dev)
L = native.CorrTables.lag_for_geometry(node.mic_pos, node.FS)
corr = native.CorrTables(7, dev, max_lag=L)
for rep in range(2):                       # the second repetition is the one profiled (-s 14)
    n_sel, off, wid, pk = fe.select(fe.score(x)[0])
    shifts, mi, ci, cstart, ntot, status, cnt = fe.fine_table(n_sel, off, wid, 16 * 1280)
    tabs = corr.compute(x)
    n = int(ntot[0])
    fe.stack_norm_counted(x, shifts, mi, ntot, min(n, 2048), tables=tabs, max_lag=L)
torch.cuda.synchronize()
print("fine rows", n, "max_lag", L)
