"""Probe (NOT product code): BASELINE configs[3] -- 256 synthetic 7-mic mixtures, candidate hypercubes sharded over
the ranks, NCCL top-K all-gather + merge.  Launch with torch.distributed.run, one rank per GPU.  Reports the time of
scoring all 256 mixtures (in batches of 32) + exchange, and checks the merged top-K against an unsharded run."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acousticswarms_speech_b200 import dist as adist, native, synth
from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft, window_length
from acousticswarms_speech_b200.srp_phat import SRP_PHAT

rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if "RANK" in os.environ:
    dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size() if dist.is_initialized() else 1
scene = synth.desk_array(7, np.random.default_rng(1), 48000)
node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=48000, n_fft=n_fft, grid_size=0.05,
                threshold=list(SRP_THRESHOLDS), WIDTH=8, device=dev)
G, T, NB, BB, K = node.grids.shape[0], 144000, 8, 32, 128
base = torch.from_numpy(synth.mixtures(scene, 5, T, seeds=list(range(200, 200 + BB)))).to(dev)
batches = [torch.roll(base, shifts=i, dims=0).contiguous() for i in range(NB)]      # 256 mixtures in 8 batches
lag = native.pair_lags(node.grids, scene.mic_positions, 48000, 343.0)
fval, fidx = native.map_topk(node.native.score(batches[3], window_length(T)), K)
for name, make in (("recompute transform on every rank", adist.native_sharded_srp),
                   ("all-gather GCC tables", adist.native_table_exchange_srp)):
    sharded, handle = make(lag, 7, dev)
    for mix in batches[:2]:
        sharded.topk(mix, K)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs = [sharded.topk(mix, K) for mix in batches]
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ok = torch.equal(outs[3][0], fval) and torch.equal(outs[3][1], fidx)
    if rank == 0:
        print(f"C4 [{name}]: {NB * BB} mixtures x G={G} hypercubes sharded over {world} GPU(s): {ms.item():.2f} ms "
              f"= {NB * BB * G / ms.item() * 1e3:.3e} hypercubes/s; merged top-{K} equals the unsharded run: {ok}")
    del sharded, handle
if dist.is_initialized():
    dist.destroy_process_group()
