"""Worker for tests/test_gpu_multi.py (launched by torch.distributed.run, one rank per GPU):
hypercube-sharded scoring over NCCL must reproduce the single-GPU map and top-K bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from acousticswarms_speech_b200 import dist as adist, native, synth  # noqa: E402
from acousticswarms_speech_b200.constants import SRP_THRESHOLDS, freq_bins, n_fft, window_length  # noqa: E402
from acousticswarms_speech_b200.srp_phat import SRP_PHAT  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world = dist.get_world_size()
    scene = synth.small_scene(n_mics=5, seed=11)
    scene.roi = [0.5, 2.3, -0.9, 0.9, 0.0, 0.6]
    node = SRP_PHAT(scene.mic_positions, freq_bins, scene.roi, FS=scene.fs, n_fft=n_fft, grid_size=0.05,
                    threshold=list(SRP_THRESHOLDS), WIDTH=8, device=dev)
    G = node.grids.shape[0]
    mix = torch.from_numpy(synth.mixtures(scene, 3, 72000, seeds=[1, 2, 3, 4])).to(dev)
    # single-GPU reference on every rank (full lag table)
    full = node.native.score(mix, window_length(72000)).clone()
    K = 32
    fval, fidx = native.map_topk(full, K)
    # hypercube-sharded: this rank holds lag rows [g0, g1)
    lag = native.pair_lags(node.grids, scene.mic_positions, scene.fs, 343.0)
    sharded, handle = adist.native_sharded_srp(lag, scene.mic_positions.shape[0], dev)
    gmap = sharded.full_map(mix)
    val, idx = sharded.topk(mix, K)
    ok = torch.equal(gmap, full) and torch.equal(val, fval) and torch.equal(idx, fidx)
    # table exchange: transform stage sharded over mixtures, gather over hypercubes, GCC tables all-gathered (NCCL)
    ex, ex_handle = adist.native_table_exchange_srp(lag, scene.mic_positions.shape[0], dev)
    emap = ex.full_map(mix)
    eval_, eidx = ex.topk(mix, K)
    ok = ok and torch.equal(emap, full) and torch.equal(eval_, fval) and torch.equal(eidx, fidx)
    odd = mix[:3].contiguous()                                   # 3 mixtures over 2 ranks: ragged split
    ok = ok and torch.equal(ex.full_map(odd), full[:3])
    # the two halves on one handle are asw_srp_score
    tabs = node.native.gcc(mix, window_length(72000))
    ok = ok and torch.equal(node.native.gather(tabs, node.native.num_windows(72000, window_length(72000))), full)
    # mixture-sharded: each rank scores its slice of the batch, no communication
    b0, b1 = adist.shard_range(mix.shape[0], rank, world)
    part = node.native.score(mix[b0:b1].contiguous(), window_length(72000)) if b1 > b0 else full[:0]
    ok = ok and torch.equal(part, full[b0:b1])
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"DIST_RESULT ok={int(flag.item())} world={world} G={G} shard=[{sharded.g0},{sharded.g1})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
