"""GPU, world size 2 over NCCL (skipped with fewer than two GPUs): hypercube-sharded and mixture-sharded
scoring reproduce the single-GPU results bit for bit (each map value is computed wholly on one rank and
the frame-group size does not depend on the batch)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_scoring_two_gpus(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_RESULT ok=1 world=2" in r.stdout


def test_two_devices_in_one_process(cuda_device):
    """One host process driving handles on two devices while the *current* device stays cuda:0: handle-based
    entry points make their device current themselves, the per-kernel shared-memory attributes are set per device,
    and the handle-less launches follow their tensors' device.  Results equal cuda:0's bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import numpy as np
    from acousticswarms_speech_b200 import native, synth
    from oracle import geometry_oracle
    scene = synth.small_scene(n_mics=4, seed=2)
    geo = geometry_oracle.GeometryOracle(scene.mic_positions, scene.roi, build_fine=False)
    lag = native.pair_lags(geo.grids, scene.mic_positions, scene.fs, 343.0)
    mix = torch.from_numpy(synth.mixtures(scene, 2, 72000, seeds=[1, 2]))
    shifts = torch.tensor([[0, 5, -7, 11], [0, -300, 17, 71999]], dtype=torch.int32)
    mi = torch.tensor([1, 0], dtype=torch.int32)
    outs = []
    torch.cuda.set_device(0)
    for d in (0, 1):
        dev = torch.device("cuda", d)
        srp = native.NativeSRP(lag, 4, device=dev)
        m = srp.score(mix.to(dev), 36000)
        val, idx = native.map_topk(m, 8)
        st = native.shift_stack(mix.to(dev), shifts.to(dev), mi.to(dev))
        _, power, maxavg, _ = native.patch_powers(st.reshape(-1, st.shape[-1]).clone(), demean=True)
        outs.append([t.cpu().numpy() for t in (m, val, idx, st, power, maxavg)])
        assert torch.cuda.current_device() == 0
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_patch_sharded_spot_model_two_gpus(cuda_device):
    """DataParallelSpotModel over two devices: the patch list is sharded (each device stacks and normalises its slice
    next to its replica of the network) instead of nn.DataParallel scattering the stacked batch
    (sep/training/JointModel/network.py:30).  Outputs must equal the single-device path: bit for bit for a network
    that is pure arithmetic on the normalised stack, to float32 round-off for a convolutional one (replicated
    weights)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import numpy as np
    from acousticswarms_speech_b200.spot import DataParallelSpotModel
    from oracle import prune_oracle

    class MeanNet(torch.nn.Module):
        def forward(self, x, cond):
            return x.mean(1, keepdim=True) * (cond[:, 1:2] + 2 * cond[:, 0:1]).unsqueeze(-1)

    class ConvNet(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.c = torch.nn.Conv1d(m, 1, 9, padding=4)

        def forward(self, x, cond):
            return self.c(x) * (cond[:, 1:2] + 2 * cond[:, 0:1]).unsqueeze(-1)

    rng = np.random.default_rng(3)
    M, T, N = 7, 48000, 301
    mix = torch.from_numpy((0.05 * rng.standard_normal((M, T))).astype(np.float32))
    patches = [prune_oracle.Patch(rng.integers(-200, 201, size=M - 1).astype(np.int64), np.full(M - 1, 4), None) for _ in range(N)]
    torch.manual_seed(0)
    for net, exact in ((MeanNet(), True), (ConvNet(M), False)):
        single = DataParallelSpotModel(net, batch_size=64, data_parallel=False)
        multi = DataParallelSpotModel(net, batch_size=64, device_ids=[0, 1])
        assert multi.device_ids == [0, 1] and single.device_ids == [0]
        for strict in (0, 1):
            a = single.shift_and_sep(mix, patches, Strict=strict)
            b = multi.shift_and_sep(mix, patches, Strict=strict)
            assert a.shape == b.shape == (N, T)
            if exact:
                assert np.array_equal(a, b)
            else:
                assert np.abs(a - b).max() <= 1e-5 * np.abs(a).max()
        rows = multi.shift_and_sep_device(mix, patches[:130], Strict=1)
        assert rows.shape == (130, T) and rows.x.device.index == 0
    assert torch.cuda.current_device() == 0
