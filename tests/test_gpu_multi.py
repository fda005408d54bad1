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
