"""GPU parity: shift-and-stack vs the oracle restatement of roll_by_gather
(sep/training/JointModel/network.py:12-25, 75-83).  Integer index work: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import shift_oracle

pytestmark = pytest.mark.gpu


def _run(mix, offsets, mix_index=None):
    from acousticswarms_speech_b200 import native
    dev = torch.device("cuda", 0)
    shifts = native.offsets_to_shifts(offsets)
    out = native.shift_stack(torch.from_numpy(mix).to(dev), torch.from_numpy(shifts).to(dev),
                             None if mix_index is None else torch.from_numpy(mix_index.astype(np.int32)).to(dev))
    torch.cuda.synchronize()
    return out.cpu().numpy(), shifts


@pytest.mark.parametrize("T", [4096, 4100, 4099, 144000, 132300, 7])
@pytest.mark.parametrize("M", [2, 7])
def test_shift_stack_bit_exact(cuda_device, T, M):
    rng = np.random.default_rng(T * 31 + M)
    mix = rng.standard_normal((M, T)).astype(np.float32)
    offs = rng.integers(-300, 300, size=(9, M - 1)).astype(np.float64)
    offs[0] = 0
    offs[1] = T            # |r| >= T wraps to zero shift
    offs[2] = -T - 3
    offs[3] = 2 * T + 5
    offs[4, 0] = 2.5       # round-half-even -> 2
    offs[4, -1] = -3.5     # -> -4
    got, shifts = _run(mix, offs)
    want = shift_oracle.shift_stack(mix, list(offs))
    assert np.array_equal(shifts[:, 1:], np.stack([shift_oracle.shift_indices(o)[1:] for o in offs]))
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_shift_stack_batched_mixtures(cuda_device):
    rng = np.random.default_rng(5)
    B, M, T = 3, 5, 24000
    mix = rng.standard_normal((B, M, T)).astype(np.float32)
    offs = rng.integers(-250, 250, size=(11, M - 1)).astype(np.float64)
    mi = rng.integers(0, B, size=11)
    got, _ = _run(mix, offs, mi)
    for n in range(11):
        want = shift_oracle.shift_stack(mix[mi[n]], [offs[n]])[0]
        assert np.array_equal(got[n], want)


def test_shift_stack_empty(cuda_device):
    from acousticswarms_speech_b200 import native
    dev = torch.device("cuda", 0)
    out = native.shift_stack(torch.zeros((1, 3, 64), device=dev), torch.zeros((0, 3), dtype=torch.int32, device=dev))
    assert out.shape == (0, 3, 64)


def test_shift_stack_full_size_properties(cuda_device):
    """BASELINE-size check through size-independent properties: shifting by r then by -r is the
    identity, and per-row sums are shift invariant."""
    from acousticswarms_speech_b200 import native
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(3)
    M, T, N = 7, 144000, 128
    mix = torch.randn((M, T), generator=g).to(dev)
    r = torch.randint(-260, 260, (N, M), generator=g, dtype=torch.int32)
    r[:, 0] = 0
    r = r.to(dev)
    out = native.shift_stack(mix, r)
    # undo patch n's shift with the same kernel: `out` is read as a batch of N "mixtures"
    for n in (0, 1, 57, N - 1):
        b = native.shift_stack(out, -r[n:n + 1], torch.tensor([n], dtype=torch.int32, device=dev))
        assert torch.equal(b[0], mix)
    # every output row is a permutation of its source row
    assert torch.equal(out.sort(-1).values, mix.sort(-1).values.expand(N, M, T))


def test_shift_stack_norm(cuda_device):
    from acousticswarms_speech_b200 import native
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(11)
    M, T, N = 7, 36000, 6
    mix = (0.1 * rng.standard_normal((M, T))).astype(np.float32)
    offs = rng.integers(-200, 200, size=(N, M - 1)).astype(np.float64)
    shifts = native.offsets_to_shifts(offs)
    dn, mu, sd = native.shift_stack_norm(torch.from_numpy(mix).to(dev), torch.from_numpy(shifts).to(dev))
    torch.cuda.synchronize()
    want, wmu, wsd = shift_oracle.normalize_input(shift_oracle.shift_stack(mix, list(offs)))
    # fp32 path: 1e-4 relative (north_star), measured normwise against the tensor's max
    assert np.abs(mu.cpu().numpy() - wmu).max() <= 1e-4 * np.abs(wsd).max()
    assert np.abs(sd.cpu().numpy() - wsd).max() <= 1e-4 * np.abs(wsd).max()
    assert np.abs(dn.cpu().numpy() - want).max() <= 1e-4 * np.abs(want).max()


def test_shift_stack_norm_is_reproducible(cuda_device):
    """The statistics pass reduces in a fixed order (no atomics across CTAs): repeated calls, alone or as part of a
    larger batch, return the same bits."""
    from acousticswarms_speech_b200 import native
    rng = np.random.default_rng(12)
    B, M, T, N = 3, 7, 144000, 40
    mix = torch.from_numpy((0.1 * rng.standard_normal((B, M, T))).astype(np.float32)).cuda()
    shifts = torch.from_numpy(rng.integers(-350, 351, size=(N, M)).astype(np.int32)).cuda()
    shifts[:, 0] = 0
    mi = torch.from_numpy(rng.integers(0, B, size=N).astype(np.int32)).cuda()
    a = [t.clone() for t in native.shift_stack_norm(mix, shifts, mi)]
    for _ in range(3):
        b = native.shift_stack_norm(mix, shifts, mi)
        assert all(torch.equal(x, y) for x, y in zip(a, b))
    c = native.shift_stack_norm(mix, shifts[7:19].contiguous(), mi[7:19].contiguous())
    assert torch.equal(c[0], a[0][7:19]) and torch.equal(c[1], a[1][7:19]) and torch.equal(c[2], a[2][7:19])


def test_pcm16_ingest_is_exact(cuda_device):
    from acousticswarms_speech_b200 import native
    rng = np.random.default_rng(2)
    for n in (0, 1, 7, 8, 4099, 7 * 144000):
        pcm = rng.integers(-32768, 32768, size=n, dtype=np.int16)
        if n >= 4:
            pcm[:4] = [-32768, 32767, 0, -1]
        got = native.pcm16_to_f32(torch.from_numpy(pcm).cuda()).cpu().numpy()
        assert np.array_equal(got, pcm.astype(np.float32) / np.float32(32768.0))
