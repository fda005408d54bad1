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


# ---- correlation tables (asw_corr_tables) and the table-driven fused normalize (asw_shift_stack_norm_tab) ----------
def _circ_corr(a, b, lag):
    """sum_t a[t] b[(t + lag) mod T] in float64."""
    return float(np.dot(a, np.roll(b, -lag)))


@pytest.mark.parametrize("M,T,L", [(3, 8192, 512), (7, 144000, 512), (4, 132300, 300), (2, 4099, 100), (16, 20000, 128), (32, 8200, 40)])
def test_corr_tables_match_direct_correlation(cuda_device, M, T, L):
    """S_c, E_c exact; R_cc'(l) within fp32-FFT round-off of the float64 circular correlation of the re-quantised
    channels, for every pair at the table's edges and at random lags (T odd exercises the scalar loader, the last
    block is ragged for every T here)."""
    from acousticswarms_speech_b200 import native
    rng = np.random.default_rng(M * 1000 + L)
    B = 2
    mix = (0.2 * rng.standard_normal((B, M, T))).astype(np.float32)
    mix[1, 0] += 0.05                                   # a DC offset on one channel
    ct = native.CorrTables(M, cuda_device, max_lag=L)
    tab = ct.compute(torch.from_numpy(mix).to(cuda_device)).cpu().numpy()
    assert tab.shape == (B, 2 * M + (M * (M - 1) // 2) * (2 * L + 1))
    q = (np.rint(mix * np.float32(32768.0)) / np.float32(32768.0)).astype(np.float64)
    worst = 0.0
    for b in range(B):
        assert np.allclose(tab[b, :M], q[b].sum(1), rtol=0, atol=1e-9)
        assert np.allclose(tab[b, M:2 * M], (q[b] ** 2).sum(1), rtol=1e-13, atol=0)
        p = 0
        for i in range(M):
            for j in range(i + 1, M):
                row = tab[b, 2 * M + p * (2 * L + 1): 2 * M + (p + 1) * (2 * L + 1)]
                scale = np.sqrt((q[b, i] ** 2).sum() * (q[b, j] ** 2).sum())
                for lag in [-L, -L + 1, -1, 0, 1, L - 1, L] + list(rng.integers(-L, L + 1, size=6 if M <= 8 else 1)):
                    worst = max(worst, abs(row[lag + L] - _circ_corr(q[b, i], q[b, j], int(lag))) / scale)
                p += 1
    print(f"corr tables M={M} T={T} L={L}: worst |dR| / sqrt(E_i E_j) = {worst:.2e}")
    assert worst <= 2e-6


def test_shift_stack_norm_tab_matches_exact_pass(cuda_device):
    """Table-driven statistics vs the exact pass and vs the oracle: means / stds within 1e-6 relative of the exact
    pass, normalised samples within 1e-5, oracle bar 1e-4; a patch whose lag exceeds the table takes the exact pass
    on the device and is then bit-identical to it; results are reproducible bit for bit."""
    from acousticswarms_speech_b200 import native
    rng = np.random.default_rng(21)
    B, M, T, N, L = 3, 7, 144000, 48, 512
    t = np.arange(T)
    mix = np.zeros((B, M, T), dtype=np.float32)
    for b in range(B):                                   # correlated channels: a common source at different delays + noise
        src = np.convolve(rng.standard_normal(T + 600), np.hanning(9), mode="same")
        for c in range(M):
            d = int(rng.integers(0, 300))
            mix[b, c] = 0.05 * src[d:d + T] + 0.01 * rng.standard_normal(T)
    shifts = rng.integers(-250, 251, size=(N, M)).astype(np.int32)
    shifts[:, 0] = 0
    shifts[5, 3] = 700                                   # |r_3 - r_c| > L: exact pass for this patch
    shifts[6] = shifts[6] + T                            # shifts beyond T reduce mod T
    shifts[6, 0] = 0
    mi = rng.integers(0, B, size=N).astype(np.int32)
    x = torch.from_numpy(mix).to(cuda_device)
    sh, mid = torch.from_numpy(shifts).to(cuda_device), torch.from_numpy(mi).to(cuda_device)
    ct = native.CorrTables(M, cuda_device, max_lag=L)
    tab = ct.compute(x)
    exact = [v.clone() for v in native.shift_stack_norm(x, sh, mid)]
    fast = [v.clone() for v in native.shift_stack_norm(x, sh, mid, tables=tab, max_lag=L)]
    again = native.shift_stack_norm(x, sh, mid, tables=ct.compute(x), max_lag=L)
    assert all(torch.equal(a, b) for a, b in zip(fast, again)), "table path is not reproducible"
    mu_e, sd_e = exact[1].cpu().numpy().ravel(), exact[2].cpu().numpy().ravel()
    mu_f, sd_f = fast[1].cpu().numpy().ravel(), fast[2].cpu().numpy().ravel()
    rel_sd = np.abs(sd_f - sd_e) / sd_e
    print(f"norm_tab: max rel std diff {rel_sd.max():.2e}, max |mean diff| / std {np.abs(mu_f - mu_e).max() / sd_e.min():.2e}")
    assert rel_sd.max() <= 1e-6
    assert np.abs(mu_f - mu_e).max() <= 1e-6 * sd_e.min()
    assert torch.equal(fast[0][5], exact[0][5]) and sd_f[5] == sd_e[5], "out-of-table patch must take the exact pass"
    dn_e, dn_f = exact[0].cpu().numpy(), fast[0].cpu().numpy()
    assert np.abs(dn_f - dn_e).max() <= 1e-5 * np.abs(dn_e).max()
    for n in (0, 5, 6, N - 1):
        want, wmu, wsd = shift_oracle.normalize_input(shift_oracle.roll_by_gather(mix[mi[n]], -shifts[n].astype(np.int64))[None])
        assert abs(sd_f[n] - wsd.ravel()[0]) <= 1e-4 * wsd.ravel()[0]
        assert np.abs(dn_f[n] - want[0]).max() <= 1e-4 * np.abs(want).max()


def test_shift_stack_norm_tab_ill_conditioned_patch_takes_exact_pass(cuda_device):
    """Channels that cancel in the mic-average (variance far below the table's round-off bound) must not be
    normalised from the tables."""
    from acousticswarms_speech_b200 import native
    rng = np.random.default_rng(22)
    M, T, L = 2, 16384, 64
    a = (0.1 * rng.standard_normal(T)).astype(np.float32)
    mix = np.stack([a, -a + (1e-4 * rng.standard_normal(T)).astype(np.float32)])[None]
    x = torch.from_numpy(mix).to(cuda_device)
    sh = torch.zeros((1, M), dtype=torch.int32, device=cuda_device)
    ct = native.CorrTables(M, cuda_device, max_lag=L)
    exact = native.shift_stack_norm(x, sh)
    fast = native.shift_stack_norm(x, sh, tables=ct.compute(x), max_lag=L)
    assert all(torch.equal(p, q) for p, q in zip(exact, fast))


@pytest.mark.parametrize("B,M,T", [(3, 7, 144000), (2, 4, 4099), (5, 8, 20000), (2, 2, 700), (1, 7, 132300)])
def test_shift_stack_norm_grouped(cuda_device, B, M, T):
    """asw_shift_stack_norm_grouped (rows grouped by mixture, one tiled pass per mixture, exact integer sums): against
    the oracle's normalize_input of the shifted stack (1e-4, north_star) and against the per-patch pass of
    asw_shift_stack_norm (1e-6: that one rounds the mic average to float32 before squaring).  Covers shifts beyond the
    512-sample halo of a tile (global read path), a mixture without patches, a ragged last tile, rows beyond the
    device-resident count, a row window (n_base), and reproducibility (integer atomics: any order, same bits)."""
    from acousticswarms_speech_b200 import native
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(100 + B * M)
    mix_h = (0.1 * rng.standard_normal((B, M, T))).astype(np.float32)
    mix = torch.from_numpy(mix_h).to(dev)
    per = [9, 0, 14, 3, 40][:B]                                 # mixture 1 has no patch
    mi_h = np.concatenate([np.full(n, b, dtype=np.int32) for b, n in enumerate(per)])
    N = mi_h.size
    lim = min(300, T // 3)
    sh_h = rng.integers(-lim, lim + 1, size=(N, M)).astype(np.int32)
    sh_h[:, 0] = rng.integers(-T, T, size=N)                    # a common shift of any size is legal
    sh_h[:, 1:] += sh_h[:, :1]
    if T > 4000:
        sh_h[2, 1] += 900                                       # relative shift beyond the halo
        sh_h[5, M - 1] -= 1500
    shifts, mi = torch.from_numpy(sh_h).to(dev), torch.from_numpy(mi_h).to(dev)
    out_g, mu_g, sd_g = [t.clone() for t in native.shift_stack_norm(mix, shifts, mi, grouped=True)]
    out_e, mu_e, sd_e = native.shift_stack_norm(mix, shifts, mi)
    torch.cuda.synchronize()
    assert torch.isfinite(out_g).all()
    assert (mu_g - mu_e).abs().max().item() <= 1e-6 * sd_e.abs().max().item()
    assert ((sd_g - sd_e).abs() <= 1e-6 * sd_e.abs()).all()
    assert (out_g - out_e).abs().max().item() <= 1e-5 * out_e.abs().max().item()
    for n in (0, N // 2, N - 1):
        q = np.round(mix_h[mi_h[n]].astype(np.float64) * 32768.0) / 32768.0
        ref = np.mean([np.roll(q[c], -int(sh_h[n, c])) for c in range(M)], axis=0)
        assert abs(mu_g[n].item() - ref.mean()) <= 1e-6 * ref.std(ddof=1)
        assert abs(sd_g[n].item() - ref.std(ddof=1)) <= 1e-6 * ref.std(ddof=1)
    for _ in range(2):                                          # same bits every time
        again = native.shift_stack_norm(mix, shifts, mi, grouped=True)
        assert all(torch.equal(a, b) for a, b in zip((out_g, mu_g, sd_g), again))
    if N < 16:
        return
    # a window of rows of a counted table: rows [4, 4 + 12) of which the device count admits the first 10
    n_total = torch.tensor([14], dtype=torch.int32, device=dev)
    buf = torch.full((12, M, T), 7.0, device=dev)
    o, m_, s_ = native.shift_stack_norm(mix, shifts, mi, out=buf, n_total=n_total, n_base=4, N=12, grouped=True)
    assert torch.equal(o[:10], out_g[4:14]) and torch.equal(m_[:10], mu_g[4:14]) and torch.equal(s_[:10], sd_g[4:14])
    assert (buf[10:] == 7.0).all()


def test_shift_stack_skips_rows_with_foreign_mixture_index(cuda_device):
    from acousticswarms_speech_b200 import native
    B, M, T = 2, 3, 4096
    x = torch.randn((B, M, T), device=cuda_device)
    sh = torch.zeros((3, M), dtype=torch.int32, device=cuda_device)
    mi = torch.tensor([0, 7, -1], dtype=torch.int32, device=cuda_device)
    out = torch.full((3, M, T), 123.0, device=cuda_device)
    native.shift_stack(x, sh, mi, out=out)
    assert torch.equal(out[0], x[0]) and bool((out[1:] == 123.0).all())


def test_corr_tables_argument_errors_and_empty_batches(cuda_device):
    """Error behaviour of the table entry points: short signals are refused (callers take the exact pass), mismatched
    tables are refused, empty patch lists return empty results."""
    from acousticswarms_speech_b200 import _lib, native
    ct = native.CorrTables(3, cuda_device, max_lag=64)
    with pytest.raises(_lib.AswError):
        ct.compute(torch.zeros((1, 3, 4095), device=cuda_device))          # T < 4096
    with pytest.raises(_lib.AswError):
        ct.compute(torch.zeros((1, 4, 8192), device=cuda_device))          # wrong channel count
    with pytest.raises(_lib.AswError):
        native.CorrTables(3, cuda_device, max_lag=513)
    x = torch.randn((2, 3, 8192), device=cuda_device)
    tab = ct.compute(x)
    sh = torch.zeros((0, 3), dtype=torch.int32, device=cuda_device)
    out, mu, sd = native.shift_stack_norm(x, sh, torch.zeros((0,), dtype=torch.int32, device=cuda_device), tables=tab, max_lag=64)
    assert out.shape == (0, 3, 8192) and mu.shape == (0, 1, 1)
    sh = torch.zeros((2, 3), dtype=torch.int32, device=cuda_device)
    with pytest.raises(_lib.AswError):                                       # table of another max_lag
        native.shift_stack_norm(x, sh, tables=tab, max_lag=32)
    with pytest.raises(_lib.AswError):                                       # tables of another batch
        native.shift_stack_norm(x, sh, tables=tab[:1].contiguous(), max_lag=64)
    # all-zero channels: variance 0 -> the exact pass decides (std 0, like torch's), no NaN from the table path
    z = torch.zeros((1, 3, 8192), device=cuda_device)
    a = native.shift_stack_norm(z, sh[:1].contiguous(), tables=ct.compute(z), max_lag=64)
    b = native.shift_stack_norm(z, sh[:1].contiguous())
    assert torch.equal(a[2], b[2]) and float(a[2].max()) == 0.0


def test_fine_table_with_empty_candidate_slots(cuda_device):
    """asw_build_fine_table on a padded candidate list: slots with width <= 0 contribute nothing, the others their
    leaves followed by their own (checked-out) offsets; the capacity clamps the total."""
    from acousticswarms_speech_b200 import native
    n, L, D = 5, 4, 3
    cnt = torch.tensor([2, 0, 9, 1, 3], dtype=torch.int32, device=cuda_device)       # 9 > L: clamped to L leaves
    off = torch.arange(n * L * D, dtype=torch.int32, device=cuda_device).reshape(n, L, D)
    root = -torch.arange(n * 2 * D, dtype=torch.int32, device=cuda_device).reshape(n, 2, D)
    wid = torch.tensor([8, 0, 8, -1, 4], dtype=torch.int32, device=cuda_device)
    own = torch.tensor([0, 0, 1, 1, 2], dtype=torch.int32, device=cuda_device)
    sh, mi, ci, cs, nt = native.build_fine_table(cnt, off, root, wid, own, 64)
    want_rows, want_mi, want_ci = [], [], []
    for i in (0, 2, 4):
        k = min(int(cnt[i]), L)
        want_rows += [[0] + off[i, q].tolist() for q in range(k)] + [[0] + root[i, 0].tolist()]
        want_mi += [int(own[i])] * (k + 1)
        want_ci += [i] * (k + 1)
    total = int(nt[0])
    assert total == len(want_rows) == 3 + 5 + 4
    assert sh[:total].tolist() == want_rows and mi[:total].tolist() == want_mi and ci[:total].tolist() == want_ci
    assert cs.tolist() == [0, 3, 3, 8, 8, 12]
    sh2, mi2, ci2, cs2, nt2 = native.build_fine_table(cnt, off, root, wid, own, 7)
    assert int(nt2[0]) == 7 and sh2[:7].tolist() == want_rows[:7]
    nothing = native.build_fine_table(cnt, off, root, torch.zeros_like(wid), own, 8)
    assert int(nothing[4][0]) == 0


def test_persistent_variant_in_a_fresh_process(cuda_device):
    """ASW_STACK=persist selects shift_stack_persist_kernel (one resident CTA per SM, shuffle realignment); the switch
    is read once per process, so the bit-exact and normalize tests of this file are re-run in a child process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, ASW_STACK="persist")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_shift.py"), "-x", "-q", "-m", "gpu",
                        "-k", "bit_exact or batched or full_size or shift_stack_norm"], env=env, cwd=root,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
