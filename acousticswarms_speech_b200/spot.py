"""Shift-and-stack + batched spot-model inference.

Mirror of sep/training/JointModel/network.py: ``roll_by_gather`` (:12-25) and
``DataParallelSpotModel.shift_and_sep`` (:37-104).  The per-patch Python loop that builds an int64
index tensor and calls ``torch.gather`` (:80-83) and the separate ``normalize_input`` pass
(SpeakerLocalization/network.py:28-40) are one fused libasw.so launch per batch.  The network itself
(the existing PyTorch U-Net/transformer) is untouched: any ``nn.Module`` taking ``(B, M, T)`` and a
``(B, 2)`` window embedding works.
"""
import numpy as np
import torch
import torch.nn as nn

from . import native
from .constants import SPOT_BATCH_SIZE


def roll_by_gather(mat, dim, shifts):
    """out[c, t] = mat[c, (t - shifts[c]) mod T] for dim=1 (network.py:12-25), on the device through
    the shift-stack kernel.  ``mat`` (rows, cols) float32 CUDA; ``shifts`` LongTensor (rows, 1)."""
    if dim == 0:
        return roll_by_gather(mat.t().contiguous(), 1, shifts.reshape(-1, 1)).t().contiguous()
    r = (-shifts.reshape(1, -1)).to(device=mat.device, dtype=torch.int32).contiguous()
    return native.shift_stack(mat.contiguous(), r)[0]


def unnormalize_input(data, means, stds):
    """SpeakerLocalization/network.py:42-47."""
    return data * stds + means


class SepRows:
    """Separator outputs of one ``shift_and_sep`` call kept on the device, de-meaned (Mic_Array.py:291):
    ``power[j] = sum(x_j^2)`` (:292) and ``maxavg[j] = max_avg_power(x_j)[0]`` (:293) are host arrays; ``row(j)``
    copies one output to the host when a caller really needs the audio; ``gram(ids)`` returns the float64 inner
    products <x_a, x_b> that SI-SDR (eval_utils.py:11-39) is a function of."""

    def __init__(self, x, power, maxavg):
        self.x, self.power, self.maxavg = x, power, maxavg
        self._rows = {}

    @property
    def shape(self):
        return tuple(self.x.shape)

    def row(self, j):
        j = int(j)
        if j not in self._rows:
            self._rows[j] = self.x[j].cpu().numpy()
        return self._rows[j]

    def gram(self, ids):
        if len(ids) == 0:
            return np.zeros((0, 0))
        sel = self.x[torch.as_tensor(list(ids), device=self.x.device, dtype=torch.int64)].double()
        return (sel @ sel.t()).cpu().numpy()


def si_sdr_from_gram(ee, rr, er):
    """Scale-invariant SDR in dB (eval_utils.py:11-39) from <est,est>, <ref,ref>, <est,ref>:
    a = <ref,est>/<ref,ref>, e_true = a ref, e_res = est - e_true."""
    a = er / rr
    true = a * a * rr
    res = max(ee - 2.0 * a * er + true, 0.0)
    return 10 * np.log10(true / (res + 1e-8))


class DataParallelSpotModel(nn.Module):
    """Drop-in for the reference's class (sep/training/JointModel/network.py:27-104).

    Multi-GPU: the reference wraps the network in ``nn.DataParallel`` (:30), i.e. it stacks every 128-patch batch
    (516 MB) on ONE device and scatters it over PCIe/NVLink each forward.  Here the PATCH LIST is sharded instead: every
    visible device receives the 4 MB mixture once, stacks (and normalises) its own contiguous slice of the patches
    right next to its replica of the network, and only the (n, T) outputs travel back.  Replicas are refreshed from the
    primary module on every call (``torch.nn.parallel.replicate``, what ``nn.DataParallel`` does per forward)."""

    TABLE_MIN_PATCHES = 48      # below this the exact per-patch statistics pass is cheaper than building the tables
    TABLE_MAX_LAG = 512         # samples; patches with a larger pair lag take the exact pass on the device
    SHARD_MIN_PATCHES = 2       # patches per device below which sharding is not worth a replica refresh

    def __init__(self, model, use_fp16=False, batch_size=SPOT_BATCH_SIZE, device=None, data_parallel=True,
                 device_ids=None):
        super().__init__()
        self.model = model
        self.batch_size = batch_size
        self.dtype = torch.bfloat16 if use_fp16 else torch.float32
        if use_fp16:
            self.model.to(torch.bfloat16)
        self._device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.to(self._device)
        if device_ids is None:
            device_ids = list(range(torch.cuda.device_count())) if data_parallel else [self._device.index]
        primary = self._device.index if self._device.index is not None else torch.cuda.current_device()
        self.device_ids = [primary] + [d for d in device_ids if d != primary]
        self._bufs = {}
        self._corr = {}

    @property
    def device(self):
        return self._device

    def _stack_and_separate(self, model, mix, shifts, Strict, save_input, results, saved):
        """network.py:58-101 for the patches of ``shifts`` on the device ``mix`` lives on; fills ``results`` (n, T)."""
        B = self.batch_size
        dev = mix.device
        N = shifts.shape[0]
        M, T = mix.shape
        key = dev.index
        if key not in self._bufs or self._bufs[key].shape != (B, M, T):
            self._bufs[key] = torch.empty((B, M, T), device=dev, dtype=torch.float32)   # reused like `data` (:58)
        buf = self._bufs[key]
        cond = torch.zeros((B, 2), device=dev, dtype=self.dtype)
        cond[:, 0 if Strict == 1 else 1] = 1                                  # :62-73
        # per-mixture correlation tables: every patch's mean / std from a few look-ups instead of a pass over its
        # samples (pays off from a few dozen patches on; fewer take the exact statistics pass)
        tables = None
        if N >= self.TABLE_MIN_PATCHES and T >= native.CorrTables.MIN_T and M >= 2:
            if key not in self._corr or self._corr[key].M != M:
                self._corr[key] = native.CorrTables(M, dev, max_lag=self.TABLE_MAX_LAG)
            tables = self._corr[key].compute(mix)
        for i in range(0, N, B):
            n = min(B, N - i)
            data_norm, means, stds = native.shift_stack_norm(mix, shifts[i:i + n], out=buf, tables=tables,
                                                             max_lag=self.TABLE_MAX_LAG)
            data_norm = data_norm[:n]
            if save_input:
                saved.append(unnormalize_input(data_norm, means, stds).cpu())
            result = model(data_norm.to(self.dtype), cond[:n])
            results[i:i + n] = unnormalize_input(result, means.to(self.dtype), stds.to(self.dtype))[:, 0]

    def _shift_and_sep_device(self, input_channels, patch_list, Strict, save_input):
        """network.py:37-101 up to (not including) the copy back: (N, T) results on the primary device."""
        N = len(patch_list)
        dev = self.device
        self.model.eval()
        with torch.no_grad():
            mix = input_channels.to(dev, dtype=torch.float32).contiguous()       # copy once (:55)
            M, T = mix.shape
            results = torch.zeros((N, T), device=dev, dtype=self.dtype)
            shifts_host = torch.from_numpy(native.offsets_to_shifts(
                np.stack([p.sample_offset for p in patch_list]) if N else np.zeros((0, M - 1))))
            saved = []
            ids = self.device_ids if N >= self.SHARD_MIN_PATCHES * len(self.device_ids) else self.device_ids[:1]
            if len(ids) == 1:
                self._stack_and_separate(self.model, mix, shifts_host.to(dev), Strict, save_input, results, saved)
            else:
                # contiguous slices of the patch list, one per device; everything below is asynchronous per device, the
                # host only issues work, so the devices run concurrently
                replicas = torch.nn.parallel.replicate(self.model, ids, detach=True)
                bounds = [(N * k) // len(ids) for k in range(len(ids) + 1)]
                parts = []
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream(dev))
                for k, d in enumerate(ids):
                    lo, hi = bounds[k], bounds[k + 1]
                    ddev = torch.device("cuda", d)
                    with torch.cuda.device(ddev):
                        torch.cuda.current_stream(ddev).wait_event(ready)
                        mix_d = mix if d == dev.index else mix.to(ddev, non_blocking=True)
                        part = results[lo:hi] if d == dev.index else torch.zeros((hi - lo, T), device=ddev, dtype=self.dtype)
                        self._stack_and_separate(replicas[k], mix_d, shifts_host[lo:hi].to(ddev, non_blocking=True), Strict,
                                                 save_input, part, saved)
                        done = torch.cuda.Event()
                        done.record(torch.cuda.current_stream(ddev))
                        parts.append((lo, hi, part, done, d))
                for lo, hi, part, done, d in parts:
                    if d != dev.index:
                        torch.cuda.current_stream(dev).wait_event(done)
                        results[lo:hi].copy_(part, non_blocking=True)
        return results, (torch.cat(saved) if saved else torch.zeros((0, M, T)))

    def shift_and_sep(self, input_channels, patch_list, Strict=0, save_input=False):
        """network.py:37-104 -> np.ndarray (N, T) float32."""
        results, saved = self._shift_and_sep_device(input_channels, patch_list, Strict, save_input)
        out = results.cpu().float().numpy()
        if save_input:
            return out, saved
        return out

    def shift_and_sep_device(self, input_channels, patch_list, Strict=0, window=12000):
        """``shift_and_sep`` whose outputs stay on the device: -> ``SepRows`` (per-row power statistics on the host,
        audio rows and their Gram matrix fetched on demand)."""
        results, _ = self._shift_and_sep_device(input_channels, patch_list, Strict, False)
        x = results.float().contiguous()
        _, power, maxavg, _ = native.patch_powers(x, window=window, demean=True)
        return SepRows(x, power.cpu().numpy(), maxavg.cpu().numpy())

    def shift_and_sep_powers(self, input_channels, patch_list, Strict=0, window=12000):
        """``shift_and_sep`` followed, still on the device, by what every caller does next with each row
        (local_utils_3d.py:342-349, Mic_Array.py:288-296): de-mean in place, ``np.sum(x ** 2)`` and
        ``max_avg_power(x)[0]``.  -> (de-meaned outputs (N, T) float32 numpy, powers (N,), max_avg_powers (N,))."""
        results, _ = self._shift_and_sep_device(input_channels, patch_list, Strict, False)
        x = results.float().contiguous()
        _, power, maxavg, _ = native.patch_powers(x, window=window, demean=True)
        return x.cpu().numpy(), power.cpu().numpy(), maxavg.cpu().numpy()
