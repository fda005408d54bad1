"""Oracle: the reference's CPU path as it actually runs (torch ops, all host threads).

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is what ``bench.py`` times as the
``cpu_baseline`` leg and as ``--impl reference``: unlike ``srp_oracle`` (which recomputes the steering
table chunk by chunk to stay small), it precomputes the (G, F, P) float64 table once, exactly as
SRP_PHAT.__init__ does (sep/Traditional_SP/SRP_Prunning.py:221-243, excluded from inference time by
the reference's README.md:144), and then evaluates SRP_Map_WINDOW_torch (:387-433) and the shift loop
(sep/training/JointModel/network.py:75-83) with the same torch calls, on CPU tensors.
"""
import os

import numpy as np
import torch

from . import prune_oracle, srp_oracle
from .pra_stft import analysis


class CpuReferencePath:
    def __init__(self, geom, freq_bins, fs, nfft, C=343.0, threads=None):
        """``geom``: oracle.geometry_oracle.GeometryOracle."""
        self.threads = threads or os.cpu_count()
        torch.set_num_threads(self.threads)
        self.geom = geom
        self.freq_bins = np.asarray(freq_bins)
        self.fs, self.nfft, self.C = fs, nfft, C
        M = geom.mic_pos.shape[0]
        G = geom.grids.shape[0]
        F, P = len(self.freq_bins), M * (M - 1) // 2
        real = torch.empty((G, F, P), dtype=torch.float64)
        imag = torch.empty((G, F, P), dtype=torch.float64)
        for s in range(0, G, 1024):
            tab = srp_oracle.steering_table_chunk(geom.grids[s:s + 1024], geom.mic_pos, self.freq_bins, fs, nfft, C)
            real[s:s + 1024] = torch.from_numpy(tab.real.copy())
            imag[s:s + 1024] = torch.from_numpy(tab.imag.copy())
        self.mode_mat_flat_real, self.mode_mat_flat_imag = real, imag
        av = np.arange(M)[:, None]
        self.mask_triu = (av < av.T).flatten()
        self.cluster_offsets = [c[0] for c in geom.clusters]

    def srp_map(self, signal, window, tol=1e-8):
        """SRP_Map_WINDOW_torch (:387-433) -> float64 tensor (G,)."""
        nfft = self.nfft
        step = window // 2
        T = signal.shape[1]
        srp = torch.zeros(self.geom.grids.shape[0])
        for j in range(T // step - 1):
            if j * step + window > T:
                break
            win = signal[:, j * step:j * step + window]
            X = np.array([analysis(x, nfft, nfft // 4).T for x in win])
            X_ts = torch.tensor(X)
            absX = torch.abs(X_ts)
            absX[absX < tol] = tol
            pX = X_ts / absX
            n = pX.shape[2]
            CC = torch.stack([torch.mm(pX[:, k, :], torch.conj(pX[:, k, :]).T) / n for k in self.freq_bins])
            CC_flat = CC.reshape((-1, CC.shape[-2] * CC.shape[-2]))[:, self.mask_triu]
            res = torch.real(CC_flat) * self.mode_mat_flat_real - torch.imag(CC_flat) * self.mode_mat_flat_imag
            res = torch.sum(res, (1, 2)) / self.freq_bins.shape[0] / CC_flat.shape[1]
            srp = torch.maximum(srp, res)
        return srp

    def apply_srp_phat(self, mix):
        """Mic_Array.Apply_SRP_PHAT (sep/Mic_Array.py:152-194) -> (patches, map float64 numpy)."""
        g = self.geom
        win = srp_oracle.window_length(mix.shape[1])
        m = self.srp_map(np.asarray(mix), win).numpy()
        pm, pi = prune_oracle.fill_powermap(m, g.clusters, (g.Lx, g.Ly, g.Lz))
        ids = prune_oracle.find_valid_peaks(pm, pi, g.dis_matrix, float(m.max()), len(g.clusters))
        patches = prune_oracle.local_source_adaptive(m, ids, g.grids, self.cluster_offsets, g.mic_pos.shape[0], g)
        return patches, m

    @staticmethod
    def shift_stack(mix, patches, batch=128):
        """The shift loop of shift_and_sep (network.py:58, 75-83) on CPU tensors."""
        mix = torch.as_tensor(mix)
        M, T = mix.shape
        data = torch.zeros((batch, M, T))
        n_cols = T
        for i in range(0, len(patches), batch):
            chunk = patches[i:i + batch]
            for j, p in enumerate(chunk):
                shifts = -torch.Tensor([0, *p.sample_offset]).unsqueeze(1)
                shifts = torch.round(shifts).long()
                ar = torch.arange(n_cols).view((1, n_cols)).repeat((M, 1))
                data[j] = torch.gather(mix, 1, (ar - shifts) % n_cols)
        return data
