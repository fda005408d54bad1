"""Oracle: hypercube (Grid_cluster) table build -- literal CPU restatement.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows sep/Traditional_SP/SRP_Prunning.py:
  * grids / keep-out / distance matrix ............ :133-146, 173-180
  * 5 cm and 1 cm position+TDoA volumes ........... :148-170
  * voxel -> quantised TDoA (``calculate_offset_pair``, ``Map_3D_TDoA``) :257-263, 315-331
  * ``check_valid`` ................................ :266-275
  * BFS clustering (``search_cluster``) ............ :277-313, 333-344
The arithmetic is evaluated voxel by voxel exactly as the reference does (so
round-half-even ties and summation orders agree); only the bookkeeping differs
(a dict instead of a fresh visited volume per cluster, quirk B5).
"""
from collections import deque

import numpy as np


class GeometryOracle:
    def __init__(self, mic_pos, Range_spk, grid_size=0.05, grid_size_z=0.1,
                 C=343.0, FS=48000, sample_resolution=4, build_fine=True):
        mic_pos = np.asarray(mic_pos, dtype=np.float64)
        self.mic_pos = mic_pos
        self.num_mic = mic_pos.shape[0]
        self.mic_center = mic_pos.mean(0)
        self.C, self.FS = C, FS
        self.sample_resolution = sample_resolution
        self.Range_spk = Range_spk
        r = Range_spk
        self.x_grids = np.arange(r[0], r[1], grid_size)
        self.y_grids = np.arange(r[2], r[3], grid_size)
        self.z_grids = np.arange(r[4], r[5], grid_size_z)
        self.Lx, self.Ly, self.Lz = len(self.x_grids), len(self.y_grids), len(self.z_grids)
        self.Axis_range = [[r[0], r[1]], [r[2], r[3]], [r[4], r[5]]]

        # :140-144  horizontal distance of every (x, y) column to the array centre
        dis = np.zeros((self.Lx, self.Ly))
        for ix in range(self.Lx):
            for iy in range(self.Ly):
                p = np.array([self.x_grids[ix], self.y_grids[iy]])
                dis[ix, iy] = np.linalg.norm(p - self.mic_center[:2]) + 1e-8
        self.dis_matrix = dis

        # :173-180  keep-out box around the array
        K = 0.2
        self.array_border = [mic_pos[:, 0].min() - K, mic_pos[:, 1].min() - K,
                             mic_pos[:, 0].max() + K, mic_pos[:, 1].max() + K]

        if build_fine:
            self.Pos_5, self.Offset_5 = self._volume(0.05)
            self.Pos_1, self.Offset_1 = self._volume(0.01)
        self._cluster()

    # :148-170 -- note meshgrid default 'xy' => volumes are indexed [y, x, z] (quirk B3)
    def _volume(self, step):
        r = self.Range_spk
        xx = np.arange(r[0], r[1], step)
        yy = np.arange(r[2], r[3], step)
        zz = np.arange(r[4], r[5], 0.1)
        X, Y, Z = np.meshgrid(xx, yy, zz)
        pos = np.stack((X, Y, Z), axis=3)
        d0 = np.linalg.norm(pos - self.mic_pos[0, :], axis=3) / self.C * self.FS
        offs = [np.linalg.norm(pos - self.mic_pos[i, :], axis=3) / self.C * self.FS - d0
                for i in range(1, self.num_mic)]
        return pos, np.stack(offs, axis=3)

    def check_valid(self, ix, iy, iz):
        if ix < 0 or ix >= self.Lx or iy < 0 or iy >= self.Ly or iz < 0 or iz >= self.Lz:
            return False
        x, y = self.x_grids[ix], self.y_grids[iy]
        b = self.array_border
        return not (x > b[0] and y > b[1] and x < b[2] and y < b[3])

    def voxel_offset(self, pos):
        """:257-263 then :327-329 (quantise to multiples of sample_resolution)."""
        d0 = np.linalg.norm(pos - self.mic_pos[0])
        off = np.array([(np.linalg.norm(pos - self.mic_pos[i]) - d0) / self.C * self.FS
                        for i in range(1, self.num_mic)])
        q = np.round(off / self.sample_resolution).astype(int)
        return q * self.sample_resolution

    def _cluster(self):
        Lx, Ly, Lz, M = self.Lx, self.Ly, self.Lz, self.num_mic
        tree = np.zeros((Lx, Ly, Lz, M), dtype=int)
        for ix in range(Lx):
            for iy in range(Ly):
                if not self.check_valid(ix, iy, 0):
                    continue
                for iz in range(Lz):
                    pos = np.array([self.x_grids[ix], self.y_grids[iy], self.z_grids[iz]])
                    tree[ix, iy, iz, 0] = 1
                    tree[ix, iy, iz, 1:] = self.voxel_offset(pos)
        self.sample_tree_full = tree.copy()

        alive = tree[..., 0].astype(bool).copy()
        clusters = []      # (sample_offset, [positions], [indices])
        for ix in range(Lx):
            for iy in range(Ly):
                for iz in range(Lz):
                    if not alive[ix, iy, iz]:
                        continue
                    clusters.append(self._bfs((ix, iy, iz), tree, alive))
        self.clusters = clusters
        self.grids = np.array([np.mean(c[1], axis=0) for c in clusters])
        self.POWER_INDEX = np.zeros((Lx, Ly, Lz), dtype=int)
        self.member_mask = np.zeros((Lx, Ly, Lz), dtype=bool)
        for i, c in enumerate(clusters):
            for (a, b, d) in c[2]:
                self.POWER_INDEX[a, b, d] = i
                self.member_mask[a, b, d] = True

    def _bfs(self, start, tree, alive):
        ix, iy, iz = start
        alive[ix, iy, iz] = False
        base = tree[ix, iy, iz, 1:]
        seen = {start}
        queue = deque([start])
        members = [[ix, iy, iz]]
        while queue:
            cx, cy, cz = queue.popleft()
            for nx in range(cx - 1, cx + 2):
                for ny in range(cy - 1, cy + 2):
                    for nz in range(cz - 1, cz + 2):
                        if not self.check_valid(nx, ny, nz):
                            continue
                        if (not alive[nx, ny, nz]) or (nx, ny, nz) in seen:
                            continue
                        seen.add((nx, ny, nz))
                        if np.array_equal(base, tree[nx, ny, nz, 1:]):
                            members.append([nx, ny, nz])
                            queue.append((nx, ny, nz))
        for (a, b, d) in members[1:]:
            alive[a, b, d] = False
        pos = [[self.x_grids[a], self.y_grids[b], self.z_grids[d]] for (a, b, d) in members]
        return (base.copy(), pos, members)
