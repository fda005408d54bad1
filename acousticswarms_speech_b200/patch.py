"""``Patch``: the TDoA hypercube record passed between the stages of the path.

Mirror of the reference's value type (sep/Traditional_SP/Patch_3D.py:3-93): same attribute and
method names, same quirks (``check_out`` assigns floats into the int64 arrays, i.e. truncates).
"""
import numpy as np


class Patch(object):
    def __init__(self, sample_offset, width_list, area_points, peak_pos=None, area_fn=None, centre=None):
        self.sample_offset = sample_offset          # int64 (M-1,): hypercube centre, samples vs mic 0
        self.width_list = np.copy(width_list)       # full width per dimension (8 coarse, 4 fine, 2 centre)
        self._area_points = area_points             # (3, n) 1 cm voxels inside, or None
        self._area_fn = area_fn                     # builds area_points on first use (device-selected patches)
        self._centre = centre                       # mean of area_points when the device already reduced it
        self.num_pair = sample_offset.shape[0]
        self.peak_pos = peak_pos

    @property
    def area_points(self):
        """(3, n) 1 cm voxels whose TDoA vector lies in the hypercube.  Patches selected on the device
        carry a builder instead of the points (SRP_Prunning.py:41-61 evaluated on first use)."""
        if self._area_points is None and self._area_fn is not None:
            self._area_points = self._area_fn()
            self._area_fn = None
        return self._area_points

    @area_points.setter
    def area_points(self, value):
        self._area_points = value
        self._area_fn = None
        self._centre = None

    def area_points_getter(self):
        """A callable returning ``area_points`` (building them on first call).  Lets callers hold on to a patch's
        points without forcing the 1 cm scan before they are needed."""
        return lambda: self.area_points

    def area_size(self):
        if self.area_points is None or self.area_points.shape[1] == 0:
            return 0
        return self.area_points.shape[1]

    def center_pos(self):
        if self.peak_pos is not None:
            return self.peak_pos
        if self._area_points is None and self._centre is not None:     # asw_subdivide's leaf centre
            return self._centre
        if self.area_points is None or self.area_points.shape[1] == 0:
            return None
        return np.mean(self.area_points, axis=1)

    def _bound(self, i):
        half = self.width_list[i] / 2
        return self.sample_offset[i] - half - 1e-3, self.sample_offset[i] + half + 1e-3

    def hyperbola_general_area(self, X, Y, Z, mic_position, sound_speed, fs):
        """Patch_3D.py:28-38: which points (X, Y, Z) have their TDoA vector inside the hypercube."""
        d0 = (((X - mic_position[0, 0]) ** 2 + (Y - mic_position[0, 1]) ** 2
               + (Z - mic_position[0, 2]) ** 2) ** 0.5) / sound_speed * fs
        z = 1
        for i in range(mic_position.shape[0] - 1):
            di = (((X - mic_position[i + 1, 0]) ** 2 + (Y - mic_position[i + 1, 1]) ** 2
                   + (Z - mic_position[i + 1, 2]) ** 2) ** 0.5) / sound_speed * fs
            lo, hi = self._bound(i)
            z = z & ((di - d0) >= lo) & ((di - d0) <= hi)
        return z.astype(int)

    def hyperbola_sample(self, offset):
        """Patch_3D.py:40-47: same test on precomputed TDoA rows ``offset`` (M-1, n)."""
        z = 1
        for i in range(offset.shape[0]):
            lo, hi = self._bound(i)
            z = z & (offset[i, :] >= lo) & (offset[i, :] <= hi)
        return z.astype(int)

    def check_gt(self, sample_offsets_gt):
        """Patch_3D.py:50-66 (debug helper)."""
        for i in range(sample_offsets_gt.shape[1]):
            if all(abs(sample_offsets_gt[j, i] - self.sample_offset[j]) <= self.width_list[j] / 2 + 1
                   for j in range(self.num_pair)):
                return True
        return False

    def check_out(self, upper_bound_pairwise):
        """Patch_3D.py:69-87: halve out-of-bound dimensions towards the physically possible range."""
        for i in range(self.num_pair):
            ub = upper_bound_pairwise[i]
            while not (abs(self.sample_offset[i]) <= ub or self.width_list[i] <= 4):
                res = self.width_list[i]
                if self.sample_offset[i] > ub:
                    self.sample_offset[i] = self.sample_offset[i] - res / 4
                elif self.sample_offset[i] < -ub:
                    self.sample_offset[i] = self.sample_offset[i] + res / 4
                self.width_list[i] = res / 2

    def check_ready_Spotforming(self, MIN_TOLERANCE):
        for i in range(self.num_pair):
            if self.width_list[i] > MIN_TOLERANCE:
                return False, i
        return True, -1
