"""``Mic_Array``: host orchestration of localisation-by-separation's front end.

Mirror of the reference's sep/Mic_Array.py for the three drop-in methods of the hot path --
``Apply_SRP_PHAT`` (:152-194), ``Spotform_Big_Patch`` (:196-222) and
``Spotform_Small_Patch_Parallel`` (:225-395) -- with the same signatures, return values and
per-instance state (``SRP_node``, ``Relative_Threshold``, counters).  Scoring and shift-stack run in
libasw.so (CUDA); the reference's debug monitor / plotting is dropped.
"""
import numpy as np
import torch

from .constants import (FS, INIT_WIDTH, SPEED_OF_SOUND, SPOT_POWER_THRESHOLD2, SRP_THRESHOLDS,
                        USE_RELATIVE_SPOT_POWER, freq_bins, n_fft, window_length)
from .local_utils import (_tdoa_rows, binary_search_baseline, max_avg_power, search_area, si_sdr, split_wav,
                          split_wise_sisdr)
from .spot import si_sdr_from_gram
from .patch import Patch
from .srp_phat import SRP_PHAT


def check_sisnr_win(sisnr_list, SISNR_THRESHOLD=-2, SISNR_THRESHOLD2=-7):
    """Window-wise SI-SDR check (Mic_Array.py:18-28): some window above the first threshold, none below the second."""
    same, none_below = False, True
    for value in sisnr_list:
        if value > SISNR_THRESHOLD:
            same = True
        if value < SISNR_THRESHOLD2:
            none_below = False
    return same and none_below


def weight_mean_pos(patch_list, powers, id_lists):
    """Power-weighted mean position / offsets of a cluster of fine patches (Mic_Array.py:32-47)."""
    total_pos = np.zeros((3,))
    total_power = 0
    max_power = powers[id_lists[0]]
    total_offsets = np.zeros(patch_list[0].sample_offset.shape)
    for _id in id_lists:
        if powers[_id] < max_power * 0.75:
            continue
        total_pos += powers[_id] * patch_list[_id].center_pos()
        total_offsets += powers[_id] * patch_list[_id].sample_offset
        total_power += powers[_id]
    return total_pos / total_power, total_offsets / total_power


def find_merge_center(merged_offests, init_area, mic_positions, Big_patch_center, fs=FS):
    """Mic_Array.py:50-81 (the widening loop only ever tries factor 0, as in the reference).  ``fs``: the rate the
    offsets are in (the reference hard-codes 48 kHz)."""
    num_pair = mic_positions.shape[0] - 1
    patch_center = Patch(merged_offests, [3 for _ in range(num_pair)], None)
    area = patch_center.hyperbola_general_area(init_area[0, :], init_area[1, :], init_area[2, :], mic_positions,
                                               SPEED_OF_SOUND, fs) == 1
    if np.sum(area) == 0:
        patch_center.width_list = [3 for _ in range(num_pair)]
        area = patch_center.hyperbola_general_area(init_area[0, :], init_area[1, :], init_area[2, :],
                                                   mic_positions, SPEED_OF_SOUND, fs) == 1
        if np.sum(area) > 0:
            patch_center.area_points = init_area[:, area]
        else:
            patch_center.peak_pos = Big_patch_center
    else:
        patch_center.area_points = init_area[:, area]
    return patch_center


class Mic_Array(object):
    def __init__(self, mic_positions, demo=False, Spk_Range=None, grid_size=0.05, Prone_method="SRP",
                 MIN_TRIGGER_POWER=0.5, SRP_fast=False, cached=False, cached_folder=None, fs=FS, device=None):
        if Prone_method != "SRP":
            raise NotImplementedError("only the default SRP pruner is on the accelerated path "
                                      "(MUSIC/TOPS are out of scope, sep/Mic_Array.py:167-170)")
        self.Prone_method = Prone_method
        self.MIN_TRIGGER_POWER = MIN_TRIGGER_POWER
        self.visual_save = False
        self.Range_spk = Spk_Range
        self.fs = fs
        self.mic_positions = np.asarray(mic_positions, dtype=np.float64)
        self.num_mic = self.mic_positions.shape[0]
        self.upper_bound_pairwise = np.zeros((self.num_mic - 1,))
        for i in range(1, self.num_mic):
            self.upper_bound_pairwise[i - 1] = (np.linalg.norm(self.mic_positions[i] - self.mic_positions[0])
                                                + 0.08) / SPEED_OF_SOUND * fs
        # SRP_fast selected the torch device in the reference (:123-130); here the path is CUDA-only
        self.SRP_node = SRP_PHAT(mic_pos=self.mic_positions, freq_bins=freq_bins, Range_spk=Spk_Range,
                                 grid_size=grid_size, FS=fs, n_fft=n_fft, threshold=list(SRP_THRESHOLDS),
                                 WIDTH=INIT_WIDTH, device=device, cached=cached, cached_name=cached_folder)
        self.original_times = 0
        self.spotforming_times = 0

    def Apply_SRP_PHAT(self, mix_data):
        """(M, T) float32 tensor -> (patch_list, simple_pos)   (Mic_Array.py:152-194)."""
        self.SRP_node.reset()
        self.spotforming_times = 0
        self.original_times = 0
        WIN_SIZE = window_length(mix_data.shape[1])
        self.SRP_node.SRP_Map_WINDOW_new(mix_data, window=WIN_SIZE)
        if self.SRP_node.native is not None:
            patch_list = self.SRP_node.local_source_adaptive_device()
        else:
            patch_list = self.SRP_node.local_source_adaptive()
        return patch_list, np.zeros((3, 3))

    def Spotform_Big_Patch(self, mix_data, patch_list, spot_model):
        """Mic_Array.py:196-222."""
        self.big_spotforming_times = len(patch_list)
        candidate_finished, powers_with_dis, Relative_Threshold = binary_search_baseline(
            mix_data, spot_model, patch_list, self.mic_positions)
        self.Relative_Threshold = Relative_Threshold
        return candidate_finished

    def _search_area_device(self, candidates):
        """search_area (local_utils_3d.py:212-246) for all candidates in one launch (asw_subdivide).  The
        candidates are mutated exactly as the reference's check_out does; a leaf's ``area_points`` is rebuilt on
        first use from the candidate's points and the leaf's TDoA box, in the reference's order."""
        import torch
        from . import native
        node = self.SRP_node
        D = self.num_mic - 1
        dev = node.device
        centres = np.stack([np.asarray(c.sample_offset, dtype=np.int32) for c in candidates])
        widths = np.array([int(c.width_list[0]) for c in candidates], dtype=np.int32)
        for c in candidates:
            if not np.all(np.asarray(c.width_list) == c.width_list[0]):
                raise native._lib.AswError("coarse patches must have one width in every dimension")
        for max_leaves in (128, 1024, 8192):     # many-mic arrays split along more dimensions: retry with longer lists
            try:
                cn, off, wid, npts, box, root, centre, members = native.subdivide(
                    node.native_select, torch.from_numpy(centres).to(dev), torch.from_numpy(widths).to(dev),
                    self.upper_bound_pairwise, max_leaves=max_leaves, member_cap=1 << 17)
                break
            except native._lib.AswCapacityError:
                if max_leaves == 8192:
                    raise
        pos1 = node.Pos_1.reshape(-1, 3)
        out = []
        for i, cand in enumerate(candidates):
            if cand._area_points is None and len(members[i]) > 0:
                # the kernel already enumerated the candidate's 1 cm voxels; ascending voxel index is the order of
                # hyperbola_area_init's np.nonzero, so this is the same (3, n) array without the host scan
                cand._area_fn = (lambda m=members[i]: pos1[m].T)
            root_box_lo = centres[i].astype(np.float64) - (float(widths[i]) + 0.2) / 2
            untouched = cn[i] == 1 and np.array_equal(box[i, 0, 0], root_box_lo)    # the root itself is the leaf
            parent = cand.area_points_getter()         # bound before check_out mutates the candidate; built on demand
            cand.sample_offset[:] = root[i, 0]
            cand.width_list[:] = root[i, 1]
            if untouched:
                out.append([cand])
                continue

            rows_cache = {}

            def builder(lo, hi, parent=parent, cache=rows_cache):
                def build():
                    area = parent()
                    if "rows" not in cache:              # the candidate's TDoA rows, once for all its leaves
                        cache["rows"] = _tdoa_rows(area, self.mic_positions, self.fs)
                    rows = cache["rows"]
                    keep = np.all((rows >= lo[:, None]) & (rows <= hi[:, None]), axis=0)
                    return area[:, keep]
                return build

            # an empty leaf has no centre (center_pos() -> None, like the host path); the others carry the device mean
            out.append([Patch(off[i, l].astype(np.int64), wid[i, l].astype(np.int64), None, None,
                              area_fn=builder(box[i, l, 0].copy(), box[i, l, 1].copy()),
                              centre=centre[i, l].copy() if npts[i, l] > 0 else None) for l in range(int(cn[i]))])
        return out

    def small_patch_list(self, candidate_finished):
        """The patch-list assembly of Spotform_Small_Patch_Parallel (:244-262): fine hypercubes from
        ``search_area`` plus one width-2 centre patch per candidate."""
        width_list0 = [2 for _ in range(self.num_mic - 1)]
        total_patch, patches_indexes, init_area_total, centre_total = [], [0], [], []
        self.spotforming_times = 0
        on_device = (self.SRP_node.native is not None and self.num_mic - 1 <= 31 and len(candidate_finished) > 0)
        fine_lists = self._search_area_device(candidate_finished) if on_device else None
        for ci, cand in enumerate(candidate_finished):
            if on_device:
                patch_processed = fine_lists[ci]
            else:
                patch_processed = search_area([cand], self.mic_positions, self.upper_bound_pairwise, self.fs)
            init_area_total.append(cand.area_points_getter())    # (:246) materialised only for clusters that are output
            patch_center0 = Patch(cand.sample_offset, width_list0, None, cand.peak_pos)
            centre = patch_center0.center_pos()
            centre_total.append(centre)
            if centre is not None:
                patch_processed.append(patch_center0)
            self.spotforming_times += len(patch_processed)
            total_patch.extend(patch_processed)
            patches_indexes.append(self.spotforming_times)
        return total_patch, patches_indexes, init_area_total, centre_total

    def Spotform_Small_Patch_Parallel(self, mix_data, candidate_finished, spot_model, sample_gt=None,
                                      run_demo_folder=None):
        """Mic_Array.py:225-395 -> list of (patch_center, audio, power, tag, offsets, big_label)."""
        output_pair = []
        if USE_RELATIVE_SPOT_POWER:
            thr_new = min([SPOT_POWER_THRESHOLD2, self.Relative_Threshold])
        else:
            thr_new = SPOT_POWER_THRESHOLD2
        total_patch, patches_indexes, init_area_total, centre_total = self.small_patch_list(candidate_finished)
        device_rows = getattr(spot_model, "shift_and_sep_device", None)
        if device_rows is not None:
            # outputs stay on the device: per-row powers come back as small arrays, SI-SDR is evaluated from one
            # Gram matrix of the rows that pass the power gates, and only the audio that is returned is copied
            rows = device_rows(mix_data, total_patch, Strict=1)
            sep_data_total = None
        else:
            rows = None
            sep_data_total = spot_model.shift_and_sep(mix_data, total_patch, Strict=1)
        T = mix_data.shape[-1]

        for i in range(len(patches_indexes) - 1):
            big_offset = candidate_finished[i].sample_offset
            big_label = -1
            if sample_gt is not None:
                for k in range(sample_gt.shape[1]):
                    if np.amax(np.abs(big_offset - sample_gt[:, k])) < 3.5:
                        big_label = k
                        break
            lo, hi = patches_indexes[i], patches_indexes[i + 1]
            patch_processed = total_patch[lo:hi]
            init_area = init_area_total[i]               # callable: the candidate's area_points on demand
            Big_patch_center = centre_total[i]
            if rows is not None:
                powers, powers2 = list(rows.power[lo:hi]), list(rows.maxavg[lo:hi])
            else:
                sep_data = sep_data_total[lo:hi]
                powers, powers2 = [], []
                for j in range(len(patch_processed)):
                    sep_data[j, :] = sep_data[j, :] - np.mean(sep_data[j, :])
                    powers.append(np.sum(sep_data[j, :] ** 2))
                    powers2.append(max_avg_power(sep_data[j, :])[0])
            cpos = candidate_finished[i].center_pos()
            d = np.linalg.norm(cpos - self.mic_positions[0]) if cpos.shape[0] == 3 else 4
            if np.amax(powers2) < thr_new / (1 + d):
                continue
            sort_idx = np.argsort(-1 * np.array(powers))
            SI_SDR_THRESHOLD = -4
            clusters = {}
            MIN_TRIGGER_POWER2 = self.MIN_TRIGGER_POWER / (3 * 48000) * T
            passing = []
            for _id in sort_idx:
                d = np.linalg.norm(patch_processed[_id].center_pos() - self.mic_positions[0])
                if powers2[_id] < thr_new / (1 + d) or powers[_id] < MIN_TRIGGER_POWER2:
                    continue
                passing.append(int(_id))
            if rows is not None:
                gram = rows.gram([lo + _id for _id in passing])
                slot = {_id: k for k, _id in enumerate(passing)}

                def sdr(est, ref):
                    return si_sdr_from_gram(gram[slot[est], slot[est]], gram[slot[ref], slot[ref]], gram[slot[est], slot[ref]])
            else:
                def sdr(est, ref):
                    return si_sdr(sep_data[est, :], sep_data[ref])
            for _id in passing:
                unique = True
                for cluster_id in clusters:
                    final_candidate_id = clusters[cluster_id][0]
                    if sdr(_id, final_candidate_id) > SI_SDR_THRESHOLD:
                        clusters[final_candidate_id].append(_id)
                        unique = False
                        break
                if unique:
                    clusters[_id] = [_id]
            for cluster_id in clusters:
                position, offests = weight_mean_pos(patch_processed, powers, clusters[cluster_id])
                patch_center = find_merge_center(offests, init_area(), self.mic_positions, Big_patch_center, self.fs)
                save_offsets = {"audio_offset": patch_processed[cluster_id].sample_offset,
                                "localization_offset": offests}
                audio = rows.row(lo + cluster_id) if rows is not None else sep_data[cluster_id, :]
                output_pair.append((patch_center, audio, powers[cluster_id],
                                    str(i) + "_" + str(cluster_id), save_offsets, big_label))
        return output_pair

    def Clustering_new(self, output_pair, simple_pos=None, sample_gt=None):
        """Non-maximum suppression over the fine-stage outputs (Mic_Array.py:399-500): candidates by descending power;
        one joins an existing cluster when its SI-SDR against the cluster head exceeds -1 dB, or the window-wise SI-SDRs
        pass ``check_sisnr_win``, or the heads are closer than 0.45 m in the plane; the maximum over all heads of the
        window-wise SI-SDRs can also veto a new cluster.  -> (audio_final, patch_final, spot_times, wrong_spotforming).
        Host code on a handful of rows, downstream of the accelerated path; kept so that
        ``JointModel.localize_by_separation`` (sep/training/JointModel/network.py:151-199) runs on the drop-in class."""
        SI_SDR_THRESHOLD = -1
        candidates = sorted(output_pair, key=lambda x: -x[2])
        clusters = {}
        wrong_spotforming = []
        for _id in range(len(candidates)):
            belong_cluster = -1
            unique = True
            sisnr_seg = []
            big_label = candidates[_id][-1]
            center1 = candidates[_id][0].center_pos()
            audio1 = candidates[_id][1]
            power1 = candidates[_id][2]
            seg_win = split_wav(audio1)
            if len(seg_win) == 0:
                continue
            for cluster_id in clusters:
                final_candidate_id = clusters[cluster_id][0]
                audio2 = candidates[final_candidate_id][1]
                center2 = candidates[final_candidate_id][0].center_pos()
                similarity = si_sdr(audio1, audio2)
                sisdr_list = split_wise_sisdr(audio1, audio2, seg_win)
                sisnr_seg.append(sisdr_list)
                dis = np.linalg.norm(center1[:2] - center2[:2])
                if (similarity > SI_SDR_THRESHOLD) or check_sisnr_win(sisdr_list) or dis < 0.45:
                    clusters[final_candidate_id].append(_id)
                    unique = False
                    belong_cluster = cluster_id
                    break
            if len(sisnr_seg) != 0:
                if check_sisnr_win(np.amax(np.array(sisnr_seg), axis=0), SISNR_THRESHOLD=-1, SISNR_THRESHOLD2=-5):
                    unique = False
            if unique:
                clusters[_id] = [_id]
            elif big_label >= 0 and sample_gt is not None and belong_cluster >= 0:
                final_candidate_id = clusters[belong_cluster][0]
                cluster_label = candidates[final_candidate_id][-1]
                power2 = candidates[final_candidate_id][2]
                offset1 = candidates[final_candidate_id][-2]["audio_offset"]
                delta_offset = (offset1 - sample_gt[:, big_label]).astype(int)
                if cluster_label == -1:
                    wrong_spotforming.append((big_label, cluster_label, delta_offset, power1 / power2))
        patch_final, audio_final = [], []
        for cluster_id in clusters:
            final_candidate_id = clusters[cluster_id][0]
            patch_final.append(candidates[final_candidate_id])
            audio_final.append(candidates[final_candidate_id][1])
        return audio_final, patch_final, self.big_spotforming_times + self.spotforming_times, wrong_spotforming
