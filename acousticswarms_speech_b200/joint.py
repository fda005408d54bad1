"""The localisation half of the reference's ``JointModel`` (sep/training/JointModel/network.py:106-199): ``setup``
(:125-137, one ``Mic_Array`` per geometry, reused while the geometry string is unchanged) and
``localize_by_separation`` (:151-199), the caller of the accelerated path:

    Apply_SRP_PHAT -> Spotform_Big_Patch -> Spotform_Small_Patch_Parallel -> Clustering_new

The separation half (``separate_by_localization``, the second network and the checkpoints ``load_model_from_exp``
reads) is out of scope; ``spot_model`` is whatever module the caller has (wrapped in ``DataParallelSpotModel`` like
the reference does when ``use_spot_dataparallel`` is set).
"""
import time

import numpy as np
import torch.nn as nn

from .mic_array import Mic_Array
from .spot import DataParallelSpotModel


class JointLocalizer:
    def __init__(self, spot_model, use_spot_dataparallel=True, use_fp16=False, spot_batch_size=128, device=None):
        if use_spot_dataparallel and isinstance(spot_model, nn.Module) and not isinstance(spot_model, DataParallelSpotModel):
            spot_model = DataParallelSpotModel(spot_model, use_fp16, spot_batch_size, device=device)
        self.spot_model = spot_model
        self.device = device
        self.times = [0, 0, 0, 0, 0]
        self.previous_config = None
        self.Mic_processor = None

    def setup(self, mic_positions, speaker_range, cached=False, cached_folder=None):
        """:125-137 -- the array is rebuilt only when the geometry changes."""
        mic_positions = np.asarray(mic_positions)
        current_config = "~".join(f"{x:.05f}" for x in mic_positions.flatten()) + "|" + \
                         "~".join(f"{x:.05f}" for x in speaker_range)
        if current_config != self.previous_config:
            self.Mic_processor = Mic_Array(mic_positions, Spk_Range=speaker_range, cached=cached,
                                           cached_folder=cached_folder, device=self.device)
            self.previous_config = current_config

    def localize_by_separation(self, mix_data, run_demo_folder=None, tracking=False):
        """:151-199 -> (patch_final, audio_final, SRP_drop, stage1_drop, spot_times)."""
        assert self.previous_config is not None, \
            "Microphone positions and spk range were not provided, did you forget to call .setup()?"
        if tracking:
            raise NotImplementedError("Clustering_tracking_new does not exist in the reference's Mic_Array either "
                                      "(sep/training/JointModel/network.py:190 is dead code)")
        t0 = time.time()
        patch_list, simple_pos = self.Mic_processor.Apply_SRP_PHAT(mix_data)
        self.times[0] = time.time() - t0
        if len(patch_list) <= 0:
            return [], [], 0, 0, 0
        t0 = time.time()
        patch_list = self.Mic_processor.Spotform_Big_Patch(mix_data, patch_list, self.spot_model)
        self.times[1] = time.time() - t0
        if len(patch_list) <= 0:
            return [], [], 0, 0, 0
        t0 = time.time()
        output_pair = self.Mic_processor.Spotform_Small_Patch_Parallel(mix_data, patch_list, self.spot_model,
                                                                       run_demo_folder=run_demo_folder)
        self.times[2] = time.time() - t0
        if len(output_pair) <= 0:
            return [], [], 0, 0, 0
        t0 = time.time()
        audio_final, patch_final, spot_times, _ = self.Mic_processor.Clustering_new(output_pair)
        self.times[3] = time.time() - t0
        if len(patch_final) <= 0:
            return [], [], 0, 0, 0
        return patch_final, np.array(audio_final), 0, 0, spot_times
