"""Constants of the SRP-PHAT + shift-stack path.

Mirror of the values the reference fixes in sep/helpers/constants.py (:7-8,
:11, :23-27, :32-41) and the literals in sep/Mic_Array.py:120, 160-163 and
sep/Traditional_SP/SRP_Prunning.py:17, 112-114, 174.  The sampling rate is a
parameter everywhere in this package (BASELINE.json quotes 44.1 kHz, the
reference hard-codes 48 kHz); ``FS`` is only the default.
"""
import numpy as np

SPEED_OF_SOUND = 343.0
FS = 48000

MAX_SHIFTS = [2, 4]
INIT_WIDTH = 8
BIN0 = 2
BIN1 = 200
freq_bins = np.arange(BIN0, BIN1)
n_fft = 2048
HOP = n_fft // 4

MIN_AREA = 400
MIN_WIDTH = 3
MIN_TOLERANCE = 4
MAX_BIG_PATCH = 30
MIN_WIDTH_REQUIRED = 2
USE_RELATIVE_SPOT_POWER = False
SPOT_POWER_THRESHOLD1 = 0.008
SPOT_POWER_THRESHOLD2 = 0.01
SI_SNR_POWER_THRESHOLD = 4e-3

SRP_THRESHOLDS = (0.15, 0.015, 0.05)   # sep/Mic_Array.py:120  [ratio, floor, ceiling]
SRP_THRESHOLD_RATIO = 4                # SRP_Prunning.py:500
ERR_TOLERANCE = 0.2                    # SRP_Prunning.py:17
SAMPLE_RESOLUTION = 4                  # SRP_Prunning.py:112
KEEPOUT = 0.2                          # SRP_Prunning.py:174
PHAT_TOL = 1e-8                        # SRP_Prunning.py:384
SPOT_BATCH_SIZE = 128                  # sep/training/JointModel/network.py:112


def window_length(T):
    """sep/Mic_Array.py:160-163."""
    return 36000 if T >= 72000 else 24000


def window_starts(T, window):
    """sep/Traditional_SP/SRP_Prunning.py:393-403."""
    step = window // 2
    out = []
    for j in range(T // step - 1):
        if j * step + window > T:
            break
        out.append(j * step)
    return out


def frames_per_window(window, nfft=n_fft, hop=HOP):
    """Frame count of the rectangular-window STFT (pyroomacoustics semantics, A1)."""
    return (window - nfft) // hop + 1
