"""ctypes binding of libasw.so (include/asw.h).  There is no CPU fallback: if the
library is missing or a call fails, this raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libasw.so")

# every symbol include/asw.h declares (tests check the export list against the header)
SYMBOLS = [
    "asw_version", "asw_last_error", "asw_launch_count",
    "asw_srp_create", "asw_srp_destroy", "asw_srp_score",
    "asw_srp_num_windows", "asw_srp_num_frames",
    "asw_srp_read_cc", "asw_srp_gcc_layout", "asw_srp_read_gcc",
    "asw_map_topk", "asw_shift_stack", "asw_shift_stack_norm",
    "asw_geometry_cluster",
    "asw_peaks_create", "asw_peaks_destroy", "asw_peaks_find",
    "asw_select_create", "asw_select_destroy", "asw_select_patches", "asw_subdivide",
    "asw_build_shift_table", "asw_shift_stack_counted", "asw_pcm16_to_f32", "asw_patch_powers",
    "asw_srp_set_frame_mode", "asw_srp_num_frames_mode", "asw_select_set_grid1", "asw_srp_set_stft_path", "asw_srp_gcc", "asw_srp_gather",
    "asw_corr_create", "asw_corr_destroy", "asw_corr_table_len", "asw_corr_tables", "asw_shift_stack_norm_tab",
    "asw_shift_stack_norm_grouped",
    "asw_build_fine_table",
]


class AswError(RuntimeError):
    pass


class AswCapacityError(AswError):
    """A device list was too short for the result (the call can be repeated with a larger capacity)."""


_lib = None


def load():
    """Load libasw.so (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AswError(
            f"{LIB_PATH} is missing: build it with `python -m acousticswarms_speech_b200.build` "
            "(or __graft_entry__.build()).  This package has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i32, f32 = c.c_void_p, c.c_int, c.c_float
    lib.asw_version.restype = i32
    lib.asw_last_error.restype = c.c_char_p
    lib.asw_launch_count.restype = c.c_longlong
    lib.asw_srp_create.argtypes = [c.POINTER(vp), i32, i32, i32, vp, i32, i32, i32, i32, f32, i32]
    lib.asw_srp_destroy.argtypes = [vp]
    lib.asw_srp_score.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.asw_srp_num_windows.argtypes = [i32, i32]
    lib.asw_srp_num_frames.argtypes = [i32, i32, i32]
    lib.asw_srp_read_cc.argtypes = [vp, vp, vp]
    lib.asw_srp_gcc_layout.argtypes = [vp, vp, vp, vp, c.POINTER(i32), c.POINTER(i32)]
    lib.asw_srp_read_gcc.argtypes = [vp, vp, vp]
    lib.asw_map_topk.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp]
    lib.asw_shift_stack.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp]
    lib.asw_shift_stack_norm.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.asw_geometry_cluster.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.asw_peaks_create.argtypes = [c.POINTER(vp), i32, i32, i32, i32, i32, vp, vp, vp, c.c_double]
    lib.asw_peaks_destroy.argtypes = [vp]
    lib.asw_peaks_find.argtypes = [vp, vp, i32, vp, i32, vp, vp, vp]
    lib.asw_select_create.argtypes = [c.POINTER(vp), i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32,
                                      i32, vp, vp, vp, vp, i32, i32, i32]
    lib.asw_select_destroy.argtypes = [vp]
    lib.asw_select_patches.argtypes = [vp, vp, vp, i32, vp, i32, vp, vp, vp, vp, i32, vp]
    lib.asw_subdivide.argtypes = [vp, vp, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp]
    lib.asw_select_set_grid1.argtypes = [vp, vp, vp, vp]
    lib.asw_build_shift_table.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, i32, vp]
    lib.asw_shift_stack_counted.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]
    lib.asw_pcm16_to_f32.argtypes = [vp, vp, c.c_longlong, vp]
    lib.asw_srp_set_frame_mode.argtypes = [vp, i32]
    lib.asw_srp_set_stft_path.argtypes = [vp, i32]
    lib.asw_srp_gcc.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.asw_srp_gather.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.asw_srp_num_frames_mode.argtypes = [i32, i32, i32, i32]
    lib.asw_patch_powers.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.asw_corr_create.argtypes = [c.POINTER(vp), i32, i32, i32]
    lib.asw_corr_destroy.argtypes = [vp]
    lib.asw_corr_table_len.argtypes = [vp]
    lib.asw_corr_tables.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.asw_shift_stack_norm_tab.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp, i32, vp]
    lib.asw_shift_stack_norm_grouped.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp]
    lib.asw_build_fine_table.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name, None)
        if fn is not None and name not in ("asw_last_error", "asw_launch_count"):
            fn.restype = i32
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().asw_last_error().decode("utf-8", "replace")
        raise AswError(f"libasw error {rc}: {msg}")


def launch_count():
    return int(load().asw_launch_count())
