"""ctypes binding of libb200master.so -- mirrors include/b200_master.h one to one.

There is no CPU fallback: if the shared library has not been built, or no CUDA device is
present, importing / creating a handle raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200M_LIB") or os.path.join(os.path.dirname(HERE), "lib", "libb200master.so")     # B200M_LIB: build experiments

OK, ERR_INVALID, ERR_CUDA, ERR_TOO_SHORT, ERR_NOMEM = 0, 1, 2, 3, 4
FMT_S16, FMT_S24, FMT_F32 = 0, 1, 2


class Biquad(C.Structure):
    _fields_ = [("b0", C.c_double), ("b1", C.c_double), ("b2", C.c_double), ("a1", C.c_double), ("a2", C.c_double)]


class Settings(C.Structure):
    _fields_ = [("saturation", C.c_double), ("bass_boost", C.c_double), ("mid_cut", C.c_double),
                ("presence_boost", C.c_double), ("treble_boost", C.c_double), ("width", C.c_double),
                ("multiband", C.c_int32), ("has_lufs", C.c_int32),
                ("low_thresh", C.c_double), ("low_ratio", C.c_double), ("mid_thresh", C.c_double),
                ("mid_ratio", C.c_double), ("high_thresh", C.c_double), ("high_ratio", C.c_double),
                ("lufs", C.c_double)]


class Band(C.Structure):
    _fields_ = [("thresh_rms", C.c_double), ("attack_frames", C.c_double), ("release_frames", C.c_double),
                ("slope", C.c_double), ("look_frames", C.c_int32), ("reserved", C.c_int32)]


class Plan(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("channels", C.c_int32), ("sat_on", C.c_int32),
                ("sat_clean", C.c_float), ("sat_mix", C.c_float), ("sat_drive", C.c_float),
                ("n_eq", C.c_int32), ("width_on", C.c_int32),
                ("eq", Biquad * 4), ("width", C.c_double),
                ("multiband", C.c_int32), ("has_lufs", C.c_int32),
                ("lp", Biquad * 2), ("hp", Biquad * 2), ("band", Band * 3), ("kw", Biquad * 2),
                ("lufs", C.c_double),
                ("sat_lut", C.c_void_p), ("sat_lut_key", C.c_uint64)]

ABI_VERSION = 2


_lib = None

EXPORTS = [
    "b200m_abi_version", "b200m_create", "b200m_destroy", "b200m_last_error", "b200m_set_stream",
    "b200m_synchronize", "b200m_set_workspace_limit", "b200m_launch_count", "b200m_set_profiling",
    "b200m_kernel_time_ms", "b200m_reset_profile", "b200m_set_recur_tiling", "b200m_recur_stats", "b200m_set_segment_tiles", "b200m_set_pipeline", "b200m_set_chain_kernel", "b200m_set_pipeline_shape", "b200m_plan_from_settings", "b200m_master_batch", "b200m_master_batch_targets", "b200m_master_batch_wav", "b200m_wav_header",
    "b200m_pcm16_to_float", "b200m_float_to_pcm16", "b200m_saturation", "b200m_saturation_pcm", "b200m_stereo_width",
    "b200m_sosfilt", "b200m_multiband", "b200m_compress_dynamic_range", "b200m_integrated_loudness",
    "b200m_normalize_to_lufs", "b200m_soft_limiter",
    "b200m_stage_pcm", "b200m_slice_halo", "b200m_slice_chain", "b200m_slice_energies", "b200m_track_blocks",
    "b200m_gate", "b200m_slice_final",
]


def load():
    """dlopen libb200master.so (built in-tree by build.py) and declare prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python python-audio-mastering_b200/build.py` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
    lib.b200m_abi_version.restype = C.c_int
    lib.b200m_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.b200m_destroy.argtypes = [vp]
    lib.b200m_destroy.restype = None
    lib.b200m_last_error.argtypes = [vp]
    lib.b200m_last_error.restype = C.c_char_p
    lib.b200m_set_stream.argtypes = [vp, vp]
    lib.b200m_synchronize.argtypes = [vp]
    lib.b200m_set_workspace_limit.argtypes = [vp, i64]
    lib.b200m_launch_count.argtypes = [vp]
    lib.b200m_launch_count.restype = i64
    lib.b200m_set_profiling.argtypes = [vp, C.c_int]
    lib.b200m_kernel_time_ms.argtypes = [vp, C.c_char_p, C.POINTER(dbl), C.POINTER(i64)]
    lib.b200m_reset_profile.argtypes = [vp]
    lib.b200m_set_recur_tiling.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.b200m_set_segment_tiles.argtypes = [vp, C.c_int, C.c_int]
    lib.b200m_set_pipeline.argtypes = [vp, C.c_int]
    lib.b200m_set_chain_kernel.argtypes = [vp, C.c_int]
    lib.b200m_set_pipeline_shape.argtypes = [vp, C.c_int, C.c_int]
    lib.b200m_recur_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.c_int]
    lib.b200m_plan_from_settings.argtypes = [C.POINTER(Settings), C.c_int, C.c_int, C.POINTER(Plan)]
    lib.b200m_master_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                                       C.POINTER(Plan), C.c_int, vp, vp, C.c_int, vp, vp]
    lib.b200m_master_batch_targets.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                                               C.POINTER(Plan), C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, vp]
    lib.b200m_master_batch_wav.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                                           C.POINTER(Plan), C.c_int, vp, vp, vp, C.c_int, vp, vp]
    lib.b200m_wav_header.argtypes = [C.c_int, C.c_int, i64, vp]
    lib.b200m_pcm16_to_float.argtypes = [vp, vp, i64, vp]
    lib.b200m_float_to_pcm16.argtypes = [vp, vp, C.c_int, i64, vp]
    lib.b200m_saturation.argtypes = [vp, vp, i64, dbl, vp]
    lib.b200m_saturation_pcm.argtypes = [vp, vp, i64, vp, C.c_uint64, vp]
    lib.b200m_stereo_width.argtypes = [vp, vp, C.c_int, i64, dbl, vp]
    lib.b200m_sosfilt.argtypes = [vp, C.POINTER(Biquad), C.c_int, vp, C.c_int, i64, C.c_int, vp]
    lib.b200m_multiband.argtypes = [vp, C.POINTER(Plan), vp, i64, vp]
    lib.b200m_compress_dynamic_range.argtypes = [vp, vp, i64, C.c_int, C.POINTER(Band), vp, vp, vp]
    lib.b200m_integrated_loudness.argtypes = [vp, C.POINTER(Biquad), vp, i64, C.c_int, C.POINTER(dbl)]
    lib.b200m_normalize_to_lufs.argtypes = [vp, C.POINTER(Biquad), vp, i64, C.c_int, C.c_int, dbl, vp,
                                            C.POINTER(dbl), C.POINTER(dbl)]
    lib.b200m_soft_limiter.argtypes = [vp, vp, C.c_int, i64, dbl, vp]
    lib.b200m_stage_pcm.argtypes = [vp, vp, C.c_int, i64, vp]
    lib.b200m_slice_halo.argtypes = [C.POINTER(Plan), i64, C.POINTER(i64), C.POINTER(i64)]
    lib.b200m_slice_chain.argtypes = [vp, vp, i64, i64, C.POINTER(Plan), vp]
    lib.b200m_slice_energies.argtypes = [vp, vp, i64, i64, i64, i64, i64, C.POINTER(Plan), vp,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.b200m_track_blocks.argtypes = [i64, C.c_int]
    lib.b200m_gate.argtypes = [vp, vp, C.c_int32, C.POINTER(Plan), C.POINTER(dbl), C.POINTER(dbl)]
    lib.b200m_slice_final.argtypes = [vp, vp, i64, C.POINTER(Plan), C.c_int, dbl, vp]
    for name in EXPORTS:
        if name not in ("b200m_destroy", "b200m_last_error", "b200m_launch_count"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def check(lib, handle, rc):
    """Map a C-ABI error code to the exception the reference's callers would see."""
    if rc == OK:
        return
    msg = (lib.b200m_last_error(handle) or b"").decode()
    if rc in (ERR_INVALID, ERR_TOO_SHORT):
        raise ValueError(msg)              # pyloudnorm raises ValueError for short audio
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
