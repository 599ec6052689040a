"""ctypes binding of include/talfe.h.  Loading fails loudly: there is no CPU or PyTorch fallback."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

from . import _build

# enums of include/talfe.h
F32, F16, I16 = 0, 1, 2
NORM_NONE, NORM_BATCH_MEAN, NORM_ROW_MEAN, NORM_ROW_MEL_MEAN, NORM_ROW_MEL_MEANVAR = 0, 1, 2, 3, 4
LAYOUT_TM, LAYOUT_MT = 0, 1
ERR_TOO_SHORT = -2

EXPECTED_VERSION = 105     # TALFE_VERSION of include/talfe.h this binding was written against

EXPORTED = [
    "talfe_version", "talfe_job_size", "talfe_probe_fp32_fma_rate", "talfe_launches_per_forward",
    "talfe_strerror", "talfe_last_cuda_error", "talfe_num_frames", "talfe_plan_create", "talfe_plan_create_ex",
    "talfe_plan_geometry", "talfe_plan_num_frames", "talfe_resample",
    "talfe_plan_destroy", "talfe_plan_n_mels", "talfe_workspace_bytes", "talfe_run", "talfe_logmel_forward",
    "talfe_apply_stats", "talfe_allreduce_stats", "talfe_synth_fill", "talfe_detect_padding", "talfe_stream_staging_bytes",
    "talfe_stream_episode",
]


class Job(ctypes.Structure):
    """struct talfe_job (include/talfe.h) — field order and types must match exactly."""
    _fields_ = [
        ("wave", c_void_p), ("wave_dtype", c_int32), ("norm", c_int32),
        ("batch", c_int64), ("row_stride", c_int64), ("buf_len", c_int64), ("origin", c_int64),
        ("total_len", c_int64), ("lens", c_void_p), ("frame0", c_int64), ("n_frames", c_int64),
        ("out", c_void_p), ("out_row_stride", c_int64), ("out_layout", c_int32), ("accumulate_stats", c_int32),
        ("eps", c_float), ("defer_normalise", c_int32), ("stats", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t), ("stream", c_void_p),
        ("out_offsets", c_void_p), ("freq_bands", c_void_p), ("time_bands", c_void_p), ("n_bands", c_int32),
        ("given_stats", c_void_p), ("lens_are_padding_hint", c_int32),
    ]


_LIB = None


def stats_doubles(n_mels: int) -> int:
    return 3 + 2 * n_mels


def load() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("TALFE_LIB")                                  # TALFE_LIB: an experiment build (tools/_abl/)
    if path:
        if not os.path.isfile(path):
            raise FileNotFoundError(f"TALFE_LIB={path} does not exist")
    else:
        # (re)build where a toolchain exists and a source is newer than the binary; the GPU box receives the prebuilt
        # .so together with its sources (same mtimes), so nothing is compiled there
        path = _build.build_if_possible()
    lib = ctypes.CDLL(path)
    lib.talfe_version.restype = c_int
    if os.environ.get("TALFE_ABI_CHECK") == "0":                        # A/B against an older build (tools/ab_libs.py)
        return _finish_binding(lib, optional=True)
    # a stale or foreign binary must fail here, not pass wrong pointers later
    if not hasattr(lib, "talfe_job_size"):
        raise RuntimeError(f"{path} predates this binding (no talfe_job_size): rebuild with tal_asrd_b200._build.build(force=True)")
    lib.talfe_job_size.restype = c_size_t
    if lib.talfe_version() != EXPECTED_VERSION or lib.talfe_job_size() != ctypes.sizeof(Job):
        raise RuntimeError(f"{path}: TALFE_VERSION {lib.talfe_version()} / sizeof(talfe_job) {lib.talfe_job_size()} do not match "
                           f"this binding ({EXPECTED_VERSION} / {ctypes.sizeof(Job)}): rebuild the library")
    lib.talfe_probe_fp32_fma_rate.restype = c_int
    lib.talfe_probe_fp32_fma_rate.argtypes = [c_int, POINTER(c_double)]
    lib.talfe_launches_per_forward.restype = c_int
    lib.talfe_launches_per_forward.argtypes = [c_void_p, c_int64, c_int64]
    return _finish_binding(lib)


def _finish_binding(lib, optional: bool = False):
    global _LIB
    lib.talfe_version.restype = c_int
    lib.talfe_strerror.restype = c_char_p
    lib.talfe_strerror.argtypes = [c_int]
    lib.talfe_last_cuda_error.restype = c_int
    lib.talfe_num_frames.restype = c_int64
    lib.talfe_num_frames.argtypes = [c_int64]
    lib.talfe_plan_create.restype = c_int
    lib.talfe_plan_create.argtypes = [POINTER(c_void_p), c_int, c_int, c_void_p, c_void_p]
    if hasattr(lib, "talfe_plan_create_ex"):              # (absent from pre-round-2 builds loaded for A/B runs)
        lib.talfe_plan_create_ex.restype = c_int
        lib.talfe_plan_create_ex.argtypes = [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_void_p, c_void_p]
        lib.talfe_plan_geometry.restype = c_int
        lib.talfe_plan_geometry.argtypes = [c_void_p, POINTER(c_int), POINTER(c_int)]
        lib.talfe_plan_num_frames.restype = c_int64
        lib.talfe_plan_num_frames.argtypes = [c_void_p, c_int64]
    if hasattr(lib, "talfe_resample"):
        lib.talfe_resample.restype = c_int
        lib.talfe_resample.argtypes = [c_void_p, c_int, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_int64, c_int64, c_void_p]
    lib.talfe_plan_destroy.restype = None
    lib.talfe_plan_destroy.argtypes = [c_void_p]
    lib.talfe_plan_n_mels.restype = c_int
    lib.talfe_plan_n_mels.argtypes = [c_void_p]
    lib.talfe_workspace_bytes.restype = c_size_t
    lib.talfe_workspace_bytes.argtypes = [c_void_p, c_int64, c_int64]
    lib.talfe_run.restype = c_int
    lib.talfe_run.argtypes = [c_void_p, POINTER(Job)]
    lib.talfe_logmel_forward.restype = c_int
    lib.talfe_logmel_forward.argtypes = [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_float,
                                         c_void_p, c_size_t, c_void_p]
    lib.talfe_apply_stats.restype = c_int
    lib.talfe_apply_stats.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p,
                                      c_void_p, c_void_p]
    lib.talfe_stream_staging_bytes.restype = c_size_t
    lib.talfe_stream_staging_bytes.argtypes = [c_int, c_int64]
    lib.talfe_stream_episode.restype = c_int
    lib.talfe_stream_episode.argtypes = [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p,
                                         c_float, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]
    lib.talfe_allreduce_stats.restype = c_int
    lib.talfe_allreduce_stats.argtypes = [c_void_p, c_int64, c_void_p, c_void_p]
    lib.talfe_detect_padding.restype = c_int
    lib.talfe_detect_padding.argtypes = [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p]
    lib.talfe_synth_fill.restype = c_int
    lib.talfe_synth_fill.argtypes = [c_void_p, c_int, c_int64, c_int64, c_int64, c_uint64, c_int64, c_int64, c_void_p]
    _LIB = lib
    return lib


def check(status: int, what: str = "talfe") -> None:
    if status == 0:
        return
    lib = load()
    msg = lib.talfe_strerror(status).decode()
    if status == -5:
        msg += f" (cudaError {lib.talfe_last_cuda_error()})"
    if status == ERR_TOO_SHORT:
        raise RuntimeError(f"{what}: {msg}")          # same exception type the reference raises from F.pad
    if status == -1:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg}")
