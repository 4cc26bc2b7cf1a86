"""ctypes binding of lyon_b200/liblyon_cuda.so (the C ABI of include/lyon_cuda.h).

The product path fails loudly when the CUDA library is missing: there is no CPU fallback."""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

from .options import CStrokeOptions

# LYON_CUDA_LIB points at an alternative build of the same library (instrumented builds of tools/dev_build.sh)
_LIB_PATH = Path(os.environ.get("LYON_CUDA_LIB") or Path(__file__).resolve().parent / "liblyon_cuda.so")
_LIB = None

LYON_OK = 0
ERR_UNSUPPORTED_PARAMETER = -1
ERR_TOO_MANY_VERTICES = -2
ERR_INTERNAL = -3
ERR_CUDA = -4
ERR_INVALID_ARGUMENT = -5
ERR_OUT_OF_MEMORY = -6
ERR_BUFFER_TOO_SMALL = -7

# every symbol include/lyon_cuda.h declares (tests check the library exports all of them)
SYMBOLS = [
    "lyon_cuda_create", "lyon_cuda_destroy", "lyon_cuda_set_stream", "lyon_cuda_synchronize", "lyon_cuda_last_error",
    "lyon_cuda_launch_count", "lyon_cuda_workspace_bytes",
    "lyon_cuda_flatten_cubics_count", "lyon_cuda_flatten_cubics_emit",
    "lyon_cuda_flatten_quadratics_count", "lyon_cuda_flatten_quadratics_emit",
    "lyon_cuda_flatten_cubics_fused", "lyon_cuda_flatten_quadratics_fused",
    "lyon_cuda_flatten_paths_count", "lyon_cuda_flatten_paths_emit",
    "lyon_cuda_measure_paths_count", "lyon_cuda_measure_paths_emit",
    "lyon_cuda_stroke_plan", "lyon_cuda_stroke_emit", "lyon_cuda_check", "lyon_cuda_stroke_fused", "lyon_cuda_rebase_indices", "lyon_cuda_set_pipeline",
    "lyon_cuda_interpolate_attributes",
    "lyon_cuda_enable_timing", "lyon_cuda_last_kernel_ms",
    "lyon_cuda_tessellate_paths", "lyon_cuda_free_host",
    "lyon_cuda_tessellate_paths_into", "lyon_cuda_host_alloc", "lyon_cuda_host_free",
    "lyon_cuda_comm_unique_id", "lyon_cuda_comm_init", "lyon_cuda_comm_destroy", "lyon_cuda_stroke_fused_sharded",
]


class CPathBatch(ctypes.Structure):
    _fields_ = [
        ("verbs", ctypes.c_void_p),
        ("points", ctypes.c_void_p),
        ("path_verb_begin", ctypes.c_void_p),
        ("path_point_begin", ctypes.c_void_p),
        ("n_paths", ctypes.c_uint32),
        ("num_attributes", ctypes.c_uint32),
        ("n_verbs", ctypes.c_uint32),
        ("n_points", ctypes.c_uint32),
    ]


def lib_path() -> Path:
    return _LIB_PATH


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "lyon_b200 has no CPU fallback."
        )
    L = ctypes.CDLL(str(_LIB_PATH))
    vp, u64, u32, f32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_float
    pu64 = ctypes.POINTER(ctypes.c_uint64)
    L.lyon_cuda_create.argtypes = [ctypes.c_int, vp, ctypes.POINTER(vp)]
    L.lyon_cuda_destroy.argtypes = [vp]
    L.lyon_cuda_destroy.restype = None
    L.lyon_cuda_set_stream.argtypes = [vp, vp]
    L.lyon_cuda_synchronize.argtypes = [vp]
    L.lyon_cuda_last_error.argtypes = [vp]
    L.lyon_cuda_last_error.restype = ctypes.c_char_p
    L.lyon_cuda_launch_count.argtypes = [vp]
    L.lyon_cuda_launch_count.restype = u64
    L.lyon_cuda_workspace_bytes.argtypes = [vp]
    L.lyon_cuda_workspace_bytes.restype = u64
    for name in ("lyon_cuda_flatten_cubics_count", "lyon_cuda_flatten_quadratics_count"):
        getattr(L, name).argtypes = [vp, vp, u64, f32, vp, pu64]
    for name in ("lyon_cuda_flatten_cubics_emit", "lyon_cuda_flatten_quadratics_emit"):
        getattr(L, name).argtypes = [vp, vp, u64, f32, vp, u64, vp, vp]
    for name in ("lyon_cuda_flatten_cubics_fused", "lyon_cuda_flatten_quadratics_fused"):
        getattr(L, name).argtypes = [vp, vp, u64, f32, vp, vp, u64, vp, pu64]
    L.lyon_cuda_flatten_paths_count.argtypes = [vp, ctypes.POINTER(CPathBatch), f32, pu64, pu64]
    L.lyon_cuda_flatten_paths_emit.argtypes = [vp, vp, vp, vp, vp]
    L.lyon_cuda_measure_paths_count.argtypes = [vp, ctypes.POINTER(CPathBatch), f32, pu64]
    L.lyon_cuda_measure_paths_emit.argtypes = [vp, vp, vp, vp, vp, vp]
    L.lyon_cuda_stroke_plan.argtypes = [vp, ctypes.POINTER(CPathBatch), ctypes.POINTER(CStrokeOptions), pu64, pu64]
    L.lyon_cuda_stroke_emit.argtypes = [vp, vp, vp, u64, vp, vp]
    L.lyon_cuda_check.argtypes = [vp]
    L.lyon_cuda_stroke_fused.argtypes = [vp, ctypes.POINTER(CPathBatch), ctypes.POINTER(CStrokeOptions), vp, u64, vp, u64, u64,
                                         vp, vp, pu64, pu64]
    L.lyon_cuda_rebase_indices.argtypes = [vp, vp, u64, u64, u64]
    L.lyon_cuda_set_pipeline.argtypes = [vp, ctypes.c_int]
    L.lyon_cuda_interpolate_attributes.argtypes = [vp, ctypes.POINTER(CPathBatch), vp, u64, vp, vp]
    L.lyon_cuda_enable_timing.argtypes = [vp, ctypes.c_int]
    L.lyon_cuda_last_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    L.lyon_cuda_tessellate_paths.argtypes = [vp, ctypes.POINTER(CPathBatch), ctypes.POINTER(CStrokeOptions),
                                             ctypes.POINTER(vp), pu64, ctypes.POINTER(vp), pu64]
    L.lyon_cuda_free_host.argtypes = [vp]
    L.lyon_cuda_free_host.restype = None
    L.lyon_cuda_tessellate_paths_into.argtypes = [vp, ctypes.POINTER(CPathBatch), ctypes.POINTER(CStrokeOptions), vp, u64, vp, u64,
                                                  u32, pu64, pu64]
    L.lyon_cuda_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
    L.lyon_cuda_host_free.argtypes = [vp]
    L.lyon_cuda_host_free.restype = None
    L.lyon_cuda_comm_unique_id.argtypes = [vp]
    L.lyon_cuda_comm_init.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int]
    L.lyon_cuda_comm_destroy.argtypes = [vp]
    L.lyon_cuda_stroke_fused_sharded.argtypes = [vp, ctypes.POINTER(CPathBatch), ctypes.POINTER(CStrokeOptions), vp, u64, vp, u64,
                                                 vp, vp, pu64, pu64, pu64, pu64]
    _LIB = L
    return L
