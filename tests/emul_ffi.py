"""ctypes binding of tests/emul/libemul.so: the CUDA per-item code compiled for the CPU (test-only)."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

from lyon_b200.geometry import STROKE_VERTEX_DTYPE, VertexBuffers
from lyon_b200.options import CStrokeOptions, StrokeOptions
from lyon_b200.path import PathBatch

ROOT = Path(__file__).resolve().parent.parent
_LIB = None


class CPathBatch(ctypes.Structure):
    _fields_ = [("verbs", ctypes.c_void_p), ("points", ctypes.c_void_p), ("path_verb_begin", ctypes.c_void_p),
                ("path_point_begin", ctypes.c_void_p), ("n_paths", ctypes.c_uint32), ("num_attributes", ctypes.c_uint32),
                ("n_verbs", ctypes.c_uint32), ("n_points", ctypes.c_uint32)]


def lib():
    global _LIB
    if _LIB is None:
        subprocess.run(["make", "-C", str(ROOT / "tests" / "emul"), "-s"], check=True)
        L = ctypes.CDLL(str(ROOT / "tests" / "emul" / "libemul.so"))
        L.emul_stroke_batch.restype = ctypes.c_int
        L.emul_stroke_batch.argtypes = [ctypes.POINTER(CPathBatch), ctypes.POINTER(CStrokeOptions), ctypes.c_uint64,
                                        ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64),
                                        ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64),
                                        ctypes.c_void_p, ctypes.c_void_p]
        L.emul_stroke_batch_tile.restype = ctypes.c_int
        L.emul_stroke_batch_tile.argtypes = L.emul_stroke_batch.argtypes
        L.emul_stroke_batch_wtile.restype = ctypes.c_int
        L.emul_stroke_batch_wtile.argtypes = L.emul_stroke_batch.argtypes
        L.emul_stroke_batch_vw.restype = ctypes.c_int
        L.emul_stroke_batch_vw.argtypes = L.emul_stroke_batch.argtypes
        L.emul_free.argtypes = [ctypes.c_void_p]
        L.emul_flatten.restype = ctypes.c_uint64
        L.emul_flatten.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p]
        _LIB = L
    return _LIB


class EmulError(Exception):
    def __init__(self, code):
        super().__init__(f"emul error {code}")
        self.code = code


PIPELINE = "general"  # or "tile": which device pipeline's per-item code stroke_batch emulates


def stroke_batch(batch: PathBatch, opts: StrokeOptions, vertex_base: int = 0, with_ranges=False, pipeline=None):
    L = lib()
    which = pipeline or PIPELINE
    fn = {"tile": L.emul_stroke_batch_tile, "wtile": L.emul_stroke_batch_wtile}.get(which, L.emul_stroke_batch)
    if opts.variable_line_width is not None and batch.num_attributes > 0:
        fn = L.emul_stroke_batch_vw  # variable width: the thread-per-path state machine, whatever the pipeline
    verbs = np.ascontiguousarray(batch.verbs, np.uint8)
    pts = np.ascontiguousarray(batch.points, np.float32)
    vb = np.ascontiguousarray(batch.path_verb_begin, np.uint32)
    pb = np.ascontiguousarray(batch.path_point_begin, np.uint32)
    cb = CPathBatch(verbs.ctypes.data, pts.ctypes.data, vb.ctypes.data, pb.ctypes.data, batch.n_paths,
                    batch.num_attributes, verbs.shape[0], pts.shape[0])
    co = opts.to_c()
    pv, pi = ctypes.c_void_p(), ctypes.c_void_p()
    nv, ni = ctypes.c_uint64(), ctypes.c_uint64()
    pve = np.zeros(batch.n_paths, np.uint64)
    pie = np.zeros(batch.n_paths, np.uint64)
    st = fn(ctypes.byref(cb), ctypes.byref(co), vertex_base, ctypes.byref(pv), ctypes.byref(nv),
                             ctypes.byref(pi), ctypes.byref(ni), pve.ctypes.data, pie.ctypes.data)
    if st != 0:
        if pv.value:
            L.emul_free(pv)
            L.emul_free(pi)
        raise EmulError(st)
    v = np.frombuffer(ctypes.string_at(pv.value, nv.value * 32), dtype=STROKE_VERTEX_DTYPE).copy()
    i = np.frombuffer(ctypes.string_at(pi.value, ni.value * 4), dtype=np.uint32).copy()
    L.emul_free(pv)
    L.emul_free(pi)
    if with_ranges:
        return VertexBuffers(v, i), pve, pie
    return VertexBuffers(v, i)


def flatten(ctrl: np.ndarray, npts: int, tol: float):
    L = lib()
    ctrl = np.ascontiguousarray(ctrl, np.float32).reshape(-1, 2 * npts)
    n = ctrl.shape[0]
    off = np.zeros(n + 1, np.uint64)
    total = L.emul_flatten(ctrl.ctypes.data, n, npts, tol, off.ctypes.data, None, None)
    pts = np.zeros((total, 2), np.float32)
    t = np.zeros(total, np.float32)
    L.emul_flatten(ctrl.ctypes.data, n, npts, tol, off.ctypes.data, pts.ctypes.data, t.ctypes.data)
    return off, pts, t
