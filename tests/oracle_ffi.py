"""ctypes binding of oracle/liblyon_oracle.so -- the CPU parity oracle (test infrastructure only)."""
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


def build():
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        so = ROOT / "oracle" / "liblyon_oracle.so"
        if not so.exists():
            build()
        L = ctypes.CDLL(str(so))
        L.oracle_stroke_batch.restype = ctypes.c_void_p
        L.oracle_stroke_batch.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint32, ctypes.c_uint32,
                                                                  ctypes.POINTER(CStrokeOptions), ctypes.c_uint64,
                                                                  ctypes.POINTER(ctypes.c_int)]
        L.oracle_mesh_num_vertices.restype = ctypes.c_uint64
        L.oracle_mesh_num_vertices.argtypes = [ctypes.c_void_p]
        L.oracle_mesh_num_indices.restype = ctypes.c_uint64
        L.oracle_mesh_num_indices.argtypes = [ctypes.c_void_p]
        L.oracle_mesh_copy.argtypes = [ctypes.c_void_p] * 5
        L.oracle_mesh_free.argtypes = [ctypes.c_void_p]
        L.oracle_mesh_interpolated_attributes.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_uint32, ctypes.c_void_p]
        L.oracle_stroke_batch_timed.restype = ctypes.c_double
        L.oracle_stroke_batch_timed.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint32, ctypes.c_uint32,
                                                                        ctypes.POINTER(CStrokeOptions), ctypes.c_int,
                                                                        ctypes.POINTER(ctypes.c_uint64),
                                                                        ctypes.POINTER(ctypes.c_uint64)]
        L.oracle_flatten_cubics.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_float] + [ctypes.c_void_p] * 3
        L.oracle_flatten_quadratics.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_float] + [ctypes.c_void_p] * 3
        L.oracle_flatten_cubics_timed.restype = ctypes.c_double
        L.oracle_flatten_cubics_timed.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_float, ctypes.c_int,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.oracle_flatten_paths.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint32, ctypes.c_float] + \
            [ctypes.c_void_p] * 4 + [ctypes.POINTER(ctypes.c_uint64)] * 2
        L.oracle_flatten_paths_attr.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float] + \
            [ctypes.c_void_p] * 4 + [ctypes.POINTER(ctypes.c_uint64)] * 2
        L.oracle_measure_paths.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float] + \
            [ctypes.c_void_p] * 5 + [ctypes.POINTER(ctypes.c_uint64)]
        L.oracle_compute_normal.argtypes = [ctypes.c_float] * 4 + [ctypes.c_void_p]
        L.oracle_fast_atan2.restype = ctypes.c_float
        L.oracle_fast_atan2.argtypes = [ctypes.c_float] * 2
        L.oracle_angle_to.restype = ctypes.c_float
        L.oracle_angle_to.argtypes = [ctypes.c_float] * 2
        L.oracle_cubic_wang.restype = ctypes.c_float
        L.oracle_cubic_wang.argtypes = [ctypes.c_void_p, ctypes.c_float]
        L.oracle_quad_wang.restype = ctypes.c_float
        L.oracle_quad_wang.argtypes = [ctypes.c_void_p, ctypes.c_float]
        L.oracle_find_sharp_turn.restype = ctypes.c_int
        L.oracle_find_sharp_turn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_circle_flattening_step.restype = ctypes.c_float
        L.oracle_circle_flattening_step.argtypes = [ctypes.c_float] * 2
        L.oracle_line_intersection.restype = ctypes.c_int
        L.oracle_line_intersection.argtypes = [ctypes.c_void_p] * 3
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleError(Exception):
    def __init__(self, code):
        super().__init__(f"oracle error {code}")
        self.code = code


def stroke_batch(batch: PathBatch, opts: StrokeOptions, max_index: int = 0xFFFFFFFE, with_ranges=False,
                 with_attributes=False):
    L = lib()
    c = opts.to_c()
    err = ctypes.c_int(0)
    verbs = np.ascontiguousarray(batch.verbs, np.uint8)
    pts = np.ascontiguousarray(batch.points, np.float32)
    vb = np.ascontiguousarray(batch.path_verb_begin, np.uint32)
    pb = np.ascontiguousarray(batch.path_point_begin, np.uint32)
    m = L.oracle_stroke_batch(_p(verbs), _p(pts), _p(vb), _p(pb), batch.n_paths, batch.num_attributes,
                              ctypes.byref(c), max_index, ctypes.byref(err))
    try:
        if err.value != 0:
            raise OracleError(err.value)
        nv, ni = L.oracle_mesh_num_vertices(m), L.oracle_mesh_num_indices(m)
        v = np.zeros(nv, STROKE_VERTEX_DTYPE)
        i = np.zeros(ni, np.uint32)
        pve = np.zeros(batch.n_paths, np.uint64)
        pie = np.zeros(batch.n_paths, np.uint64)
        L.oracle_mesh_copy(m, _p(v), _p(i), _p(pve), _p(pie))
        if with_attributes:  # StrokeVertex::interpolated_attributes of every vertex, [V, num_attributes]
            attrs = np.zeros((nv, batch.num_attributes), np.float32)
            L.oracle_mesh_interpolated_attributes(m, _p(pts), _p(pb), batch.num_attributes, _p(attrs))
    finally:
        L.oracle_mesh_free(m)
    if with_attributes:
        return (VertexBuffers(v, i), pve, pie, attrs) if with_ranges else (VertexBuffers(v, i), attrs)
    if with_ranges:
        return VertexBuffers(v, i), pve, pie
    return VertexBuffers(v, i)


def stroke_path(path, opts: StrokeOptions, **kw):
    return stroke_batch(PathBatch.from_paths([path]), opts, **kw)


def stroke_batch_timed(batch: PathBatch, opts: StrokeOptions, n_threads: int):
    L = lib()
    c = opts.to_c()
    tv, ti = ctypes.c_uint64(0), ctypes.c_uint64(0)
    verbs = np.ascontiguousarray(batch.verbs, np.uint8)
    pts = np.ascontiguousarray(batch.points, np.float32)
    vb = np.ascontiguousarray(batch.path_verb_begin, np.uint32)
    pb = np.ascontiguousarray(batch.path_point_begin, np.uint32)
    secs = L.oracle_stroke_batch_timed(_p(verbs), _p(pts), _p(vb), _p(pb), batch.n_paths, batch.num_attributes,
                                       ctypes.byref(c), n_threads, ctypes.byref(tv), ctypes.byref(ti))
    return secs, tv.value, ti.value


def flatten_cubics(ctrl: np.ndarray, tol: float, with_t=False):
    L = lib()
    ctrl = np.ascontiguousarray(ctrl, np.float32).reshape(-1, 8)
    n = ctrl.shape[0]
    off = np.zeros(n + 1, np.uint64)
    L.oracle_flatten_cubics(_p(ctrl), n, tol, _p(off), None, None)
    pts = np.zeros((int(off[-1]), 2), np.float32)
    t = np.zeros(int(off[-1]), np.float32)
    L.oracle_flatten_cubics(_p(ctrl), n, tol, _p(off), _p(pts), _p(t))
    return (off, pts, t) if with_t else (off, pts)


def flatten_quadratics(ctrl: np.ndarray, tol: float, with_t=False):
    L = lib()
    ctrl = np.ascontiguousarray(ctrl, np.float32).reshape(-1, 6)
    n = ctrl.shape[0]
    off = np.zeros(n + 1, np.uint64)
    L.oracle_flatten_quadratics(_p(ctrl), n, tol, _p(off), None, None)
    pts = np.zeros((int(off[-1]), 2), np.float32)
    t = np.zeros(int(off[-1]), np.float32)
    L.oracle_flatten_quadratics(_p(ctrl), n, tol, _p(off), _p(pts), _p(t))
    return (off, pts, t) if with_t else (off, pts)


def flatten_cubics_timed(ctrl: np.ndarray, tol: float, n_threads: int):
    L = lib()
    ctrl = np.ascontiguousarray(ctrl, np.float32).reshape(-1, 8)
    n = ctrl.shape[0]
    off = np.zeros(n + 1, np.uint64)
    L.oracle_flatten_cubics(_p(ctrl), n, tol, _p(off), None, None)
    pts = np.zeros((int(off[-1]), 2), np.float32)
    tot = ctypes.c_uint64(0)
    secs = L.oracle_flatten_cubics_timed(_p(ctrl), n, tol, n_threads, _p(pts), _p(off), ctypes.byref(tot))
    return secs, tot.value


def flatten_paths(batch: PathBatch, tol: float) -> PathBatch:
    """Path-level flattening (iterator.rs:125-226; with custom attributes: private.rs:56-122)."""
    L = lib()
    verbs = np.ascontiguousarray(batch.verbs, np.uint8)
    pts = np.ascontiguousarray(batch.points, np.float32)
    vb = np.ascontiguousarray(batch.path_verb_begin, np.uint32)
    pb = np.ascontiguousarray(batch.path_point_begin, np.uint32)
    na = int(batch.num_attributes)
    nv, npnt = ctypes.c_uint64(0), ctypes.c_uint64(0)
    L.oracle_flatten_paths_attr(_p(verbs), _p(pts), _p(vb), _p(pb), batch.n_paths, na, ctypes.c_float(tol), None, None, None, None,
                                ctypes.byref(nv), ctypes.byref(npnt))
    ov = np.zeros(nv.value, np.uint8)
    op = np.zeros((npnt.value, 2), np.float32)
    ovb = np.zeros(batch.n_paths + 1, np.uint32)
    opb = np.zeros(batch.n_paths + 1, np.uint32)
    L.oracle_flatten_paths_attr(_p(verbs), _p(pts), _p(vb), _p(pb), batch.n_paths, na, ctypes.c_float(tol), _p(ov), _p(op), _p(ovb),
                                _p(opb), ctypes.byref(nv), ctypes.byref(npnt))
    return PathBatch(ov, op, ovb, opb, na)


def measure_paths(batch: PathBatch, tol: float):
    """PathMeasurements::from_path for every path (measure.rs:140-260): (distance, event_index, t, path_edge_begin,
    path_length)."""
    L = lib()
    verbs = np.ascontiguousarray(batch.verbs, np.uint8)
    pts = np.ascontiguousarray(batch.points, np.float32)
    vb = np.ascontiguousarray(batch.path_verb_begin, np.uint32)
    pb = np.ascontiguousarray(batch.path_point_begin, np.uint32)
    ne = ctypes.c_uint64(0)
    args = (_p(verbs), _p(pts), _p(vb), _p(pb), batch.n_paths, int(batch.num_attributes), ctypes.c_float(tol))
    L.oracle_measure_paths(*args, None, None, None, None, None, ctypes.byref(ne))
    dist = np.zeros(ne.value, np.float32)
    idx = np.zeros(ne.value, np.uint32)
    t = np.zeros(ne.value, np.float32)
    begin = np.zeros(batch.n_paths + 1, np.uint64)
    length = np.zeros(batch.n_paths, np.float32)
    L.oracle_measure_paths(*args, _p(dist), _p(idx), _p(t), _p(begin), _p(length), ctypes.byref(ne))
    return dist, idx, t, begin, length


def compute_normal(v1, v2):
    out = np.zeros(2, np.float32)
    lib().oracle_compute_normal(v1[0], v1[1], v2[0], v2[1], _p(out))
    return out
