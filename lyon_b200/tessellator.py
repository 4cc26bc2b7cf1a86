"""Host-side mirror of lyon's stroke/flatten entry points over the CUDA C ABI.

  StrokeTessellator.tessellate_path(path, options)      <- stroke.rs:164-177
  tessellate_paths(paths | PathBatch, options)          <- the north-star batch API:
      result == for p in paths { StrokeTessellator::new().tessellate_path(p, opts, &mut BuffersBuilder::new(&mut out, ctor)) }
  flatten_cubics / flatten_quadratics                   <- CubicBezierSegment::flattened (cubic_bezier.rs:563-614) batched

`DeviceTessellator` is the HBM-resident flavour (torch tensors in, torch tensors out) used by bench.py and
by multi-GPU sharding; the plain functions move host numpy arrays across PCIe around it.
PyTorch is only used for device memory and streams."""
from __future__ import annotations

import ctypes
from typing import Iterable, Optional, Sequence, Union

import numpy as np

from . import _ffi
from .geometry import STROKE_VERTEX_DTYPE, TessellationError, TooManyVertices, VertexBuffers
from .options import StrokeOptions
from .path import Path, PathBatch


def _raise(L, ctx, code: int):
    msg = L.lyon_cuda_last_error(ctx)
    msg = msg.decode() if msg else ""
    names = {-1: "UnsupportedParameter", -2: "TooManyVertices", -3: "Internal", -4: "CudaError",
             -5: "InvalidArgument", -6: "OutOfMemory", -7: "BufferTooSmall"}
    what = f"{names.get(code, 'Error')}: {msg}" if msg else names.get(code, "Error")
    if code == _ffi.ERR_TOO_MANY_VERTICES:
        raise TooManyVertices(code, what)
    raise TessellationError(code, what)


class DeviceBatch:
    """A PathBatch resident in HBM (torch uint8/float32/int32 tensors)."""

    def __init__(self, batch: PathBatch, device, pinned: bool = False):
        import torch

        self.n_paths = batch.n_paths
        self.num_attributes = batch.num_attributes
        self.n_verbs = int(batch.verbs.shape[0])
        self.n_points = int(batch.points.shape[0])

        def up(a, dt):
            t = torch.from_numpy(np.ascontiguousarray(a).view(dt))
            if pinned:
                t = t.pin_memory()
            return t.to(device, non_blocking=pinned)

        self.verbs = up(batch.verbs, np.uint8)
        self.points = up(batch.points.reshape(-1), np.float32)
        self.path_verb_begin = up(batch.path_verb_begin, np.int32)
        self.path_point_begin = up(batch.path_point_begin, np.int32)

    def c_struct(self) -> _ffi.CPathBatch:
        return _ffi.CPathBatch(self.verbs.data_ptr(), self.points.data_ptr(), self.path_verb_begin.data_ptr(),
                               self.path_point_begin.data_ptr(), self.n_paths, self.num_attributes, self.n_verbs,
                               self.n_points)

    def input_bytes(self) -> int:
        return self.n_verbs + 8 * self.n_points + 8 * (self.n_paths + 1)


class DeviceTessellator:
    """One CUDA context (device + stream + grow-only workspace).  Not thread-safe, like `&mut StrokeTessellator`."""

    def __init__(self, device: Union[int, str, None] = None, stream=None):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("lyon_b200 needs a CUDA device; there is no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device if isinstance(device, int) else torch.device(device).index or 0))
        self.L = _ffi.lib()
        self.ctx = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            # launch on torch's current stream (or the given torch stream) so torch ops and events order with ours
            self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)
            st = self.L.lyon_cuda_create(self.device.index, ctypes.c_void_p(self.stream.cuda_stream),
                                         ctypes.byref(self.ctx))
        if st != 0:
            raise TessellationError(st, "lyon_cuda_create failed")

    def close(self):
        if getattr(self, "ctx", None) and self.ctx.value:
            self.L.lyon_cuda_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- stroke ---------------------------------------------------------------------------------------
    def upload(self, batch: PathBatch, pinned: bool = False) -> DeviceBatch:
        return DeviceBatch(batch, self.device, pinned)

    def stroke_plan(self, dbatch: DeviceBatch, options: StrokeOptions):
        """Count passes + scans. Returns (n_vertices, n_indices)."""
        nv, ni = ctypes.c_uint64(), ctypes.c_uint64()
        cb, co = dbatch.c_struct(), options.to_c()
        st = self.L.lyon_cuda_stroke_plan(self.ctx, ctypes.byref(cb), ctypes.byref(co), ctypes.byref(nv), ctypes.byref(ni))
        if st != 0:
            _raise(self.L, self.ctx, st)
        self._planned = dbatch  # keep the device inputs alive until emit
        return nv.value, ni.value

    def stroke_emit(self, n_vertices: int, n_indices: int, vertex_base: int = 0, vertices=None, indices=None,
                    with_ranges: bool = False):
        """Emit pass into (given or freshly allocated) device tensors: vertices float32 [V, 8] (the 32-byte
        lyon_stroke_vertex rows), indices int32 [I] (u32 bit patterns)."""
        torch = self.torch
        if vertices is None:
            vertices = torch.empty((n_vertices, 8), dtype=torch.float32, device=self.device)
        if indices is None:
            indices = torch.empty((n_indices,), dtype=torch.int32, device=self.device)
        pve = pie = None
        if with_ranges:
            pve = torch.empty((self._planned.n_paths,), dtype=torch.int64, device=self.device)
            pie = torch.empty((self._planned.n_paths,), dtype=torch.int64, device=self.device)
        st = self.L.lyon_cuda_stroke_emit(self.ctx, vertices.data_ptr(), indices.data_ptr(), vertex_base,
                                          pve.data_ptr() if with_ranges else None, pie.data_ptr() if with_ranges else None)
        if st != 0:
            _raise(self.L, self.ctx, st)
        if with_ranges:
            return vertices, indices, pve, pie
        return vertices, indices

    def stroke_fused(self, dbatch: DeviceBatch, options: StrokeOptions, vertices=None, indices=None, vertex_base: int = 0,
                     with_ranges: bool = False):
        """Single-pass stroke (lyon_cuda_stroke_fused): count, scan and emit in one kernel, into caller-kept
        device buffers -- `vertices` float32 [capV, 8], `indices` int32 [capI].  Without buffers the tessellator
        keeps its own grow-only pair (the way bench/tess keeps one VertexBuffers across iterations); a call whose
        result does not fit grows them and runs once more.  Returns views (vertices[:nv], indices[:ni])."""
        torch = self.torch
        own = vertices is None or indices is None
        if own:
            vertices, indices = getattr(self, "_own_v", None), getattr(self, "_own_i", None)
            if vertices is None:
                vertices = self._own_v = torch.empty((0, 8), dtype=torch.float32, device=self.device)
                indices = self._own_i = torch.empty((0,), dtype=torch.int32, device=self.device)
        pve = pie = None
        if with_ranges:
            pve = torch.empty((dbatch.n_paths,), dtype=torch.int64, device=self.device)
            pie = torch.empty((dbatch.n_paths,), dtype=torch.int64, device=self.device)
        nv, ni = ctypes.c_uint64(), ctypes.c_uint64()
        cb, co = dbatch.c_struct(), options.to_c()
        for attempt in range(2):
            st = self.L.lyon_cuda_stroke_fused(self.ctx, ctypes.byref(cb), ctypes.byref(co), vertices.data_ptr(),
                                               vertices.shape[0], indices.data_ptr(), indices.shape[0], vertex_base,
                                               pve.data_ptr() if with_ranges else None,
                                               pie.data_ptr() if with_ranges else None, ctypes.byref(nv), ctypes.byref(ni))
            if st == _ffi.ERR_BUFFER_TOO_SMALL and own and attempt == 0:
                del vertices, indices
                self._own_v = self._own_i = None
                vertices = self._own_v = torch.empty((nv.value + nv.value // 64 + 64, 8), dtype=torch.float32, device=self.device)
                indices = self._own_i = torch.empty((ni.value + ni.value // 64 + 256,), dtype=torch.int32, device=self.device)
                continue
            break
        if st != 0:
            _raise(self.L, self.ctx, st)
        out = (vertices[:nv.value], indices[:ni.value])
        return out + (pve, pie) if with_ranges else out

    # ---- multi-GPU (lyon_cuda_comm_* / lyon_cuda_stroke_fused_sharded) ---------------------------------------------
    def comm_unique_id(self) -> bytes:
        """The 128-byte NCCL id rank 0 creates; the application hands it to every rank (any transport)."""
        buf = (ctypes.c_uint8 * 128)()
        st = self.L.lyon_cuda_comm_unique_id(buf)
        if st != 0:
            _raise(self.L, self.ctx, st)
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        """Collective: join the communicator of the job (one context per GPU)."""
        assert len(unique_id) == 128
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        with self.torch.cuda.device(self.device):
            st = self.L.lyon_cuda_comm_init(self.ctx, buf, int(rank), int(nranks))
        if st != 0:
            _raise(self.L, self.ctx, st)
        self.comm_rank, self.comm_size = int(rank), int(nranks)

    def comm_destroy(self):
        self.L.lyon_cuda_comm_destroy(self.ctx)

    def stroke_fused_sharded(self, dbatch: DeviceBatch, options: StrokeOptions, vertices=None, indices=None):
        """Collective single-pass stroke of this rank's slice: fused kernels with slice-local ids, the all-gather of
        {V_r, I_r} and the index rebase, all enqueued on the context's stream.  Returns (vertices[:nv], indices[:ni],
        vertex_base, totals[nranks, 2])."""
        torch = self.torch
        own = vertices is None or indices is None
        if own:
            vertices, indices = getattr(self, "_own_v", None), getattr(self, "_own_i", None)
            if vertices is None:
                vertices = self._own_v = torch.empty((0, 8), dtype=torch.float32, device=self.device)
                indices = self._own_i = torch.empty((0,), dtype=torch.int32, device=self.device)
        nv, ni, base = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        totals = np.zeros((self.comm_size, 2), dtype=np.uint64)
        cb, co = dbatch.c_struct(), options.to_c()
        for attempt in range(2):
            st = self.L.lyon_cuda_stroke_fused_sharded(
                self.ctx, ctypes.byref(cb), ctypes.byref(co), vertices.data_ptr(), vertices.shape[0], indices.data_ptr(),
                indices.shape[0], None, None, ctypes.byref(nv), ctypes.byref(ni), ctypes.byref(base),
                totals.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
            if st == _ffi.ERR_BUFFER_TOO_SMALL and attempt == 0:
                # some rank's buffers were too small: every rank calls again (the exchange is collective)
                if own and (nv.value > vertices.shape[0] or ni.value > indices.shape[0]):
                    del vertices, indices
                    self._own_v = self._own_i = None
                    vertices = self._own_v = torch.empty((nv.value + nv.value // 64 + 64, 8), dtype=torch.float32, device=self.device)
                    indices = self._own_i = torch.empty((ni.value + ni.value // 64 + 256,), dtype=torch.int32, device=self.device)
                continue
            break
        if st != 0:
            _raise(self.L, self.ctx, st)
        return vertices[:nv.value], indices[:ni.value], int(base.value), totals

    def rebase_indices(self, indices, n_vertices: int, vertex_base: int):
        """indices += vertex_base in place (lyon_cuda_rebase_indices): global index rebasing after a single-pass emit."""
        st = self.L.lyon_cuda_rebase_indices(self.ctx, indices.data_ptr(), int(indices.shape[0]), int(n_vertices), int(vertex_base))
        if st != 0:
            _raise(self.L, self.ctx, st)
        return indices

    def interpolate_attributes(self, dbatch: DeviceBatch, vertices, path_vertex_end):
        """StrokeVertex::interpolated_attributes (stroke.rs:2728-2747) of every vertex: float32 [V, num_attributes] on the
        device.  `vertices` / `path_vertex_end` as returned by stroke(..., with_ranges=True) for the same batch."""
        torch = self.torch
        out = torch.empty((vertices.shape[0], dbatch.num_attributes), dtype=torch.float32, device=self.device)
        cb = dbatch.c_struct()
        st = self.L.lyon_cuda_interpolate_attributes(self.ctx, ctypes.byref(cb), vertices.data_ptr(), vertices.shape[0],
                                                     path_vertex_end.data_ptr(), out.data_ptr())
        if st != 0:
            _raise(self.L, self.ctx, st)
        return out

    def set_pipeline(self, general):
        """False / 0: automatic (shared-memory tiles owned by a CTA when every path fits one, else the general
        pipeline); True / 1: always the general multi-kernel pipeline; 2: tiles owned by a single warp where they fit
        (an experiment kept for comparison, slower).  All produce identical output."""
        st = self.L.lyon_cuda_set_pipeline(self.ctx, int(general))
        if st != 0:
            _raise(self.L, self.ctx, st)

    def enable_timing(self, on: bool = True):
        self.L.lyon_cuda_enable_timing(self.ctx, 1 if on else 0)

    def last_kernel_ms(self) -> float:
        """CUDA-event duration of the last tile kernel launch (needs enable_timing)."""
        ms = ctypes.c_float()
        st = self.L.lyon_cuda_last_kernel_ms(self.ctx, ctypes.byref(ms))
        if st != 0:
            _raise(self.L, self.ctx, st)
        return float(ms.value)

    def check(self):
        st = self.L.lyon_cuda_check(self.ctx)
        if st != 0:
            _raise(self.L, self.ctx, st)

    def stroke(self, dbatch: DeviceBatch, options: StrokeOptions, vertex_base: int = 0, with_ranges: bool = False):
        nv, ni = self.stroke_plan(dbatch, options)
        return self.stroke_emit(nv, ni, vertex_base, with_ranges=with_ranges)

    # ---- flatten --------------------------------------------------------------------------------------
    def flatten_curves(self, d_ctrl, npts: int, tolerance: float, with_t: bool = False):
        """d_ctrl: float32 device tensor [n, 2*npts] (npts = 4 cubic / 3 quadratic).
        Returns (offsets int64 [n+1], points float32 [total, 2][, t float32 [total]]) on the device."""
        torch = self.torch
        n = int(d_ctrl.shape[0])
        offsets = torch.empty((n + 1,), dtype=torch.int64, device=self.device)
        total = ctypes.c_uint64()
        cnt = self.L.lyon_cuda_flatten_cubics_count if npts == 4 else self.L.lyon_cuda_flatten_quadratics_count
        emit = self.L.lyon_cuda_flatten_cubics_emit if npts == 4 else self.L.lyon_cuda_flatten_quadratics_emit
        st = cnt(self.ctx, d_ctrl.data_ptr(), n, tolerance, offsets.data_ptr(), ctypes.byref(total))
        if st != 0:
            _raise(self.L, self.ctx, st)
        pts = torch.empty((total.value, 2), dtype=torch.float32, device=self.device)
        t = torch.empty((total.value,), dtype=torch.float32, device=self.device) if with_t else None
        st = emit(self.ctx, d_ctrl.data_ptr(), n, tolerance, offsets.data_ptr(), total.value, pts.data_ptr(),
                  t.data_ptr() if with_t else None)
        if st != 0:
            _raise(self.L, self.ctx, st)
        return (offsets, pts, t) if with_t else (offsets, pts)

    def flatten_paths(self, dbatch: DeviceBatch, tolerance: float) -> DeviceBatch:
        """`path.iter().flattened(tolerance)` for every path of the batch (lyon_cuda_flatten_paths_*): a device batch
        of polylines in the same encoding."""
        torch = self.torch
        nv, npt = ctypes.c_uint64(), ctypes.c_uint64()
        cb = dbatch.c_struct()
        st = self.L.lyon_cuda_flatten_paths_count(self.ctx, ctypes.byref(cb), tolerance, ctypes.byref(nv), ctypes.byref(npt))
        if st != 0:
            _raise(self.L, self.ctx, st)
        out = DeviceBatch.__new__(DeviceBatch)
        out.n_paths, out.num_attributes = dbatch.n_paths, dbatch.num_attributes
        out.n_verbs, out.n_points = int(nv.value), int(npt.value)
        out.verbs = torch.empty((out.n_verbs,), dtype=torch.uint8, device=self.device)
        out.points = torch.empty((2 * out.n_points,), dtype=torch.float32, device=self.device)
        out.path_verb_begin = torch.empty((dbatch.n_paths + 1,), dtype=torch.int32, device=self.device)
        out.path_point_begin = torch.empty((dbatch.n_paths + 1,), dtype=torch.int32, device=self.device)
        st = self.L.lyon_cuda_flatten_paths_emit(self.ctx, out.verbs.data_ptr(), out.points.data_ptr(),
                                                 out.path_verb_begin.data_ptr(), out.path_point_begin.data_ptr())
        if st != 0:
            _raise(self.L, self.ctx, st)
        return out

    def measure_paths(self, dbatch: DeviceBatch, tolerance: float):
        """`PathMeasurements::from_path(path, tolerance)` for every path of the batch (lyon_cuda_measure_paths_*).
        Returns (distance f32[E], event_index i32[E], t f32[E], path_edge_begin i64[N + 1], path_length f32[N])."""
        torch = self.torch
        ne = ctypes.c_uint64()
        cb = dbatch.c_struct()
        st = self.L.lyon_cuda_measure_paths_count(self.ctx, ctypes.byref(cb), tolerance, ctypes.byref(ne))
        if st != 0:
            _raise(self.L, self.ctx, st)
        n = int(ne.value)
        dist = torch.empty((n,), dtype=torch.float32, device=self.device)
        idx = torch.empty((n,), dtype=torch.int32, device=self.device)
        t = torch.empty((n,), dtype=torch.float32, device=self.device)
        begin = torch.empty((dbatch.n_paths + 1,), dtype=torch.int64, device=self.device)
        length = torch.empty((dbatch.n_paths,), dtype=torch.float32, device=self.device)
        st = self.L.lyon_cuda_measure_paths_emit(self.ctx, dist.data_ptr(), idx.data_ptr(), t.data_ptr(), begin.data_ptr(),
                                                 length.data_ptr())
        if st != 0:
            _raise(self.L, self.ctx, st)
        return dist, idx, t, begin, length

    def flatten_curves_fused(self, d_ctrl, npts: int, tolerance: float, with_t: bool = False):
        """Single pass (count + scan + emit in one kernel) into buffers kept across calls; they grow when the batch
        needs more room (one extra pass).  Returns views of the kept buffers, like flatten_curves."""
        torch = self.torch
        n = int(d_ctrl.shape[0])
        fn = self.L.lyon_cuda_flatten_cubics_fused if npts == 4 else self.L.lyon_cuda_flatten_quadratics_fused
        if getattr(self, "_fl_off", None) is None or self._fl_off.shape[0] < n + 1:
            self._fl_off = torch.empty((n + 1,), dtype=torch.int64, device=self.device)
        if getattr(self, "_fl_pts", None) is None:
            self._fl_pts = torch.empty((0, 2), dtype=torch.float32, device=self.device)
            self._fl_t = torch.empty((0,), dtype=torch.float32, device=self.device)
        total = ctypes.c_uint64()
        for _ in range(2):
            if with_t and self._fl_t.shape[0] < self._fl_pts.shape[0]:
                self._fl_t = torch.empty((self._fl_pts.shape[0],), dtype=torch.float32, device=self.device)
            cap = int(self._fl_pts.shape[0])
            st = fn(self.ctx, d_ctrl.data_ptr(), n, tolerance, self._fl_off.data_ptr(), self._fl_pts.data_ptr() if cap else None,
                    cap, self._fl_t.data_ptr() if (with_t and cap) else None, ctypes.byref(total))
            if st != _ffi.ERR_BUFFER_TOO_SMALL:
                break
            self._fl_pts = None
            self._fl_pts = torch.empty((int(total.value * 1.02) + 16, 2), dtype=torch.float32, device=self.device)
        if st != 0:
            _raise(self.L, self.ctx, st)
        nt = int(total.value)
        out = (self._fl_off[: n + 1], self._fl_pts[:nt])
        return out + (self._fl_t[:nt],) if with_t else out

    def launch_count(self) -> int:
        return int(self.L.lyon_cuda_launch_count(self.ctx))

    def workspace_bytes(self) -> int:
        return int(self.L.lyon_cuda_workspace_bytes(self.ctx))

    def synchronize(self):
        self.L.lyon_cuda_synchronize(self.ctx)


_DEFAULT: Optional[DeviceTessellator] = None


def _default() -> DeviceTessellator:
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = DeviceTessellator()
    return _DEFAULT


def _as_batch(paths) -> PathBatch:
    if isinstance(paths, PathBatch):
        return paths
    if isinstance(paths, Path):
        return PathBatch.from_paths([paths])
    return PathBatch.from_paths(list(paths))


def tessellate_paths(paths: Union[PathBatch, Iterable[Path]], options: StrokeOptions,
                     tessellator: Optional[DeviceTessellator] = None) -> VertexBuffers:
    """`lyon_cuda::tessellate_paths(&[Path], &StrokeOptions) -> VertexBuffers<StrokeVertexGpu, u32>` with host
    buffers in and out, through the single C entry point lyon_cuda_tessellate_paths."""
    tz = tessellator or _default()
    batch = _as_batch(paths)
    L = tz.L
    verbs = np.ascontiguousarray(batch.verbs, np.uint8)
    pts = np.ascontiguousarray(batch.points, np.float32)
    vb = np.ascontiguousarray(batch.path_verb_begin, np.uint32)
    pb = np.ascontiguousarray(batch.path_point_begin, np.uint32)
    cb = _ffi.CPathBatch(verbs.ctypes.data, pts.ctypes.data, vb.ctypes.data, pb.ctypes.data, batch.n_paths,
                         batch.num_attributes, verbs.shape[0], pts.shape[0])
    co = options.to_c()
    pv, pi = ctypes.c_void_p(), ctypes.c_void_p()
    nv, ni = ctypes.c_uint64(), ctypes.c_uint64()
    with tz.torch.cuda.device(tz.device):
        st = L.lyon_cuda_tessellate_paths(tz.ctx, ctypes.byref(cb), ctypes.byref(co), ctypes.byref(pv), ctypes.byref(nv),
                                          ctypes.byref(pi), ctypes.byref(ni))
    if st != 0:
        _raise(L, tz.ctx, st)
    try:
        v = np.frombuffer(ctypes.string_at(pv.value, nv.value * 32), dtype=STROKE_VERTEX_DTYPE).copy() if nv.value else \
            np.zeros(0, STROKE_VERTEX_DTYPE)
        i = np.frombuffer(ctypes.string_at(pi.value, ni.value * 4), dtype=np.uint32).copy() if ni.value else \
            np.zeros(0, np.uint32)
    finally:
        if pv.value:
            L.lyon_cuda_free_host(pv)
        if pi.value:
            L.lyon_cuda_free_host(pi)
    return VertexBuffers(v, i)


def device_mesh_to_host(vertices, indices) -> VertexBuffers:
    """Device tensors of DeviceTessellator.stroke_emit -> host VertexBuffers."""
    v = vertices.cpu().numpy().view(STROKE_VERTEX_DTYPE).reshape(-1)
    i = indices.cpu().numpy().view(np.uint32)
    return VertexBuffers(v, i)


def flatten_cubics(ctrl: np.ndarray, tolerance: float, tessellator: Optional[DeviceTessellator] = None, with_t=False):
    """Batched `CubicBezierSegment::flattened(tolerance)`: ctrl [n, 8] f32 -> (offsets u64 [n+1], points f32 [total, 2])."""
    tz = tessellator or _default()
    d = tz.torch.from_numpy(np.ascontiguousarray(ctrl, np.float32).reshape(-1, 8)).to(tz.device)
    out = tz.flatten_curves(d, 4, tolerance, with_t)
    return tuple(x.cpu().numpy().view(np.uint64) if k == 0 else x.cpu().numpy() for k, x in enumerate(out))


def flatten_quadratics(ctrl: np.ndarray, tolerance: float, tessellator: Optional[DeviceTessellator] = None, with_t=False):
    tz = tessellator or _default()
    d = tz.torch.from_numpy(np.ascontiguousarray(ctrl, np.float32).reshape(-1, 6)).to(tz.device)
    out = tz.flatten_curves(d, 3, tolerance, with_t)
    return tuple(x.cpu().numpy().view(np.uint64) if k == 0 else x.cpu().numpy() for k, x in enumerate(out))


class _PinnedArray:
    """A numpy view of page-locked host memory from lyon_cuda_host_alloc (freed with the object)."""

    def __init__(self, L, nbytes: int):
        self.L = L
        self.ptr = ctypes.c_void_p()
        st = L.lyon_cuda_host_alloc(ctypes.byref(self.ptr), max(int(nbytes), 1))
        if st != _ffi.LYON_OK:
            raise MemoryError(f"lyon_cuda_host_alloc({nbytes}) failed ({st})")
        self.nbytes = max(int(nbytes), 1)
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * self.nbytes).from_address(self.ptr.value))

    def __del__(self):
        try:
            if self.ptr:
                self.L.lyon_cuda_host_free(self.ptr)
                self.ptr = ctypes.c_void_p()
        except Exception:
            pass


class HostPipeline:
    """End-to-end host->host tessellation of a large batch: the public call a lyon user makes
    (`tessellate_paths(&[Path], &StrokeOptions) -> VertexBuffers`), a thin wrapper over the C entry point
    `lyon_cuda_tessellate_paths_into` (include/lyon_cuda.h), which cuts the batch into chunks of paths and alternates two
    streams so that the PCIe download of chunk i overlaps the upload + kernels of chunk i + 1.  Chunk c is emitted with
    vertex_base = vertices of chunks < c, so the concatenated host buffers equal the one-shot result.

    Inputs are staged in pinned host memory once (`stage`), outputs land in pinned host arrays that are reused
    across calls (like bench/tess reuses its VertexBuffers)."""

    API = "lyon_cuda_tessellate_paths_into"

    def __init__(self, device=None, chunk_paths: int = 32768):
        import torch

        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.chunk_paths = int(chunk_paths)
        self.tz = DeviceTessellator(self.device.index)
        self.L = self.tz.L
        self.out_v = self.out_i = None
        self.cap_v = self.cap_i = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def stage(self, batch: PathBatch):
        """Copy the input batch into pinned host memory (outside any timed region)."""
        self.batch = batch
        nv, npt, np_ = int(batch.verbs.shape[0]), int(batch.points.shape[0]), batch.n_paths
        self._in = [_PinnedArray(self.L, nv), _PinnedArray(self.L, 8 * npt), _PinnedArray(self.L, 4 * (np_ + 1)),
                    _PinnedArray(self.L, 4 * (np_ + 1))]
        self._in[0].array[:nv] = np.ascontiguousarray(batch.verbs, np.uint8)
        self._in[1].array[:8 * npt] = np.ascontiguousarray(batch.points, np.float32).reshape(-1).view(np.uint8)
        self._in[2].array[:4 * (np_ + 1)] = np.ascontiguousarray(batch.path_verb_begin, np.uint32).view(np.uint8)
        self._in[3].array[:4 * (np_ + 1)] = np.ascontiguousarray(batch.path_point_begin, np.uint32).view(np.uint8)
        self.cbatch = _ffi.CPathBatch(self._in[0].ptr, self._in[1].ptr, self._in[2].ptr, self._in[3].ptr, np_,
                                      batch.num_attributes, nv, npt)
        self.h2d_bytes = nv + 8 * npt + 8 * (np_ + 1)

    def _alloc_outputs(self, cap_v: int, cap_i: int):
        self.out_v = self.out_i = None  # release the old pinned buffers first
        self.out_v = _PinnedArray(self.L, 32 * cap_v)
        self.out_i = _PinnedArray(self.L, 4 * cap_i)
        self.cap_v, self.cap_i = cap_v, cap_i

    def run(self, options: StrokeOptions):
        """One end-to-end pass. Returns (n_vertices, n_indices); results are in the pinned arrays (`result`)."""
        copts = options.to_c()
        nv, ni = ctypes.c_uint64(0), ctypes.c_uint64(0)
        for _attempt in range(2):
            st = self.L.lyon_cuda_tessellate_paths_into(
                self.tz.ctx, ctypes.byref(self.cbatch), ctypes.byref(copts),
                self.out_v.ptr if self.out_v else None, self.cap_v, self.out_i.ptr if self.out_i else None, self.cap_i,
                self.chunk_paths, ctypes.byref(nv), ctypes.byref(ni))
            if st != _ffi.ERR_BUFFER_TOO_SMALL:
                break
            self._alloc_outputs(int(nv.value) + 1024, int(ni.value) + 4096)  # exact sizes are known now (first call only)
        if st != _ffi.LYON_OK:
            _raise(self.L, self.tz.ctx, st)
        self.d2h_bytes = 32 * int(nv.value) + 4 * int(ni.value)
        return int(nv.value), int(ni.value)

    def result(self, nv: int, ni: int) -> VertexBuffers:
        return VertexBuffers(self.out_v.array[:32 * nv].view(STROKE_VERTEX_DTYPE).reshape(-1),
                             self.out_i.array[:4 * ni].view(np.uint32))

    def launch_count(self) -> int:
        return self.tz.launch_count()
