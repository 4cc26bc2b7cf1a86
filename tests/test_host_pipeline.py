"""-m gpu: the host-buffer entry point `lyon_cuda_tessellate_paths_into` (the call the e2e number of bench.py goes through)
against the oracle, and the two-phase entry points' plan invalidation."""
import ctypes

import numpy as np
import pytest

import oracle_ffi as O
import workloads as W
from lyon_b200 import LineCap, LineJoin, PathBatch, StrokeOptions, _ffi
from parity import assert_mesh_parity, bbox_extent

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tz():
    import torch
    from lyon_b200 import DeviceTessellator
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    t = DeviceTessellator(0)
    yield t
    t.close()


def test_host_pipeline_matches_oracle_across_ragged_chunks():
    """Seven chunks of 157 paths and a ragged last one; paths of different lengths, closed and open, so every chunk boundary
    falls at a different verb / point / vertex offset.  The concatenated pinned buffers must be the oracle's mesh: the
    vertex_base of chunk c is the vertex count of the chunks before it (geometry_builder.rs:325-349)."""
    from lyon_b200 import HostPipeline
    parts = [W.c3_mixed_paths(400, 64, seed=11, closed_fraction=0.3), W.random_small_paths(500, seed=5),
             W.c3_mixed_paths(231, 17, seed=12)]
    paths = [b.path(i) for b in parts for i in range(b.n_paths)]
    batch = PathBatch.from_paths(paths)
    o = StrokeOptions(line_width=2.0, line_join=LineJoin.Round, tolerance=0.1).with_line_cap(LineCap.Round)
    want = O.stroke_batch(batch, o)
    pipe = HostPipeline(0, chunk_paths=157)
    assert pipe.API == "lyon_cuda_tessellate_paths_into"
    pipe.stage(batch)
    for _ in range(2):  # the second run reuses the pinned outputs and the slots' device buffers
        nv, ni = pipe.run(o)
        got = pipe.result(nv, ni)
        assert_mesh_parity(got, want, bbox_extent(batch), what="host pipeline")
    assert pipe.d2h_bytes == 32 * len(want.vertices) + 4 * len(want.indices)
    # another option set through the same pipeline object (different sizes: the outputs grow or are reused)
    o2 = StrokeOptions(line_width=1.0, line_join=LineJoin.MiterClip, tolerance=0.01)
    nv, ni = pipe.run(o2)
    assert_mesh_parity(pipe.result(nv, ni), O.stroke_batch(batch, o2), bbox_extent(batch), what="host pipeline, miter clip")


def test_host_pipeline_reports_exact_sizes_when_too_small(tz):
    batch = W.c3_mixed_paths(300, 32, seed=3)
    o = StrokeOptions(line_width=2.0, line_join=LineJoin.Round).with_line_cap(LineCap.Round)
    want = O.stroke_batch(batch, o)
    L = tz.L
    verbs = np.ascontiguousarray(batch.verbs, np.uint8)
    pts = np.ascontiguousarray(batch.points, np.float32)
    vb = np.ascontiguousarray(batch.path_verb_begin, np.uint32)
    pb = np.ascontiguousarray(batch.path_point_begin, np.uint32)
    cb = _ffi.CPathBatch(verbs.ctypes.data, pts.ctypes.data, vb.ctypes.data, pb.ctypes.data, batch.n_paths, 0,
                         verbs.shape[0], pts.shape[0])
    co = o.to_c()
    nv, ni = ctypes.c_uint64(), ctypes.c_uint64()
    st = L.lyon_cuda_tessellate_paths_into(tz.ctx, ctypes.byref(cb), ctypes.byref(co), None, 0, None, 0, 64,
                                           ctypes.byref(nv), ctypes.byref(ni))
    assert st == _ffi.ERR_BUFFER_TOO_SMALL
    assert (nv.value, ni.value) == (len(want.vertices), len(want.indices))
    # pageable (plain numpy) output buffers of the exact size work too
    hv = np.zeros(nv.value, dtype=want.vertices.dtype)
    hi = np.zeros(ni.value, dtype=np.uint32)
    st = L.lyon_cuda_tessellate_paths_into(tz.ctx, ctypes.byref(cb), ctypes.byref(co), hv.ctypes.data, nv.value, hi.ctypes.data,
                                           ni.value, 64, ctypes.byref(nv), ctypes.byref(ni))
    assert st == _ffi.LYON_OK
    assert np.array_equal(hi, want.indices)
    assert hv.tobytes() == want.vertices.tobytes()


def test_stale_plans_are_rejected(tz):
    """A plan lives in workspace the other pipelines reuse: an emit after another call that may have overwritten it must
    fail with INVALID_ARGUMENT instead of reading foreign data (plan -> flatten -> emit; count -> plan -> emit)."""
    import torch
    L = tz.L
    batch = W.random_small_paths(50, seed=9)
    o = StrokeOptions(line_width=1.0)
    d = tz.upload(batch)
    cb, co = d.c_struct(), o.to_c()
    nv, ni = ctypes.c_uint64(), ctypes.c_uint64()
    ctrl = torch.from_numpy(W.c2_cubics(100)).cuda()
    offs = torch.empty(101, dtype=torch.int64, device="cuda")
    total = ctypes.c_uint64()
    for pipeline in (0, 1):  # tile and general
        tz.set_pipeline(pipeline)
        try:
            assert L.lyon_cuda_stroke_plan(tz.ctx, ctypes.byref(cb), ctypes.byref(co), ctypes.byref(nv), ctypes.byref(ni)) == 0
            assert L.lyon_cuda_flatten_cubics_count(tz.ctx, ctrl.data_ptr(), 100, 0.01, offs.data_ptr(), ctypes.byref(total)) == 0
            v = torch.empty((max(nv.value, 1), 8), dtype=torch.float32, device="cuda")
            i = torch.empty(max(ni.value, 1), dtype=torch.int32, device="cuda")
            assert L.lyon_cuda_stroke_emit(tz.ctx, v.data_ptr(), i.data_ptr(), 0, None, None) == _ffi.ERR_INVALID_ARGUMENT
        finally:
            tz.set_pipeline(False)
    no_v, no_p = ctypes.c_uint64(), ctypes.c_uint64()
    assert L.lyon_cuda_flatten_paths_count(tz.ctx, ctypes.byref(cb), 0.05, ctypes.byref(no_v), ctypes.byref(no_p)) == 0
    assert L.lyon_cuda_stroke_plan(tz.ctx, ctypes.byref(cb), ctypes.byref(co), ctypes.byref(nv), ctypes.byref(ni)) == 0
    ov = torch.empty(max(no_v.value, 1), dtype=torch.uint8, device="cuda")
    op = torch.empty((max(no_p.value, 1), 2), dtype=torch.float32, device="cuda")
    r1 = torch.empty(batch.n_paths + 1, dtype=torch.int32, device="cuda")
    r2 = torch.empty(batch.n_paths + 1, dtype=torch.int32, device="cuda")
    assert L.lyon_cuda_flatten_paths_emit(tz.ctx, ov.data_ptr(), op.data_ptr(), r1.data_ptr(), r2.data_ptr()) == _ffi.ERR_INVALID_ARGUMENT
    # and the plain sequence still works afterwards
    got = tz.stroke(d, o)
    torch.cuda.synchronize()
    tz.check()
    from lyon_b200 import device_mesh_to_host
    assert np.array_equal(device_mesh_to_host(*got).indices, O.stroke_batch(batch, o).indices)


def test_sharded_entry_single_rank(tz):
    """lyon_cuda_comm_init / lyon_cuda_stroke_fused_sharded with a communicator of one rank: the exchange degenerates (base 0)
    but every step runs -- NCCL communicator, all-gather on the context's stream, rebase kernel, totals read-back."""
    import torch
    from lyon_b200 import device_mesh_to_host
    batch = W.c3_mixed_paths(500, 24, seed=21, closed_fraction=0.2)
    o = StrokeOptions(line_width=2.0, line_join=LineJoin.Round).with_line_cap(LineCap.Round)
    want = O.stroke_batch(batch, o)
    tz.comm_init(tz.comm_unique_id(), 0, 1)
    try:
        v, i, base, totals = tz.stroke_fused_sharded(tz.upload(batch), o)
        torch.cuda.synchronize()
        tz.check()
        assert base == 0 and totals.tolist() == [[len(want.vertices), len(want.indices)]]
        assert_mesh_parity(device_mesh_to_host(v, i), want, bbox_extent(batch), what="sharded, one rank")
    finally:
        tz.comm_destroy()


def test_sharded_entry_two_ranks():
    """Two processes, two GPUs (skipped on a one-GPU box): every rank strokes its slice through the C entry point, the
    concatenation of the per-rank buffers must be the oracle's mesh of the whole batch."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = O.ROOT / "tools" / "sharded_check.py"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", str(script)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sharded_check: OK" in r.stdout
