"""-m gpu: the real CUDA path (through the C ABI) against the oracle on the same seeded inputs, the committed
fixtures, edge cases, and -- at BASELINE.json's full sizes -- size-independent properties."""
import numpy as np
import pytest

import oracle_ffi as O
import workloads as W
from lyon_b200 import LineCap, LineJoin, Path, PathBatch, StrokeOptions
from parity import assert_mesh_parity, bbox_extent

pytestmark = pytest.mark.gpu
GOLDEN = O.ROOT / "tests" / "golden"


@pytest.fixture(scope="module")
def tz():
    import torch
    from lyon_b200 import DeviceTessellator
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    t = DeviceTessellator(0)
    yield t
    t.close()


ROUTE = "tile"


@pytest.fixture(autouse=True, params=["tile", "fused", "wtile", "wfused", "general"])
def route(request):
    """Every stroke test runs five ways: plan+emit and the single-pass fused call through the automatic choice
    (shared-memory tiles owned by a CTA when the paths fit), the same two through the warp-owned tile kernel (kept as a
    measured experiment), and plan+emit through the general multi-kernel pipeline (any path length)."""
    global ROUTE
    ROUTE = request.param
    yield request.param
    ROUTE = "tile"


def gpu_stroke(tz, batch, o, vertex_base=0, with_ranges=False):
    import torch
    from lyon_b200 import device_mesh_to_host
    d = tz.upload(batch)
    tz.set_pipeline(1 if ROUTE == "general" else (2 if ROUTE in ("wtile", "wfused") else 0))
    try:
        if ROUTE in ("fused", "wfused"):
            out = tz.stroke_fused(d, o, vertex_base=vertex_base, with_ranges=with_ranges)
        else:
            out = tz.stroke(d, o, vertex_base=vertex_base, with_ranges=with_ranges)
    finally:
        tz.set_pipeline(False)
    torch.cuda.synchronize()
    tz.check()
    if with_ranges:
        return device_mesh_to_host(out[0], out[1]), out[2].cpu().numpy().astype(np.uint64), out[3].cpu().numpy().astype(np.uint64)
    return device_mesh_to_host(out[0], out[1])


def all_options(tol=0.1, width=1.0):
    for join in LineJoin:
        for cap in LineCap:
            yield StrokeOptions(tolerance=tol, line_width=width, line_join=join).with_line_cap(cap)


def test_native_library_is_loaded(tz):
    from lyon_b200 import _ffi
    maps = open("/proc/self/maps").read()
    assert str(_ffi.lib_path()) in maps


@pytest.mark.parametrize("tol", [0.05, 0.1, 0.01, 0.001])
def test_c1_fixtures(tz, tol):
    """C1: assets/logo.path (width 1, tol 0.05, Miter/Butt) and the Rust logo (bench/tess stroke_01..03)."""
    for path in W.fixture_paths(GOLDEN):
        batch = PathBatch.from_paths([path])
        for o in all_options(tol):
            assert_mesh_parity(gpu_stroke(tz, batch, o), O.stroke_batch(batch, o), bbox_extent(batch), what=str(o))
    logo = PathBatch.from_paths([W.fixture_paths(GOLDEN)[0]])
    vb = gpu_stroke(tz, logo, StrokeOptions(tolerance=0.05))
    assert (len(vb.vertices), len(vb.indices)) == (2546, 7554)


@pytest.mark.parametrize("seed", range(6))
def test_random_adversarial_paths(tz, seed):
    batch = W.random_small_paths(300, 1000 + seed)
    for tol, w, ml in ((0.1, 1.0, 4.0), (0.02, 3.0, 2.0), (0.5, 0.3, 1.0)):
        for join in LineJoin:
            for cap in LineCap:
                o = StrokeOptions(tolerance=tol, line_width=w, line_join=join, miter_limit=ml).with_line_cap(cap)
                got, pve, pie = gpu_stroke(tz, batch, o, with_ranges=True)
                want, wpve, wpie = O.stroke_batch(batch, o, with_ranges=True)
                assert_mesh_parity(got, want, bbox_extent(batch), what=f"seed {seed} {o}")
                assert np.array_equal(pve, wpve) and np.array_equal(pie, wpie)


def test_edge_cases(tz):
    o = StrokeOptions()
    # empty batch / empty path
    empty = PathBatch.from_paths([Path.builder().build()])
    assert len(gpu_stroke(tz, empty, o).vertices) == 0
    # reference's pinned counts (stroke.rs:2852-2974)
    b = Path.builder()
    b.begin((-1.0, 1.0)); b.line_to((1.0, 1.0)); b.line_to((1.0, -1.0)); b.line_to((-1.0, -1.0)); b.end(False)
    b.begin((-1.0, -1.0)); b.line_to((1.0, -1.0)); b.line_to((1.0, 1.0)); b.line_to((-1.0, 1.0)); b.end(False)
    sq = PathBatch.from_paths([b.build()])
    assert len(gpu_stroke(tz, sq, o.with_line_join(LineJoin.Miter).with_line_cap(LineCap.Butt)).indices) == 36
    assert len(gpu_stroke(tz, sq, o.with_line_join(LineJoin.Bevel).with_line_cap(LineCap.Square)).indices) == 48
    b = Path.builder()
    b.begin((1.0, 0.0)); b.end(True)
    b.begin((2.0, 0.0)); b.line_to((2.0, 0.0)); b.end(False)
    b.begin((3.0, 0.0)); b.line_to((3.0, 0.0)); b.end(True)
    b.begin((3.0, 0.0)); b.line_to((3.0, 0.0)); b.line_to((3.0, 0.0)); b.line_to((3.0, 0.0)); b.end(True)
    b.begin((4.0, 0.0)); b.end(False)
    caps = PathBatch.from_paths([b.build()])
    assert len(gpu_stroke(tz, caps, o.with_line_cap(LineCap.Butt)).indices) == 0
    assert len(gpu_stroke(tz, caps, o.with_line_cap(LineCap.Square)).indices) == 24
    for cap in LineCap:
        assert_mesh_parity(gpu_stroke(tz, caps, o.with_line_cap(cap)), O.stroke_batch(caps, o.with_line_cap(cap)), 4.0)
    # issue_819 closing-point snap
    b = Path.builder()
    b.begin((650.539978027344, 173.559997558594)); b.line_to((650.609985351562, 173.529998779297))
    b.line_to((650.729980468750, 173.630004882812)); b.line_to((650.559997558594, 173.570007324219)); b.end(True)
    snap = PathBatch.from_paths([b.build()])
    for oo in all_options(0.1):
        assert_mesh_parity(gpu_stroke(tz, snap, oo), O.stroke_batch(snap, oo), 1.0)
    # correct_miter_clip_length: exact f32 max-y (stroke.rs:3381-3416)
    b = Path.builder(); b.begin((-70.0, -200.0)); b.line_to((0.0, 300.0)); b.line_to((70.0, -200.0)); b.end(False)
    mc = gpu_stroke(tz, PathBatch.from_paths([b.build()]),
                    StrokeOptions().with_line_width(30.0).with_line_join(LineJoin.MiterClip).with_miter_limit(3.0))
    assert mc.positions()[:, 1].max() == np.float32(345.0)


def test_verb_round_boundaries(tz):
    """The verb phase and the pre-pass run a path's verbs in rounds of 32 with its first Begin and last End / Close set
    aside: paths whose verb counts sit around the round boundaries (1 .. 3, 32 .. 35, 64 .. 68, 96 .. 99 verbs), single and
    multiple sub-paths, open and closed, mixed segment kinds."""
    rng = np.random.default_rng(77)
    paths = []
    for nseg in (0, 1, 2, 30, 31, 32, 33, 34, 62, 63, 64, 65, 66, 67, 94, 95, 96, 97):
        for closed in (False, True):
            b = Path.builder()
            p0 = rng.uniform(-50, 50, 2)
            b.begin((float(p0[0]), float(p0[1])))
            for k in range(nseg):
                q = rng.uniform(-50, 50, 6)
                kind = (k + nseg) % 3
                if kind == 0:
                    b.line_to((float(q[0]), float(q[1])))
                elif kind == 1:
                    b.quadratic_bezier_to((float(q[0]), float(q[1])), (float(q[2]), float(q[3])))
                else:
                    b.cubic_bezier_to((float(q[0]), float(q[1])), (float(q[2]), float(q[3])), (float(q[4]), float(q[5])))
            b.end(closed)
            if nseg in (31, 64):  # a second sub-path: the set-aside End is not the only End
                b.begin((1.0, 2.0)); b.line_to((3.0, 5.0)); b.line_to((-4.0, 1.0)); b.end(not closed)
            paths.append(b.build())
    batch = PathBatch.from_paths(paths)
    for o in (StrokeOptions(line_width=2.0, line_join=LineJoin.Round, tolerance=0.1).with_line_cap(LineCap.Round),
              StrokeOptions(line_width=1.0, line_join=LineJoin.MiterClip, tolerance=0.02).with_line_cap(LineCap.Square)):
        assert_mesh_parity(gpu_stroke(tz, batch, o), O.stroke_batch(batch, o), bbox_extent(batch), what=str(o))


def test_ragged_batch_and_tile_boundaries(tz):
    """Paths of very different sizes so that sub-paths straddle scan tiles (256 slots / 256 verbs)."""
    rng = np.random.default_rng(9)
    paths = []
    for n in [1, 255, 256, 257, 2, 511, 513, 1000, 3, 0, 64]:
        b = Path.builder()
        if n:
            cur = rng.uniform(0, 100, 2).astype(np.float32)
            b.begin(tuple(cur))
            for _ in range(n):
                cur = (cur + rng.uniform(-5, 5, 2)).astype(np.float32)
                b.line_to(tuple(cur))
            b.end(bool(n % 2))
        paths.append(b.build())
    batch = PathBatch.from_paths(paths)
    for o in all_options(0.1, 1.5):
        got, pve, pie = gpu_stroke(tz, batch, o, with_ranges=True)
        want, wpve, wpie = O.stroke_batch(batch, o, with_ranges=True)
        assert_mesh_parity(got, want, bbox_extent(batch), what=str(o))
        assert np.array_equal(pve, wpve) and np.array_equal(pie, wpie)


def test_c3_sample(tz):
    """C3 at a size the oracle finishes in seconds: 20k paths x 64 mixed segments, width 2, round joins/caps."""
    batch = W.c3_mixed_paths(20000, 64, seed=2024)
    o = StrokeOptions(line_width=2.0, line_join=LineJoin.Round).with_line_cap(LineCap.Round)
    assert_mesh_parity(gpu_stroke(tz, batch, o), O.stroke_batch(batch, o), bbox_extent(batch))
    batch = W.c3_mixed_paths(3000, 64, seed=11, closed_fraction=0.5)
    assert_mesh_parity(gpu_stroke(tz, batch, o), O.stroke_batch(batch, o), bbox_extent(batch))


def test_c4_sample_fixed_width(tz):
    """C4 geometry (1k-segment polylines, 10% reversals, miter limit 4, butt caps) on the fixed-width path."""
    batch = W.c4_polylines(300, 1000, with_attribute=False)
    o = StrokeOptions(line_join=LineJoin.Miter, miter_limit=4.0)
    assert_mesh_parity(gpu_stroke(tz, batch, o), O.stroke_batch(batch, o), bbox_extent(batch))
    batch = W.c4_polylines(50, 1000, with_attribute=True)  # attributes present, constant width: with-ids path
    assert_mesh_parity(gpu_stroke(tz, batch, o), O.stroke_batch(batch, o), bbox_extent(batch))


def test_c4_sample_variable_width(tz):
    """C4 as BASELINE.json states it: 1k-segment polylines, miter joins with limit 4, butt caps, line width from
    custom attribute 0 (StrokeOptions::variable_line_width) -- the thread-per-path kernels of stroke_vw.cuh."""
    batch = W.c4_polylines(300, 1000, with_attribute=True)
    o = StrokeOptions(line_join=LineJoin.Miter, miter_limit=4.0).with_variable_line_width(0)
    got, pve, pie = gpu_stroke(tz, batch, o, with_ranges=True)
    want, wpve, wpie = O.stroke_batch(batch, o, with_ranges=True)
    assert_mesh_parity(got, want, bbox_extent(batch))
    assert np.array_equal(pve, wpve) and np.array_equal(pie, wpie)


@pytest.mark.parametrize("num_attributes,attr", [(1, 0), (2, 1), (3, 2)])
def test_variable_width_adversarial_paths(tz, num_attributes, attr):
    batch = W.random_width_paths(400, seed=70 + num_attributes, num_attributes=num_attributes)
    for tol, w, ml in ((0.1, 1.0, 4.0), (0.02, 3.0, 2.0)):
        for join in LineJoin:
            for cap in LineCap:
                o = StrokeOptions(tolerance=tol, line_width=w, line_join=join, miter_limit=ml).with_line_cap(cap)
                o = o.with_variable_line_width(attr)
                assert_mesh_parity(gpu_stroke(tz, batch, o), O.stroke_batch(batch, o), 100.0,
                                   what=str(o))


@pytest.mark.parametrize("num_attributes", [1, 2, 3])
def test_interpolated_attributes(tz, num_attributes):
    """StrokeVertex::interpolated_attributes (stroke.rs:2728-2747) on the device against the oracle, bit for bit: endpoint
    vertices copy their endpoint's attributes, curve vertices lerp between the curve's endpoints with the vertex's t."""
    import torch
    if ROUTE != "tile":
        pytest.skip("one route")
    batch = W.random_width_paths(300, seed=90 + num_attributes, num_attributes=num_attributes)
    for o in (StrokeOptions(tolerance=0.05, line_width=1.5).with_variable_line_width(0),
              StrokeOptions(tolerance=0.1, line_width=2.0, line_join=LineJoin.Round).with_line_cap(LineCap.Round)):
        d = tz.upload(batch)
        v, i, pve, pie = tz.stroke(d, o, with_ranges=True)
        got = tz.interpolate_attributes(d, v, pve)
        torch.cuda.synchronize()
        tz.check()
        want_mesh, want = O.stroke_batch(batch, o, with_attributes=True)
        assert got.shape == want.shape and want.shape[0] == v.shape[0]
        edge = want_mesh.vertices["t"] != 1.0
        assert edge.any() and (~edge).any()
        assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32)), str(o)


def test_interpolated_attributes_reference_case(tz):
    """stroke_vertex_source_01 (stroke.rs:3040-3124) through the device path."""
    import torch
    from lyon_b200 import Path, PathBatch
    b = Path.builder_with_attributes(1)
    b.begin((0.0, 0.0), [1.0]); b.line_to((10.0, 10.0), [2.0])
    b.quadratic_bezier_to((10.0, 20.0), (0.0, 20.0), [3.0]); b.end(True)
    batch = PathBatch.from_paths([b.build()])
    d = tz.upload(batch)
    v, i, pve, pie = tz.stroke(d, StrokeOptions().with_variable_line_width(0), with_ranges=True)
    a = tz.interpolate_attributes(d, v, pve).cpu().numpy()
    pos = v.cpu().numpy()[:, 0:2]
    for k in range(len(pos)):
        def eq(q): return abs(pos[k][0] - q[0]) < 1e-5 and abs(pos[k][1] - q[1]) < 1e-5
        if eq((0, 0)): assert a[k, 0] == 1.0
        elif eq((10, 10)): assert a[k, 0] == 2.0
        elif eq((0, 20)): assert a[k, 0] == 3.0
        else: assert 2.0 < a[k, 0] < 3.0


@pytest.mark.parametrize("tol", [1.0, 0.3, 0.1, 0.03, 0.01, 0.003, 0.001])
def test_c5_tolerance_sweep(tz, tol):
    batch = W.c5_logo_copies(W.fixture_paths(GOLDEN), 40)
    o = StrokeOptions(tolerance=tol)
    assert_mesh_parity(gpu_stroke(tz, batch, o), O.stroke_batch(batch, o), bbox_extent(batch))


def test_vertex_base_and_too_many_vertices(tz):
    from lyon_b200 import TooManyVertices
    batch = W.c3_mixed_paths(100, 8, seed=1)
    o = StrokeOptions()
    a = gpu_stroke(tz, batch, o)
    b = gpu_stroke(tz, batch, o, vertex_base=12345)
    assert np.array_equal(a.indices + np.uint32(12345), b.indices)
    with pytest.raises(TooManyVertices):
        gpu_stroke(tz, batch, o, vertex_base=0xFFFFFFFE)


def test_flatten_paths(tz, route):
    """The path-level `Flattened` adapter (iterator.rs:125-226) on the device against the oracle: same verbs, same
    points byte for byte, same ranges; and the flattened batch strokes like the original does with Bevel joins...
    (it is a valid batch for every entry point)."""
    if route != "tile":
        pytest.skip("one route")
    import torch
    for batch in (W.random_small_paths(500, 21, degenerate=0.1), W.c3_mixed_paths(2000, 64, seed=4, closed_fraction=0.3),
                  PathBatch.from_paths(W.fixture_paths(GOLDEN)), PathBatch.from_paths([Path.builder().build()])):
        for tol in (0.5, 0.05, 0.002):
            d = tz.upload(batch)
            out = tz.flatten_paths(d, tol)
            torch.cuda.synchronize()
            want = O.flatten_paths(batch, tol)
            assert np.array_equal(out.verbs.cpu().numpy(), want.verbs)
            assert out.points.cpu().numpy().tobytes() == np.ascontiguousarray(want.points, np.float32).tobytes()
            assert np.array_equal(out.path_verb_begin.cpu().numpy().astype(np.uint32), want.path_verb_begin)
            assert np.array_equal(out.path_point_begin.cpu().numpy().astype(np.uint32), want.path_point_begin)
    # the flattened batch is itself a batch: stroke it on the device and with the oracle
    batch = W.c3_mixed_paths(300, 16, seed=8)
    out = tz.flatten_paths(tz.upload(batch), 0.1)
    o = StrokeOptions(line_join=LineJoin.Bevel)
    v, i = tz.stroke(out, o)
    torch.cuda.synchronize()
    from lyon_b200 import device_mesh_to_host
    assert_mesh_parity(device_mesh_to_host(v, i), O.stroke_batch(O.flatten_paths(batch, 0.1), o), bbox_extent(batch))


def test_flatten_paths_with_attributes(tz, route):
    """Path-level flattening of paths that carry custom attributes (crates/path/src/private.rs:56-122): byte-identical to
    the oracle for 1, 2 and 3 attributes (odd counts are padded), incl. the reference's `flattened_custom_attributes`
    path (path.rs:2106-2142); and the flattened batch strokes with a variable line width like the oracle's."""
    if route != "tile":
        pytest.skip("one route")
    import torch
    b = Path.builder_with_attributes(1)
    b.begin((0.0, 0.0), [0.0])
    b.quadratic_bezier_to((1.0, 0.0), (1.0, 1.0), [1.0])
    b.cubic_bezier_to((1.0, 2.0), (0.0, 2.0), (0.0, 1.0), [2.0])
    b.end(False)
    batches = [(PathBatch.from_paths([b.build()]), 0.01)]
    for na in (1, 2, 3):
        batches.append((W.random_width_paths(300, seed=40 + na, num_attributes=na), 0.05))
    for batch, tol in batches:
        out = tz.flatten_paths(tz.upload(batch), tol)
        torch.cuda.synchronize()
        want = O.flatten_paths(batch, tol)
        assert out.num_attributes == batch.num_attributes
        assert np.array_equal(out.verbs.cpu().numpy(), want.verbs)
        assert out.points.cpu().numpy().tobytes() == np.ascontiguousarray(want.points, np.float32).tobytes()
        assert np.array_equal(out.path_verb_begin.cpu().numpy().astype(np.uint32), want.path_verb_begin)
        assert np.array_equal(out.path_point_begin.cpu().numpy().astype(np.uint32), want.path_point_begin)
    # a flattened batch with attributes is a batch: variable line width on the device and with the oracle
    batch = W.random_width_paths(200, seed=77, num_attributes=1)
    out = tz.flatten_paths(tz.upload(batch), 0.05)
    o = StrokeOptions(line_join=LineJoin.Miter).with_variable_line_width(0)
    v, i = tz.stroke(out, o)
    torch.cuda.synchronize()
    from lyon_b200 import device_mesh_to_host
    assert np.array_equal(device_mesh_to_host(v, i).indices, O.stroke_batch(O.flatten_paths(batch, 0.05), o).indices)


def test_measure_paths(tz, route):
    """`PathMeasurements::from_path` on the device (lyon_cuda_measure_paths_*, a consumer of flattening: measure.rs:140-260)
    against the oracle: the same edges, byte for byte -- the f32 running distance is replayed in event order."""
    if route != "tile":
        pytest.skip("one route")
    import torch
    for batch in (W.random_small_paths(400, 33, degenerate=0.1), W.c3_mixed_paths(1500, 64, seed=6, closed_fraction=0.3),
                  W.random_width_paths(200, seed=51, num_attributes=3), PathBatch.from_paths(W.fixture_paths(GOLDEN)),
                  PathBatch.from_paths([Path.builder().build()])):
        for tol in (0.25, 0.01, 1e-6):
            got = tz.measure_paths(tz.upload(batch), tol)
            torch.cuda.synchronize()
            want = O.measure_paths(batch, tol)
            assert np.array_equal(got[3].cpu().numpy().astype(np.uint64), want[3])
            assert got[0].cpu().numpy().tobytes() == want[0].tobytes(), "distance"
            assert np.array_equal(got[1].cpu().numpy().astype(np.uint32), want[1]), "event index"
            assert got[2].cpu().numpy().tobytes() == want[2].tobytes(), "t"
            assert got[4].cpu().numpy().tobytes() == want[4].tobytes(), "path length"


def test_rebase_indices(tz, route):
    """lyon_cuda_rebase_indices: a mesh emitted with local ids, rebased afterwards == emitted with the base (the
    multi-GPU single-pass flow), for every alignment of the index buffer."""
    if route != "tile":
        pytest.skip("one route")
    import torch
    from lyon_b200 import TooManyVertices
    batch = W.c3_mixed_paths(500, 16, seed=3)
    o = StrokeOptions(line_join=LineJoin.Round)
    d = tz.upload(batch)
    v0, i0 = tz.stroke_fused(d, o)
    i0 = i0.clone()
    v1, i1 = tz.stroke_fused(d, o, vertex_base=54321)
    want = i1.clone()
    for shift in range(4):  # index views starting at every 4-byte offset inside a 16-byte line
        buf = torch.zeros(i0.shape[0] + 8, dtype=torch.int32, device=i0.device)
        view = buf[shift:shift + i0.shape[0]]
        view.copy_(i0)
        tz.rebase_indices(view, int(v0.shape[0]), 54321)
        torch.cuda.synchronize()
        assert torch.equal(view, want) and int(buf[:shift].abs().sum()) == 0 and int(buf[shift + i0.shape[0]:].abs().sum()) == 0
    for n in (0, 1, 2, 3, 5):
        small = i0[:n].clone()
        tz.rebase_indices(small, int(v0.shape[0]), 7)
        torch.cuda.synchronize()
        assert torch.equal(small, i0[:n] + 7)
    with pytest.raises(TooManyVertices):
        tz.rebase_indices(i0.clone(), int(v0.shape[0]), 0xFFFFFFFE)


def test_host_buffer_entry_point(tz):
    from lyon_b200 import StrokeTessellator, tessellate_paths
    batch = W.random_small_paths(200, 77)
    o = StrokeOptions(line_join=LineJoin.Round).with_line_cap(LineCap.Square)
    assert_mesh_parity(tessellate_paths(batch, o, tz), O.stroke_batch(batch, o), bbox_extent(batch))
    p = W.fixture_paths(GOLDEN)[0]
    vb = StrokeTessellator(tz).tessellate_path(p, StrokeOptions(tolerance=0.05))
    assert (len(vb.vertices), len(vb.indices)) == (2546, 7554)


@pytest.mark.parametrize("npts", [3, 4])
def test_flatten_curves_fused(tz, npts, route):
    """Single-pass flattening (count + look-back + emit in one kernel, buffers kept across calls): same bytes as the
    oracle, including after the kept buffers had to grow and when they are larger than needed."""
    if route != "tile":
        pytest.skip("flattening has a single route")
    import torch
    rng = np.random.default_rng(5)
    for n in (70000, 1000, 257, 1, 0):
        ctrl = rng.uniform(-300, 300, size=(n, 2 * npts)).astype(np.float32)
        d = torch.from_numpy(ctrl).cuda().reshape(n, 2 * npts)
        off, pts, t = tz.flatten_curves_fused(d, npts, 0.02, with_t=True)
        torch.cuda.synchronize()
        ooff, opts, ot = (O.flatten_cubics if npts == 4 else O.flatten_quadratics)(ctrl, 0.02, with_t=True)
        assert np.array_equal(off.cpu().numpy().astype(np.uint64), ooff)
        assert pts.cpu().numpy().tobytes() == opts.tobytes() and t.cpu().numpy().tobytes() == ot.tobytes()


@pytest.mark.parametrize("npts", [3, 4])
def test_flatten_curves(tz, npts, route):
    if route != "tile":
        pytest.skip("flattening has a single route")
    from lyon_b200 import flatten_cubics, flatten_quadratics
    rng = np.random.default_rng(3)
    ctrl = rng.uniform(-500, 500, size=(100000, 2 * npts)).astype(np.float32)
    ctrl[:10] = 0
    for tol in (0.5, 0.01, 0.0005):
        off, pts, t = (flatten_cubics if npts == 4 else flatten_quadratics)(ctrl, tol, tz, with_t=True)
        ooff, opts, ot = (O.flatten_cubics if npts == 4 else O.flatten_quadratics)(ctrl, tol, with_t=True)
        assert np.array_equal(off, ooff) and pts.tobytes() == opts.tobytes() and t.tobytes() == ot.tobytes()


def test_c2_sample_and_full_size_properties(tz, route):
    if route != "tile":
        pytest.skip("flattening has a single route")
    """C2: exact parity on 200k curves; at the full 10M curves only size-independent properties:
    offsets == Wang counts recomputed in numpy for a sample, last point of every curve == `to`."""
    import torch
    from lyon_b200 import flatten_cubics
    ctrl = W.c2_cubics(200000)
    off, pts = flatten_cubics(ctrl, 0.01, tz)
    ooff, opts = O.flatten_cubics(ctrl, 0.01)
    assert np.array_equal(off, ooff) and pts.tobytes() == opts.tobytes()
    n = 10_000_000
    ctrl = W.c2_cubics(n)
    d = torch.from_numpy(ctrl).cuda()
    doff, dpts = tz.flatten_curves(d, 4, 0.01)
    torch.cuda.synchronize()
    counts = (doff[1:] - doff[:-1])
    assert int(counts.min()) >= 1
    last = dpts[(doff[1:] - 1)]
    assert torch.equal(last, d[:, 6:8])  # test_flatten_end: last point == `to` for every curve
    sample = np.arange(0, n, 997)
    so, _ = O.flatten_cubics(ctrl[sample], 0.01)
    assert np.array_equal(np.diff(so).astype(np.int64), counts[torch.from_numpy(sample).cuda()].cpu().numpy())
