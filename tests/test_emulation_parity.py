"""CPU (no GPU) parity: the CUDA per-item code, compiled for the host by tests/emul (same source as the
kernels), against the oracle.  Catches logic errors in the container; the -m gpu tests then only have to
prove the parallel plumbing (scans, indexing) on real hardware."""
import numpy as np
import pytest

import emul_ffi as E
import oracle_ffi as O
import workloads as W
from lyon_b200 import LineCap, LineJoin, Path, PathBatch, StrokeOptions
from parity import assert_mesh_parity, bbox_extent

GOLDEN = O.ROOT / "tests" / "golden"


@pytest.fixture(autouse=True, params=["general", "tile", "wtile"])
def pipeline(request, monkeypatch):
    """Every test runs against both device pipelines' per-item code: the general multi-kernel one
    (stroke_core.cuh / Views) and the fused tile one (stroke_tile.cuh / TileViews)."""
    monkeypatch.setattr(E, "PIPELINE", request.param)
    return request.param


def all_options(tol=0.1, width=1.0):
    for join in LineJoin:
        for cap in LineCap:
            yield StrokeOptions(tolerance=tol, line_width=width, line_join=join).with_line_cap(cap)


@pytest.mark.parametrize("tol", [0.05, 0.1, 0.01])
def test_fixtures(tol):
    for path in W.fixture_paths(GOLDEN):
        batch = PathBatch.from_paths([path])
        for o in all_options(tol):
            assert_mesh_parity(E.stroke_batch(batch, o), O.stroke_batch(batch, o), bbox_extent(batch), what=str(o))


@pytest.mark.parametrize("seed", range(12))
def test_random_adversarial_paths(seed):
    batch = W.random_small_paths(40, seed)
    for tol, w, ml in ((0.1, 1.0, 4.0), (0.02, 3.0, 2.0), (0.5, 0.3, 1.0)):
        for join in LineJoin:
            for cap in LineCap:
                o = StrokeOptions(tolerance=tol, line_width=w, line_join=join, miter_limit=ml).with_line_cap(cap)
                got, pve, pie = E.stroke_batch(batch, o, with_ranges=True)
                want, wpve, wpie = O.stroke_batch(batch, o, with_ranges=True)
                assert_mesh_parity(got, want, bbox_extent(batch), what=f"seed {seed} {o}")
                assert np.array_equal(pve, wpve) and np.array_equal(pie, wpie)


def test_adversarial_paths_hit_the_rare_branches():
    """The random generator must actually exercise merge drops, folds, closing snaps and sharp-turn splits."""
    batch = W.random_small_paths(400, 123)
    o = StrokeOptions(tolerance=0.1, line_join=LineJoin.Round).with_line_cap(LineCap.Round)
    want = O.stroke_batch(batch, o)
    # quads with a far control point exist -> find_sharp_turn split
    L = O.lib()
    n_split = 0
    for i in range(batch.n_paths):
        p = batch.path(i)
        for ev in p.iter():
            if ev[0] == "quadratic":
                c = np.array([*ev[1], *ev[2], *ev[3]], np.float32)
                t = np.zeros(1, np.float32)
                n_split += L.oracle_find_sharp_turn(c.ctypes.data, t.ctypes.data)
    assert n_split > 5
    assert len(want.vertices) > 10000


def test_known_answer_cases_through_the_device_code():
    """The reference's pinned counts (stroke.rs:2852-2974) hold for the device code as well (fixed width)."""
    def square():
        b = Path.builder()
        b.begin((-1.0, 1.0)); b.line_to((1.0, 1.0)); b.line_to((1.0, -1.0)); b.line_to((-1.0, -1.0)); b.end(False)
        b.begin((-1.0, -1.0)); b.line_to((1.0, -1.0)); b.line_to((1.0, 1.0)); b.line_to((-1.0, 1.0)); b.end(False)
        return PathBatch.from_paths([b.build()])
    o = StrokeOptions()
    assert len(E.stroke_batch(square(), o.with_line_join(LineJoin.Miter).with_line_cap(LineCap.Butt)).indices) == 36
    assert len(E.stroke_batch(square(), o.with_line_join(LineJoin.Bevel).with_line_cap(LineCap.Square)).indices) == 48
    b = Path.builder()
    b.begin((1.0, 0.0)); b.end(True)
    b.begin((2.0, 0.0)); b.line_to((2.0, 0.0)); b.end(False)
    b.begin((3.0, 0.0)); b.line_to((3.0, 0.0)); b.end(True)
    b.begin((3.0, 0.0)); b.line_to((3.0, 0.0)); b.line_to((3.0, 0.0)); b.line_to((3.0, 0.0)); b.end(True)
    b.begin((4.0, 0.0)); b.end(False)
    caps = PathBatch.from_paths([b.build()])
    assert len(E.stroke_batch(caps, o.with_line_cap(LineCap.Butt)).indices) == 0
    assert len(E.stroke_batch(caps, o.with_line_cap(LineCap.Square)).indices) == 24
    for cap in LineCap:
        assert_mesh_parity(E.stroke_batch(caps, o.with_line_cap(cap)), O.stroke_batch(caps, o.with_line_cap(cap)), 4.0)
    empty = PathBatch.from_paths([Path.builder().build()])
    assert len(E.stroke_batch(empty, o).vertices) == 0


def test_issue_819_closing_point_snap():  # stroke.rs:3290-3321
    b = Path.builder()
    b.begin((650.539978027344, 173.559997558594)); b.line_to((650.609985351562, 173.529998779297))
    b.line_to((650.729980468750, 173.630004882812)); b.line_to((650.559997558594, 173.570007324219)); b.end(True)
    batch = PathBatch.from_paths([b.build()])
    for o in all_options(0.1):
        assert_mesh_parity(E.stroke_batch(batch, o), O.stroke_batch(batch, o), 1.0)


@pytest.mark.parametrize("n_paths,segs", [(300, 64), (20, 300)])
def test_c3_like_batches(n_paths, segs):
    batch = W.c3_mixed_paths(n_paths, segs, seed=5, closed_fraction=0.3)
    o = StrokeOptions(line_width=2.0, line_join=LineJoin.Round).with_line_cap(LineCap.Round)
    assert_mesh_parity(E.stroke_batch(batch, o), O.stroke_batch(batch, o), bbox_extent(batch))


def test_c4_like_polylines_fixed_width():
    batch = W.c4_polylines(50, 400, with_attribute=False)
    o = StrokeOptions(line_join=LineJoin.Miter, miter_limit=4.0)
    assert_mesh_parity(E.stroke_batch(batch, o), O.stroke_batch(batch, o), bbox_extent(batch))
    # with an attribute but no variable width -> tessellate_with_ids_fw: endpoint ids are point offsets
    batch = W.c4_polylines(20, 100, with_attribute=True)
    assert_mesh_parity(E.stroke_batch(batch, o), O.stroke_batch(batch, o), bbox_extent(batch))


@pytest.mark.parametrize("tol", [1.0, 0.1, 0.003])
def test_c5_tolerance_sweep(tol):
    batch = W.c5_logo_copies(W.fixture_paths(GOLDEN), 6)
    o = StrokeOptions(tolerance=tol)
    assert_mesh_parity(E.stroke_batch(batch, o), O.stroke_batch(batch, o), bbox_extent(batch))


@pytest.mark.parametrize("num_attributes,attr", [(1, 0), (2, 1), (3, 2)])
def test_variable_line_width_state_machine(num_attributes, attr):
    """stroke_vw.cuh (the code k_vw_count / k_vw_emit run per path) against the oracle's variable-width path."""
    batch = W.random_width_paths(120, seed=40 + num_attributes, num_attributes=num_attributes)
    for o in all_options(0.1):
        o = o.with_variable_line_width(attr)
        assert_mesh_parity(E.stroke_batch(batch, o), O.stroke_batch(batch, o), 100.0)


def test_c4_polylines_variable_width():
    batch = W.c4_polylines(40, 300, with_attribute=True)
    o = StrokeOptions(line_join=LineJoin.Miter, miter_limit=4.0).with_variable_line_width(0)
    assert_mesh_parity(E.stroke_batch(batch, o), O.stroke_batch(batch, o), bbox_extent(batch))


def test_vertex_base_rebasing_and_overflow():
    batch = W.c3_mixed_paths(10, 8, seed=1)
    o = StrokeOptions()
    a = E.stroke_batch(batch, o)
    b = E.stroke_batch(batch, o, vertex_base=1000)
    assert np.array_equal(a.indices + np.uint32(1000), b.indices)
    with pytest.raises(E.EmulError) as e:
        E.stroke_batch(batch, o, vertex_base=0xFFFFFFFE)  # ids would pass u32::MAX - 1
    assert e.value.code == -2


@pytest.mark.parametrize("npts", [3, 4])
def test_flatten_device_code_matches_oracle(npts):
    rng = np.random.default_rng(3)
    ctrl = rng.uniform(-500, 500, size=(3000, 2 * npts)).astype(np.float32)
    ctrl[:10] = 0  # degenerate curves
    for tol in (0.5, 0.01, 0.0005):
        off, pts, t = E.flatten(ctrl, npts, tol)
        ooff, opts, ot = (O.flatten_cubics if npts == 4 else O.flatten_quadratics)(ctrl, tol, with_t=True)
        assert np.array_equal(off, ooff) and pts.tobytes() == opts.tobytes() and t.tobytes() == ot.tobytes()
