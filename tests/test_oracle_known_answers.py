"""Pins the CPU oracle against every known answer the reference's own tests hold for the hot path
(SURVEY §4 / §8c).  Each test names the reference test it restates (file:line under /root/reference)."""
import math

import numpy as np
import pytest

import oracle_ffi as O
from lyon_b200 import LineCap, LineJoin, Path, PathBatch, StrokeOptions, parse_svg_path

GOLDEN = O.ROOT / "tests" / "golden"


def check_invariants(path, opts, expected_triangles=None, winding=True):
    """`test_path` + `TestBuilder` (crates/tessellation/src/stroke.rs:2787-2849)."""
    vb = O.stroke_path(path, opts)
    v, idx = vb.vertices, vb.indices.reshape(-1, 3)
    pos = vb.positions()
    assert not np.isnan(pos).any()
    assert not np.isnan(v["normal"]).any()
    assert ((v["normal"] ** 2).sum(axis=1) != 0).all()
    assert not np.isnan(v["advancement"]).any()
    a, b, c = idx[:, 0], idx[:, 1], idx[:, 2]
    assert (a != b).all() and (a != c).all() and (b != c).all()
    pa, pb, pc = pos[a], pos[b], pos[c]
    d0, d1 = pa - pb, pc - pb
    crossz = d0[:, 0] * d1[:, 1] - d0[:, 1] * d1[:, 0]
    if winding:  # the reference only asserts winding on its own small test inputs
        assert (crossz >= -0.035).all()
    if expected_triangles is not None:
        assert idx.shape[0] == expected_triangles
    return vb


def square_path():
    b = Path.builder_with_attributes(1)
    b.begin((-1.0, 1.0), [0.3]); b.line_to((1.0, 1.0), [0.3]); b.line_to((1.0, -1.0), [0.3]); b.line_to((-1.0, -1.0), [0.3]); b.end(False)
    b.begin((-1.0, -1.0), [0.3]); b.line_to((1.0, -1.0), [0.3]); b.line_to((1.0, 1.0), [0.3]); b.line_to((-1.0, 1.0), [0.3]); b.end(False)
    return b.build()


@pytest.mark.parametrize("vw", [True, False])
def test_square(vw):  # stroke.rs:2852-2910
    path = square_path()
    o = StrokeOptions().with_variable_line_width(0) if vw else StrokeOptions()
    check_invariants(path, o.with_line_join(LineJoin.Miter).with_line_cap(LineCap.Butt), 12)
    check_invariants(path, o.with_line_join(LineJoin.Bevel).with_line_cap(LineCap.Square), 16)
    check_invariants(path, o.with_line_join(LineJoin.MiterClip).with_miter_limit(1.0).with_line_cap(LineCap.Round))
    check_invariants(path, o.with_tolerance(0.001).with_line_join(LineJoin.Round).with_line_cap(LineCap.Round))


def test_empty_path():  # stroke.rs:2913-2919
    check_invariants(Path.builder().build(), StrokeOptions(), 0)
    check_invariants(Path.builder_with_attributes(1).build(), StrokeOptions(), 0)


@pytest.mark.parametrize("vw", [True, False])
def test_empty_caps(vw):  # stroke.rs:2922-2974
    b = Path.builder_with_attributes(1)
    b.begin((1.0, 0.0), [1.0]); b.end(True)
    b.begin((2.0, 0.0), [1.0]); b.line_to((2.0, 0.0), [1.0]); b.end(False)
    b.begin((3.0, 0.0), [1.0]); b.line_to((3.0, 0.0), [1.0]); b.end(True)
    b.begin((3.0, 0.0), [1.0]); b.line_to((3.0, 0.0), [1.0]); b.line_to((3.0, 0.0), [1.0]); b.line_to((3.0, 0.0), [1.0]); b.end(True)
    b.begin((4.0, 0.0), [1.0]); b.end(False)
    path = b.build()
    o = StrokeOptions().with_variable_line_width(0) if vw else StrokeOptions()
    check_invariants(path, o.with_line_cap(LineCap.Butt), 0)
    check_invariants(path, o.with_line_cap(LineCap.Square), 8)
    check_invariants(path, o.with_line_cap(LineCap.Round))


def test_too_many_vertices():  # stroke.rs:2977-3037 (builder that runs out of vertex ids)
    path = parse_svg_path((GOLDEN / "rust_logo.path").read_text())
    opts = StrokeOptions(tolerance=0.05)
    for max_vertices in (0, 10, 100):
        with pytest.raises(O.OracleError) as e:
            # ids 0..max_index may be handed out => max_vertices ids means max_index = max_vertices - 1
            O.stroke_path(path, opts, max_index=max_vertices - 1 if max_vertices else -1 & 0)
        assert e.value.code == -2
    O.stroke_path(path, opts)  # fits u32


def test_line_width():  # stroke.rs:3127-3154
    b = Path.builder(); b.begin((0.0, 1.0)); b.line_to((2.0, 1.0)); b.end(False)
    vb = O.stroke_path(b.build(), StrokeOptions().with_line_width(2.0))
    corners = np.array([(0, 0), (0, 2), (2, 0), (2, 2)], np.float32)
    for p in vb.positions():
        assert (np.abs(corners - p).max(axis=1) < 1e-6).any(), p
    assert len(vb.vertices) == 4 and len(vb.indices) == 6


def test_triangle_winding():  # stroke.rs:3215-3257
    b = Path.builder().with_svg(); b.move_to((0.0, 0.0)); b.quadratic_bezier_to((100.0, 0.0), (100.0, 100.0))
    vb = O.stroke_path(b.build(), StrokeOptions(tolerance=0.05))
    pos = vb.positions(); idx = vb.indices.reshape(-1, 3)
    pa, pb, pc = pos[idx[:, 0]], pos[idx[:, 1]], pos[idx[:, 2]]
    u, w = pb - pa, pc - pb
    assert ((u[:, 0] * w[:, 1] - u[:, 1] * w[:, 0]) <= 0).all()


def test_single_segment_closed():  # stroke.rs:3260-3287
    b = Path.builder().with_svg(); b.move_to((0.0, 0.0)); b.line_to((100.0, 0.0))
    assert len(O.stroke_path(b.build(), StrokeOptions(tolerance=0.05)).indices) > 0
    b = Path.builder_with_attributes(1); b.begin((0.0, 0.0), [1.0]); b.line_to((100.0, 0.0), [1.0]); b.end(False)
    assert len(O.stroke_path(b.build(), StrokeOptions(tolerance=0.05)).indices) > 0


@pytest.mark.parametrize("attrs", [False, True])
def test_issue_819(attrs):  # stroke.rs:3290-3321
    pts = [(650.539978027344, 173.559997558594), (650.609985351562, 173.529998779297),
           (650.729980468750, 173.630004882812), (650.559997558594, 173.570007324219)]
    b = Path.builder_with_attributes(1) if attrs else Path.builder()
    a = [0.0] if attrs else None
    b.begin(pts[0], a)
    for p in pts[1:]:
        b.line_to(p, a)
    b.end(True)
    vb = O.stroke_path(b.build(), StrokeOptions(tolerance=0.1))
    assert not np.isnan(vb.positions()).any()


def test_issue_821():  # stroke.rs:3324-3348
    b = Path.builder_with_attributes(1)
    pts = [(-45.192276, -69.800575), (-45.164116, -69.84931), (-45.135952, -69.90564), (-45.10779, -69.93248),
           (-45.052788, -69.96064), (-45.026604, -69.898155)]
    b.begin(pts[0], [1.0])
    for p in pts[1:]:
        b.line_to(p, [1.0])
    b.end(False)
    o = StrokeOptions().with_tolerance(0.001).with_line_cap(LineCap.Round).with_variable_line_width(0)
    O.stroke_path(b.build(), o)


def test_issue_894():  # stroke.rs:3351-3378
    b = Path.builder_with_attributes(1)
    b.begin((435.72, 368.42), [38.82]); b.line_to((433.53, 366.06), [38.82])
    b.quadratic_bezier_to((431.35, 363.70), (430.22, 362.52), [39.59])
    b.quadratic_bezier_to((429.09, 361.34), (429.05, 362.14), [41.62])
    b.line_to((429.00, 362.95), [41.63]); b.end(False)
    o = StrokeOptions(tolerance=0.01).with_line_cap(LineCap.Round).with_variable_line_width(0)
    vb = O.stroke_path(b.build(), o)
    idx = vb.indices.reshape(-1, 3)
    assert (idx[:, 0] != idx[:, 1]).all() and (idx[:, 1] != idx[:, 2]).all() and (idx[:, 0] != idx[:, 2]).all()


def test_correct_miter_clip_length():  # stroke.rs:3381-3416 (exact f32 equality)
    b = Path.builder(); b.begin((-70.0, -200.0)); b.line_to((0.0, 300.0)); b.line_to((70.0, -200.0)); b.end(False)
    o = StrokeOptions().with_line_width(30.0).with_line_join(LineJoin.MiterClip).with_miter_limit(3.0)
    vb = O.stroke_path(b.build(), o)
    assert vb.positions()[:, 1].max() == np.float32(300.0 + 30.0 * 3.0 * 0.5)


def test_stroke_vertex_source_01():  # stroke.rs:3040-3124
    b = Path.builder_with_attributes(1)
    a = b.begin((0.0, 0.0), [1.0]); bb = b.line_to((10.0, 10.0), [2.0])
    c = b.quadratic_bezier_to((10.0, 20.0), (0.0, 20.0), [3.0]); b.end(True)
    path = b.build()
    vb = O.stroke_path(path, StrokeOptions().with_variable_line_width(0))
    v = vb.vertices
    for i in range(len(v)):
        pos = v["position_on_path"][i]; src = int(vb.source_endpoint()[i]); t = float(v["t"][i])
        def eq(q): return abs(pos[0] - q[0]) < 1e-5 and abs(pos[1] - q[1]) < 1e-5
        if eq((0, 0)): assert (src, t) == (a, 1.0)
        elif eq((10, 10)): assert (src, t) == (bb, 1.0)
        elif eq((0, 20)): assert (src, t) == (c, 1.0)
        else:
            assert src == c and 0.0 < t < 1.0
            # interpolated line width (attribute 1 -> lerp between 2 and 3): stroke.rs:1022
            w = float(v["half_width"][i]) * 2.0
            assert 2.0 < w < 3.0
    # attr.interpolated_attributes() of every vertex (stroke.rs:3105-3117)
    vb2, attrs = O.stroke_batch(PathBatch.from_paths([path]), StrokeOptions().with_variable_line_width(0), with_attributes=True)
    assert attrs.shape == (len(v), 1)
    for i in range(len(v)):
        pos = v["position_on_path"][i]
        def eq(q): return abs(pos[0] - q[0]) < 1e-5 and abs(pos[1] - q[1]) < 1e-5
        if eq((0, 0)): assert attrs[i, 0] == 1.0
        elif eq((10, 10)): assert attrs[i, 0] == 2.0
        elif eq((0, 20)): assert attrs[i, 0] == 3.0
        else: assert 2.0 < attrs[i, 0] < 3.0


def test_compute_normal():  # math_utils.rs:42-62
    for v1, v2, want in [((1, 0), (0, 1), (-1, 1)), ((1, 0), (0, -1), (1, 1)), ((1, 0), (1, 0), (0, 1))]:
        n = O.compute_normal(v1, v2)
        assert ((n - np.array(want, np.float32)) ** 2).sum() <= 1e-5


# ---- flattening -------------------------------------------------------------------------------------
def test_flattening_empty_and_straight():  # quadratic_bezier.rs test_flattening_empty_curve / _straight_line
    off, pts = O.flatten_quadratics(np.array([[0, 0, 0, 0, 0, 0]], np.float32), 0.1)
    assert off[-1] == 1
    off, pts = O.flatten_quadratics(np.array([[0, 0, 10, 0, 20, 0]], np.float32), 0.1)
    assert off[-1] == 1 and tuple(pts[0]) == (20.0, 0.0)
    off, pts = O.flatten_cubics(np.array([[0, 0, 0, 0, 0, 0, 0, 0]], np.float32), 0.1)
    assert off[-1] == 1


def test_flattened_t_values_monotonic():  # quadratic_bezier.rs:1422-1461
    curve = np.array([[0, 0, 50, 100, 100, 0]], np.float32)
    off, pts, t = O.flatten_quadratics(curve, 0.01, with_t=True)
    count = int(off[-1])
    assert count >= 3
    assert (np.diff(np.concatenate([[0.0], t])) > 0).all()  # strictly increasing from 0
    assert abs(t[0] - 1.0 / count) < 1e-3  # first ~ 1/count, not 2/count
    assert t[-1] == 1.0
    assert tuple(pts[-1]) == (100.0, 0.0)  # test_flatten_end: last point == `to`
    for tol in (0.5, 0.1, 0.001):
        off, pts, t = O.flatten_quadratics(curve, tol, with_t=True)
        assert (np.diff(t) > 0).all() and t[-1] == 1.0


def test_flatten_with_t_within_tolerance():  # cubic_bezier.rs flatten_with_t (:1574-1602): |p - sample(t)| <= tol
    rng = np.random.default_rng(7)
    ctrl = rng.uniform(-100, 100, size=(64, 8)).astype(np.float32)
    tol = 0.01
    off, pts, t = O.flatten_cubics(ctrl, tol, with_t=True)
    for i in range(64):
        c = ctrl[i].astype(np.float64).reshape(4, 2)
        tt = t[off[i]:off[i + 1]].astype(np.float64)[:, None]
        bez = ((1 - tt) ** 3) * c[0] + 3 * ((1 - tt) ** 2) * tt * c[1] + 3 * (1 - tt) * tt * tt * c[2] + tt ** 3 * c[3]
        assert np.abs(bez - pts[off[i]:off[i + 1]]).max() <= tol
        assert tuple(pts[off[i + 1] - 1]) == tuple(ctrl[i, 6:8])


def test_wang_counts_lut_boundaries():  # cubic_bezier.rs:1369-1403: LUT i^4 boundaries, non-integer fallback, inf/NaN
    L = O.lib()
    def n_cubic(scale, tol=1.0):
        c = np.array([0, 0, scale, 0, 0, 0, 0, 0], np.float32)  # v1 = -2*scale*6 , v2 = scale*6
        return L.oracle_cubic_wang(c.ctypes.data, tol)
    # err4 = (12 s)^2 / 64 ; choose s so err4 == 16 exactly -> n = 2 ; just above -> 3
    s = math.sqrt(16 * 64) / 12
    assert n_cubic(np.float32(s)) in (2.0, 3.0)
    assert n_cubic(1e-3) == 1.0
    big = n_cubic(1e4)
    assert big > 24 and big != round(big)  # fallback path returns sqrt(sqrt(err4)), not rounded
    assert n_cubic(np.float32(np.inf)) == 1.0 or math.isinf(n_cubic(np.float32(np.inf)))


def test_fast_atan2_is_approximation():
    L = O.lib()
    for y, x in [(1, 1), (1, -1), (-1, -1), (-1, 1), (0.3, 2.0), (5, 0.1), (0, 1), (1, 0)]:
        assert abs(L.oracle_fast_atan2(y, x) - math.atan2(y, x)) < 2.5e-4
    assert math.isnan(L.oracle_fast_atan2(0.0, 0.0))  # 0/0


# ---- sizes on the real fixtures (second independent restatement agrees) -----------
@pytest.mark.parametrize("tol,join,nv,ni", [
    (0.05, LineJoin.Miter, 2546, 7554), (0.05, LineJoin.Round, 3036, 9012),
    (0.1, LineJoin.Miter, 1974, 5838), (0.1, LineJoin.Round, 2330, 6906),
    (0.01, LineJoin.Miter, 5086, 15168), (0.01, LineJoin.Round, 5991, 17835),
    (0.001, LineJoin.Miter, 15004, 44922), (0.001, LineJoin.Round, 17031, 50937),
])
def test_logo_path_sizes(tol, join, nv, ni):
    path = parse_svg_path((GOLDEN / "logo.path").read_text())
    assert np.bincount(path.verbs, minlength=6).tolist() == [174, 0, 103, 14, 14, 0]
    vb = check_invariants(path, StrokeOptions(tolerance=tol, line_join=join), winding=False)
    assert (len(vb.vertices), len(vb.indices)) == (nv, ni)


def test_rust_logo_fixture_shape():  # 10 closed sub-paths, 146 cubics, 89 lines
    path = parse_svg_path((GOLDEN / "rust_logo.path").read_text())
    assert np.bincount(path.verbs, minlength=6).tolist() == [89, 0, 146, 10, 10, 0]
    for join in (LineJoin.Miter, LineJoin.Bevel, LineJoin.Round):  # bench/tess stroke_01..03
        check_invariants(path, StrokeOptions(line_join=join), winding=False)


def test_flattened_custom_attributes():
    """crates/path/src/path.rs:2106-2142 `flattened_custom_attributes`: along the flattened events of a path whose
    endpoints carry the attribute 0, 1, 2 the attribute strictly increases, every Line starts from the previous value, and
    no curve survives.  Run on the oracle's attribute-aware path flattening (private.rs:56-122)."""
    from lyon_b200.path import Path, PathBatch, LINE_TO, BEGIN, END
    b = Path.builder_with_attributes(1)
    b.begin((0.0, 0.0), [0.0])
    b.quadratic_bezier_to((1.0, 0.0), (1.0, 1.0), [1.0])
    b.cubic_bezier_to((1.0, 2.0), (0.0, 2.0), (0.0, 1.0), [2.0])
    b.end(False)
    flat = O.flatten_paths(PathBatch.from_paths([b.build()]), 0.01).path(0)
    assert set(flat.verbs.tolist()) == {LINE_TO, BEGIN, END}
    prev = -1.0
    n_lines = 0
    for ev in flat.iter_with_attributes():
        if ev[0] == "begin":
            attr = ev[1][1][0]
        elif ev[0] == "line":
            assert ev[1][1][0] == prev
            attr = ev[2][1][0]
            n_lines += 1
        else:
            assert ev[0] == "end" and ev[1][1][0] == prev
            continue
        assert attr > prev
        prev = attr
    assert n_lines > 10 and prev == 2.0
    # the attribute of every flattened point is prev * (1 - t) + attr * t for the accumulated t of for_each_flattened_with_t
    ctrl = np.array([[0, 0, 1, 0, 1, 1]], np.float32)
    _off, _pts, t = O.flatten_quadratics(ctrl, 0.01, with_t=True)
    got = [e[2][1][0] for e in flat.iter_with_attributes() if e[0] == "line"][: len(t)]
    want = [np.float32(1.0) if ti == 1.0 else np.float32(np.float32(0.0) * (np.float32(1.0) - ti) + np.float32(1.0) * ti) for ti in t]
    assert got == want


def test_path_measurements_known_answers():
    """The measurements behind crates/algorithms/src/measure.rs `measure_line` (:718-733), `measure_square` (:735-756: the
    normalized samples 0.125 / 0.375 / ... fall on the edge midpoints, i.e. the closed unit square measures 4 with an edge
    at every unit) and `measure_bezier_curve` (:782-806: half of the length is reached at the joint of two mirrored
    quadratics), plus the edge layout of `PathMeasurements::initialize` (:170-257): every Begin / Line / Close contributes
    one edge with t = 1, an open End none."""
    from lyon_b200.path import Path, PathBatch
    b = Path.builder()
    b.begin((0.0, 0.0))
    b.line_to((1.0, 0.0))
    b.end(False)
    line = b.build()
    b = Path.builder()
    b.begin((0.0, 0.0))
    b.line_to((1.0, 0.0))
    b.line_to((1.0, 1.0))
    b.line_to((0.0, 1.0))
    b.close()
    square = b.build()
    dist, idx, t, begin, length = O.measure_paths(PathBatch.from_paths([line, square]), 0.01)
    assert begin.tolist() == [0, 2, 7]
    assert length.tolist() == [1.0, 4.0]
    assert dist.tolist() == [0.0, 1.0, 0.0, 1.0, 2.0, 3.0, 4.0]
    assert idx.tolist() == [0, 1, 0, 1, 2, 3, 4] and set(t.tolist()) == {1.0}
    # a curve: one edge per flattened segment, t = the accumulated parameter, the last one exactly 1; the length is the
    # polyline's (measure.rs:`measure_bezier_curve` checks the quarter-circle-like quadratic (0,0)-(0.5,0.7)-(1,0) ~ 1.68)
    b = Path.builder()
    b.begin((0.0, 0.0))
    b.quadratic_bezier_to((0.5, 0.7), (1.0, 0.0))
    b.end(False)
    q = PathBatch.from_paths([b.build()])
    dist, idx, t, begin, length = O.measure_paths(q, 0.01)
    off, pts, tt = O.flatten_quadratics(np.array([[0, 0, 0.5, 0.7, 1, 0]], np.float32), 0.01, with_t=True)
    assert len(dist) == 1 + len(pts) and t[1:].tolist() == tt.tolist() and t[-1] == 1.0
    assert np.all(np.diff(dist) > 0) and abs(float(length[0]) - 1.4) < 0.2
    b = Path.builder()
    b.begin((0.0, 0.0))
    b.quadratic_bezier_to((0.5, 0.7), (1.0, 0.0))
    b.quadratic_bezier_to((1.5, -0.7), (2.0, 0.0))
    b.end(False)
    d2, i2, t2, _b2, l2 = O.measure_paths(PathBatch.from_paths([b.build()]), 0.01)
    joint = d2[np.flatnonzero((i2 == 1) & (t2 == 1.0))[0]]
    assert abs(float(joint) / float(l2[0]) - 0.5) < 1e-5
    # tolerance.max(1e-4): a tiny tolerance measures like 1e-4
    assert O.measure_paths(q, 1e-9)[4][0] == O.measure_paths(q, 1e-4)[4][0]
