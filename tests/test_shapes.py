"""Shape helpers: the host-side generators of lyon_b200/shapes.py against the oracle's independent
restatement (byte-identical paths), the reference's own known answers, and -- on the GPU -- strokes of shape batches
through the same kernels against the oracle."""
import ctypes

import numpy as np
import pytest

import oracle_ffi as O
from lyon_b200 import LineCap, LineJoin, Path, PathBatch, StrokeOptions
from lyon_b200 import shapes as S

f = np.float32


def oracle_shape(kind, params, winding=1) -> Path:
    L = O.lib()
    L.oracle_shape_path.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
    p = np.asarray(params, np.float32)
    nv, npt = ctypes.c_uint32(), ctypes.c_uint32()
    L.oracle_shape_path(kind, p.ctypes.data, winding, None, None, ctypes.byref(nv), ctypes.byref(npt))
    verbs = np.zeros(nv.value, np.uint8)
    pts = np.zeros((npt.value, 2), np.float32)
    L.oracle_shape_path(kind, p.ctypes.data, winding, verbs.ctypes.data, pts.ctypes.data, ctypes.byref(nv), ctypes.byref(npt))
    return Path(pts, verbs, 0)


def same_path(a: Path, b: Path):
    assert a.verbs.tolist() == b.verbs.tolist()
    assert a.points.astype(np.float32).tobytes() == b.points.astype(np.float32).tobytes()


@pytest.mark.parametrize("winding", [S.POSITIVE, S.NEGATIVE])
def test_generators_match_the_oracle(winding):
    rng = np.random.default_rng(5)
    for _ in range(50):
        mn = rng.uniform(-100, 100, 2).astype(np.float32)
        mx = (mn + rng.uniform(0.01, 80, 2)).astype(np.float32)
        b = Path.builder()
        S.add_rectangle(b, mn, mx, winding)
        same_path(b.build(), oracle_shape(0, [*mn, *mx], winding))
        c, r = rng.uniform(-50, 50, 2).astype(np.float32), f(rng.uniform(0.1, 40))
        b = Path.builder()
        S.add_circle(b, c, r, winding)
        same_path(b.build(), oracle_shape(1, [*c, r], winding))
        radii, rot = rng.uniform(0.5, 30, 2).astype(np.float32), f(rng.uniform(-3, 3))
        b = Path.builder()
        S.add_ellipse(b, c, radii, rot, winding)
        same_path(b.build(), oracle_shape(2, [*c, *radii, rot], winding))
        rr = rng.uniform(0, 60, 4).astype(np.float32) * (rng.uniform(size=4) > 0.25)
        b = Path.builder()
        S.add_rounded_rectangle(b, mn, mx, rr, winding)
        same_path(b.build(), oracle_shape(3, [*mn, *mx, *rr], winding))


def test_reference_known_answers():
    # builder.rs:1807-1822 `top_right_rounded_rect`
    b = Path.builder()
    S.add_rounded_rectangle(b, (0.0, 0.0), (100.0, 100.0), (0.0, 2.0, 0.0, 0.0), S.POSITIVE)
    tr = list(b.build().iter())[2]
    assert tr[0] == "cubic" and tr[1] == (f(98.0), f(0.0)) and tr[4] == (f(100.0), f(2.0))
    # builder.rs:1508-1546: a circle is Begin + four cubics + Close, starting on the left, through top / right / bottom
    b = Path.builder()
    S.add_circle(b, (10.0, 20.0), 5.0, S.POSITIVE)
    ev = list(b.build().iter())
    assert [e[0] for e in ev] == ["begin", "cubic", "cubic", "cubic", "cubic", "end"] and ev[-1][3] is True
    assert [e[-1] for e in ev[1:5]] == [(f(10), f(15)), (f(15), f(20)), (f(10), f(25)), (f(5), f(20))]
    # builder.rs:606-636: an ellipse is the arc 0 .. 2 pi as ceil(2 pi / (pi / 4)) = 8 quadratics, closed
    b = Path.builder()
    S.add_ellipse(b, (0.0, 0.0), (10.0, 5.0), 0.0, S.POSITIVE)
    ev = list(b.build().iter())
    assert [e[0] for e in ev] == ["begin"] + ["quadratic"] * 8 + ["end"]
    assert ev[0][1] == (f(10.0), f(0.0)) and np.allclose(ev[2][3], (0.0, 5.0), atol=1e-5)
    # add_polygon (builder.rs:560-573) and FromPolyline agree
    b = Path.builder()
    S.add_polygon(b, [(0, 0), (5, 0), (5, 5)], True)
    assert [e[0] for e in b.build().iter()] == ["begin", "line", "line", "end"]


def test_thin_rectangle_rule_matches_the_oracle():
    """stroke.rs:523-537 / 2336-2369: with Miter joins a rectangle thinner than the line width becomes a line segment with
    square caps and a wider line; with the other joins only below 5 % of the line width (round joins: round caps)."""
    L = O.lib()
    L.oracle_thin_rectangle.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    rng = np.random.default_rng(9)
    n_thin = 0
    for _ in range(400):
        mn = rng.uniform(-10, 10, 2).astype(np.float32)
        ext = np.where(rng.uniform(size=2) < 0.4, rng.uniform(0, 0.2, 2), rng.uniform(0, 20, 2)).astype(np.float32)
        mx = (mn + ext).astype(np.float32)
        join = LineJoin(int(rng.integers(0, 4)))
        lw = f(rng.uniform(0.1, 3))
        o = StrokeOptions(line_width=float(lw), line_join=join)
        rect = np.array([*mn, *mx], np.float32)
        out = np.zeros(6, np.float32)
        want = L.oracle_thin_rectangle(rect.ctypes.data, lw, int(join), out.ctypes.data)
        got = S.thin_rectangle(mn, mx, o)
        assert (got is not None) == bool(want)
        if got:
            n_thin += 1
            frm, to, o2 = got
            assert [frm[0], frm[1], to[0], to[1], f(o2.line_width)] == out[:5].tolist()
            assert int(o2.start_cap) == int(out[5]) == int(o2.end_cap)
    assert n_thin > 50
    # variable line width never takes the approximation (stroke.rs:531)
    assert S.thin_rectangle((0, 0), (10, 0.01), StrokeOptions(line_width=1.0).with_variable_line_width(0)) is None


@pytest.mark.gpu
def test_shape_strokes_match_the_oracle():
    """tessellate_circle / ellipse / polygon / rectangle(s) and the recording StrokeBuilder through the batch kernels
    against the oracle stroking the oracle's own shape paths (thin rectangles with the oracle's stand-in segment)."""
    import torch
    from lyon_b200 import DeviceTessellator, StrokeTessellator
    from parity import assert_mesh_parity
    tz = DeviceTessellator(0)
    st = StrokeTessellator(tz)
    opts = [StrokeOptions(line_width=1.5, line_join=LineJoin.Round, tolerance=0.02).with_line_cap(LineCap.Round),
            StrokeOptions(line_width=0.7, line_join=LineJoin.MiterClip, tolerance=0.05),
            StrokeOptions(line_width=2.0, line_join=LineJoin.Bevel, tolerance=0.1)]
    for o in opts:
        got = st.tessellate_circle((3.0, 4.0), 10.0, o)
        assert_mesh_parity(got, O.stroke_batch(PathBatch.from_paths([oracle_shape(1, [3, 4, 10])]), o), 20.0, what="circle")
        got = st.tessellate_ellipse((0.0, 0.0), (12.0, 4.0), 0.4, S.NEGATIVE, o)
        assert_mesh_parity(got, O.stroke_batch(PathBatch.from_paths([oracle_shape(2, [0, 0, 12, 4, 0.4], -1)]), o), 24.0, what="ellipse")
        got = st.tessellate_polygon([(0, 0), (8, 1), (4, 9), (-3, 5)], True, o)
        b = Path.builder()
        b.begin((0.0, 0.0)); b.line_to((8.0, 1.0)); b.line_to((4.0, 9.0)); b.line_to((-3.0, 5.0)); b.end(True)
        assert_mesh_parity(got, O.stroke_batch(PathBatch.from_paths([b.build()]), o), 12.0, what="polygon")
    # a batch of rectangles, some of them thin: appended in order to ONE VertexBuffers (absolute indices)
    o = StrokeOptions(line_width=1.0, line_join=LineJoin.Miter, tolerance=0.05)
    rects = [((0, 0), (10, 5)), ((20, 0), (30, 0.4)), ((40, 0), (40.2, 9)), ((50, 1), (58, 7)), ((60, 0), (70, 0.4))]
    got = st.tessellate_rectangles(rects, o)
    L = O.lib()
    L.oracle_thin_rectangle.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    wv, wi = [], []
    base = 0
    for mn, mx in rects:
        rect = np.array([*mn, *mx], np.float32)
        out = np.zeros(6, np.float32)
        if L.oracle_thin_rectangle(rect.ctypes.data, f(1.0), 0, out.ctypes.data):
            b = Path.builder()
            b.begin((out[0], out[1])); b.line_to((out[2], out[3])); b.end(False)
            cap = LineCap(int(out[5]))
            m = O.stroke_batch(PathBatch.from_paths([b.build()]), StrokeOptions(line_width=float(out[4]), line_join=LineJoin.Miter,
                                                                                 tolerance=0.05, start_cap=cap, end_cap=cap))
        else:
            m = O.stroke_batch(PathBatch.from_paths([oracle_shape(0, rect)]), o)
        wv.append(m.vertices)
        wi.append(m.indices + np.uint32(base))
        base += len(m.vertices)
    assert got.vertices.tobytes() == np.concatenate(wv).tobytes()
    assert np.array_equal(got.indices, np.concatenate(wi))
    # the recording builder: two sub-paths under one set of options are one path (ids and advancement run on)
    sb = st.builder(opts[0])
    sb.begin((0.0, 0.0)); sb.line_to((5.0, 0.0)); sb.quadratic_bezier_to((8.0, 3.0), (5.0, 6.0)); sb.end(False)
    sb.add_circle((20.0, 0.0), 4.0)
    got = sb.build()
    b = Path.builder()
    b.begin((0.0, 0.0)); b.line_to((5.0, 0.0)); b.quadratic_bezier_to((8.0, 3.0), (5.0, 6.0)); b.end(False)
    S.add_circle(b, (20.0, 0.0), 4.0)
    assert_mesh_parity(got, O.stroke_batch(PathBatch.from_paths([b.build()]), opts[0]), 30.0, what="builder")
    with pytest.raises(NotImplementedError):
        sb2 = st.builder(opts[0])
        sb2.begin((0.0, 0.0))
        sb2.set_line_join(LineJoin.Bevel)
    torch.cuda.synchronize()
    tz.close()
