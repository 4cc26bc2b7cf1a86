"""The reference's own tests of the host-side path model, restated against `lyon_b200/path.py`.

Each test names the reference test it restates (file:line).  These pin the input side of the boundary: the
SVG path-data parser, `PathBuilder` / `BuilderWithAttributes` storage, `WithSvg`, the event iterators and
`FromPolyline`."""
import numpy as np
import pytest

from lyon_b200.path import (
    ParseError,
    Path,
    PathBuilder,
    from_polyline,
    parse_svg_path,
    svg_arc_to_arc,
    arc_to_quadratics,
)

f = np.float32


def P(x, y):
    return (f(x), f(y))


def events(path):
    return list(path.iter())


# ---- crates/extra/src/parser.rs:496-769 -------------------------------------------------------------------------------


def test_parser_empty():  # parser.rs:496-513
    assert events(parse_svg_path("")) == []
    assert events(parse_svg_path(" ")) == []


def test_parser_simple_square():  # parser.rs:515-526
    p = parse_svg_path("M 0 0 L 1 0 L 1 1 L 0 1 Z")
    assert events(p) == [
        ("begin", P(0, 0)),
        ("line", P(0, 0), P(1, 0)),
        ("line", P(1, 0), P(1, 1)),
        ("line", P(1, 1), P(0, 1)),
        ("end", P(0, 1), P(0, 0), True),
    ]


def test_parser_simple_attr():  # parser.rs:528-539
    p = parse_svg_path("M 0 0 1.0 L 1 0 2.0 L 1 1 3.0 L 0 1 4.0 Z", num_attributes=1)
    ev = list(p.iter_with_attributes())
    assert [e[0] for e in ev] == ["begin", "line", "line", "line", "end"]
    assert ev[0][1] == (P(0, 0), (f(1.0),))
    assert ev[3][2] == (P(0, 1), (f(4.0),))
    assert ev[4][1] == (P(0, 1), (f(4.0),)) and ev[4][2] == (P(0, 0), (f(1.0),)) and ev[4][3] is True


def test_parser_implicit_polyline():  # parser.rs:541-552
    p = parse_svg_path("0 0 0 1 1 1.0 2 2 2.0 3 3 3", num_attributes=1)
    ev = list(p.iter_with_attributes())
    assert [e[0] for e in ev] == ["begin", "line", "line", "line", "end"]
    assert [e[2][1][0] for e in ev[1:4]] == [f(1.0), f(2.0), f(3.0)]
    assert ev[4][3] is False


def test_parser_invalid_cmd():  # parser.rs:554-595
    with pytest.raises(ParseError) as e:
        parse_svg_path("x 0 0 0", num_attributes=1)
    assert e.value == ParseError("command", 0, 0, command="x")
    with pytest.raises(ParseError) as e:
        parse_svg_path("\n M 0 \n0 1 x 1 1 1", num_attributes=1)
    assert e.value == ParseError("command", 2, 4, command="x")


def test_parser_number_01():  # parser.rs:597-630
    p = parse_svg_path("M 0.6.5")
    assert events(p) == [("begin", P(0.6, 0.5)), ("end", P(0.6, 0.5), P(0.6, 0.5), False)]


def test_parser_number_scientific_notation():  # parser.rs:632-643
    p = parse_svg_path("M 1e-2 -1E3")
    assert events(p)[0] == ("begin", P(1e-2, -1e3))


@pytest.mark.parametrize("src", ["M 0 --1", "M 0 1ee2", "M 0 1e--1", "M 0 *2", "M 0 e", "M 0 1e", "M 0 +1"])
def test_parser_bad_numbers(src):  # parser.rs:645-676
    with pytest.raises(ParseError) as e:
        parse_svg_path(src)
    assert e.value.kind == "number"


@pytest.mark.parametrize("src", ["M 0 0 | xxxxxx", "M 0 0| xxxxxx", "| xxxxxx", "    | xxxxxx"])
def test_parser_stop(src):  # parser.rs:678-699
    parse_svg_path(src, stop_at="|")


def test_parser_need_start():  # parser.rs:701-741
    b = PathBuilder(0)
    with pytest.raises(ParseError) as e:
        parse_svg_path("M 0 0 Z L 1 1 2 2 L 3 3 Z M 4 4", stop_at="|", builder=b)
    assert e.value.kind == "missing_move_to"
    p1 = b.build()
    p2 = parse_svg_path("M 0 0 Z")
    assert events(p1) == events(p2)


@pytest.mark.parametrize(
    "src", ["M 1e-9 0", "M -1e-9 0", "M -1e11 0", "M 1.e-9 1.4e-4z", "M 1.6e-9 1.4e-4 z", "M0 1.6e-9L0 1.4e-4"]
)
def test_parser_issue_895(src):  # parser.rs:743-760
    parse_svg_path(src)


def test_parser_issue_919():  # parser.rs:762-769: must not panic; the result is ignored by the reference
    try:
        parse_svg_path("0 0M")
    except ParseError:
        pass


def test_parser_arc_is_quadratics():
    """parser.rs:269-306: an arc command reaches the builder as `quadratic_bezier_to` calls that end exactly on the arc
    samples; a half circle needs ceil(pi / (pi/4)) = 4 of them and ends (within f32 rounding of the ellipse sample) on `to`."""
    p = parse_svg_path("M 0 0 A 50 50 0 0 1 100 0")
    ev = events(p)
    assert [e[0] for e in ev] == ["begin"] + ["quadratic"] * 4 + ["end"]
    assert np.allclose(ev[4][3], (100.0, 0.0), atol=1e-4)
    assert np.allclose(ev[2][3], (50.0, -50.0), atol=1e-4)  # sweep flag set: through the top (negative y)
    # straight-line arcs degrade to line_to (arc.rs:763-767)
    assert [e[0] for e in events(parse_svg_path("M 0 0 A 0 50 0 0 1 100 0"))] == ["begin", "line", "end"]
    # flags are single characters (parser.rs:469-490)
    assert len(events(parse_svg_path("M0 0a50 50 0 01100 0"))) == 6
    with pytest.raises(ParseError) as e:
        parse_svg_path("M 0 0 A 50 50 0 2 1 100 0")
    assert e.value.kind == "flag"


def test_parser_arc_attributes_interpolated():
    """parser.rs:283-300: the attributes of the quadratics of an arc are `prev * (1 - t) + attr * t` with t the END of each
    piece's range."""
    p = parse_svg_path("M 0 0 10 A 50 50 0 0 1 100 0 20", num_attributes=1)
    ev = list(p.iter_with_attributes())
    assert [e[0] for e in ev] == ["begin"] + ["quadratic"] * 4 + ["end"]
    got = [e[3][1][0] for e in ev[1:5]]
    t = [f(0.25), f(0.5), f(0.75), f(1.0)]
    assert got == [f(f(f(10) * f(f(1) - ti)) + f(f(20) * ti)) for ti in t]


# ---- crates/path/src/path.rs:1762-2230 --------------------------------------------------------------------------------


def test_path_builder_1():  # path.rs:1762-1955
    p = Path.builder_with_attributes(1)
    p.begin(P(0, 0), [0.0])
    p.line_to(P(1, 0), [1.0])
    p.line_to(P(2, 0), [2.0])
    p.line_to(P(3, 0), [3.0])
    p.quadratic_bezier_to(P(4, 0), P(4, 1), [4.0])
    p.cubic_bezier_to(P(5, 0), P(5, 1), P(5, 2), [5.0])
    p.end(True)
    p.begin(P(10, 0), [6.0])
    p.line_to(P(11, 0), [7.0])
    p.line_to(P(12, 0), [8.0])
    p.line_to(P(13, 0), [9.0])
    p.quadratic_bezier_to(P(14, 0), P(14, 1), [10.0])
    p.cubic_bezier_to(P(15, 0), P(15, 1), P(15, 2), [11.0])
    p.end(True)
    p.begin(P(1, 1), [12.0])
    p.end(False)
    p.begin(P(2, 2), [13.0])
    p.end(False)
    p.begin(P(3, 3), [14.0])
    p.line_to(P(4, 4), [15.0])
    p.end(False)
    path = p.build()

    def E(x, y, a):
        return (P(x, y), (f(a),))

    assert list(path.iter_with_attributes()) == [
        ("begin", E(0, 0, 0)),
        ("line", E(0, 0, 0), E(1, 0, 1)),
        ("line", E(1, 0, 1), E(2, 0, 2)),
        ("line", E(2, 0, 2), E(3, 0, 3)),
        ("quadratic", E(3, 0, 3), P(4, 0), E(4, 1, 4)),
        ("cubic", E(4, 1, 4), P(5, 0), P(5, 1), E(5, 2, 5)),
        ("end", E(5, 2, 5), E(0, 0, 0), True),
        ("begin", E(10, 0, 6)),
        ("line", E(10, 0, 6), E(11, 0, 7)),
        ("line", E(11, 0, 7), E(12, 0, 8)),
        ("line", E(12, 0, 8), E(13, 0, 9)),
        ("quadratic", E(13, 0, 9), P(14, 0), E(14, 1, 10)),
        ("cubic", E(14, 1, 10), P(15, 0), P(15, 1), E(15, 2, 11)),
        ("end", E(15, 2, 11), E(10, 0, 6), True),
        ("begin", E(1, 1, 12)),
        ("end", E(1, 1, 12), E(1, 1, 12), False),
        ("begin", E(2, 2, 13)),
        ("end", E(2, 2, 13), E(2, 2, 13), False),
        ("begin", E(3, 3, 14)),
        ("line", E(3, 3, 14), E(4, 4, 15)),
        ("end", E(4, 4, 15), E(3, 3, 14), False),
    ]


def test_path_builder_empty():  # path.rs:1957-1963
    assert events(Path.builder_with_attributes(5).build()) == []


def test_path_builder_empty_begin():  # path.rs:1965-2021
    p = Path.builder_with_attributes(1)
    p.begin(P(1, 2), [0.0])
    p.end(False)
    p.begin(P(3, 4), [1.0])
    p.end(False)
    p.begin(P(5, 6), [2.0])
    p.end(False)
    assert events(p.build()) == [
        ("begin", P(1, 2)),
        ("end", P(1, 2), P(1, 2), False),
        ("begin", P(3, 4)),
        ("end", P(3, 4), P(3, 4), False),
        ("begin", P(5, 6)),
        ("end", P(5, 6), P(5, 6), False),
    ]


def test_extend_from_paths():  # path.rs:2023-2103
    b = Path.builder()
    b.begin(P(0, 0))
    b.line_to(P(5, 0))
    b.line_to(P(5, 5))
    b.end(True)
    path1 = b.build()
    b = Path.builder()
    b.begin(P(1, 1))
    b.line_to(P(4, 0))
    b.line_to(P(4, 4))
    b.end(True)
    path2 = b.build()
    b = Path.builder()
    b.extend_from_paths([path1, path2])
    assert events(b.build()) == [
        ("begin", P(0, 0)),
        ("line", P(0, 0), P(5, 0)),
        ("line", P(5, 0), P(5, 5)),
        ("end", P(5, 5), P(0, 0), True),
        ("begin", P(1, 1)),
        ("line", P(1, 1), P(4, 0)),
        ("line", P(4, 0), P(4, 4)),
        ("end", P(4, 4), P(1, 1), True),
    ]


def test_first_last():  # path.rs:2144-2171
    for close, last in ((False, (P(4, 4), (f(5.0),))), (True, (P(0, 0), (f(1.0),)))):
        b = Path.builder_with_attributes(1)
        b.begin(P(0, 0), [1.0])
        b.line_to(P(2, 2), [3.0])
        b.line_to(P(4, 4), [5.0])
        b.end(close)
        path = b.build()
        assert path.first_endpoint() == (P(0, 0), (f(1.0),))
        assert path.last_endpoint() == last
    assert Path.builder().build().first_endpoint() is None


def test_id_events():  # path.rs:2173-2230
    b = Path.builder_with_attributes(1)
    e1 = b.begin(P(0, 0), [1.0])
    e2 = b.line_to(P(2, 2), [3.0])
    e3 = b.line_to(P(4, 4), [5.0])
    b.end(False)
    e4 = b.begin(P(6, 6), [7.0])
    e5 = b.line_to(P(8, 8), [9.0])
    e6 = b.line_to(P(10, 10), [11.0])
    b.end(True)
    e7 = b.begin(P(12, 12), [13.0])
    e8 = b.line_to(P(14, 14), [15.0])
    b.end(False)
    assert list(b.build().id_iter()) == [
        ("begin", e1),
        ("line", e1, e2),
        ("line", e2, e3),
        ("end", e3, e1, False),
        ("begin", e4),
        ("line", e4, e5),
        ("line", e5, e6),
        ("end", e6, e4, True),
        ("begin", e7),
        ("line", e7, e8),
        ("end", e8, e7, False),
    ]


# ---- crates/path/src/builder.rs:1670-1806 -----------------------------------------------------------------------------


def test_svg_builder_line_to_after_close():  # builder.rs:1670-1718
    p = Path.builder().with_svg()
    p.line_to(P(1, 0))
    p.close()
    p.line_to(P(2, 0))
    assert events(p.build()) == [
        ("begin", P(1, 0)),
        ("end", P(1, 0), P(1, 0), True),
        ("begin", P(1, 0)),
        ("line", P(1, 0), P(2, 0)),
        ("end", P(2, 0), P(1, 0), False),
    ]


def test_svg_builder_relative_curves():  # builder.rs:1720-1762
    p = Path.builder().with_svg()
    p.move_to(P(0, 0))
    p.relative_quadratic_bezier_to((0.0, 100.0), (-100.0, 100.0))
    p.relative_line_to((-50.0, 0.0))
    assert events(p.build()) == [
        ("begin", P(0, 0)),
        ("quadratic", P(0, 0), P(0, 100), P(-100, 100)),
        ("line", P(-100, 100), P(-150, 100)),
        ("end", P(-150, 100), P(0, 0), False),
    ]


def test_svg_builder_arc_to_update_position():  # builder.rs:1764-1778
    p = Path.builder().with_svg()
    p.move_to(P(0, 0))
    assert p.current_position == P(0, 0)
    p.arc_to((100.0, 100.0), 0.0, False, False, P(0, 100))
    assert p.current_position != P(0, 0)
    assert np.allclose(p.current_position, (0.0, 100.0), atol=0.05)  # fast_atan2 is good to 2.5e-4 rad x radius 100


def test_svg_builder_issue_650():  # builder.rs:1780-1790
    b = Path.builder().with_svg()
    b.arc(P(0, 0), (50.0, 50.0), np.pi, 0.0)  # current position == center: skipped, no NaN
    assert events(b.build()) == []


def test_svg_builder_straight_line_arc():  # builder.rs:1792-1805
    p = Path.builder().with_svg()
    p.move_to(P(100, 0))
    p.arc_to((100.0, 100.0), 0.0, False, False, P(100, 0))
    assert [e[0] for e in events(p.build())] == ["begin", "line", "end"]


def test_svg_arc_to_arc_round_trip():
    """arc.rs:58-148: the centre form of an SVG arc samples `from` at t = 0 and `to` at t = 1 (the property the reference's
    `from_svg_arc` tests check, arc.rs `test_from_svg_arc`)."""
    frm, to = P(10, 20), P(60, -5)
    for large in (False, True):
        for sweep in (False, True):
            c, r, start, sw, xr = svg_arc_to_arc(frm, to, (40.0, 30.0), f(0.3), large, sweep)
            quads = arc_to_quadratics(c, r, start, sw, xr)
            assert quads and quads[-1][2] == f(1.0)
            assert np.allclose(quads[-1][1], to, atol=0.03)  # fast_atan2: 2.5e-4 rad x radius 40
            assert (abs(sw) > np.pi) == large
            assert (sw > 0) == sweep


# ---- crates/path/src/iterator.rs:340-430 ------------------------------------------------------------------------------


def test_from_polyline_open():  # iterator.rs:340-387
    pts = [P(1, 1), P(3, 1), P(4, 5), P(5, 2)]
    assert list(from_polyline(pts, False)) == [
        ("begin", P(1, 1)),
        ("line", P(1, 1), P(3, 1)),
        ("line", P(3, 1), P(4, 5)),
        ("line", P(4, 5), P(5, 2)),
        ("end", P(5, 2), P(1, 1), False),
    ]


def test_from_polyline_closed():  # iterator.rs:389-436
    pts = [P(1, 1), P(3, 1), P(4, 5), P(5, 2)]
    ev = list(from_polyline(pts, True))
    assert ev[-1] == ("end", P(5, 2), P(1, 1), True) and len(ev) == 5
