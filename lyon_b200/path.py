"""Host-side path model mirroring lyon_path (reference: crates/path/src/path.rs, builder.rs).

A `Path` is two flat buffers exactly like the reference's (`path.rs:75-79`):
  * `verbs`  u8, one per command: LineTo=0, QuadraticTo=1, CubicTo=2, Begin=3, Close=4, End=5
    (`path.rs:29-36`);
  * `points` f32 (n, 2): endpoints, control points (before their endpoint) and, after every
    endpoint, ceil(A/2) slots holding the A custom attributes (`path.rs:516-611`, `:650-665`).
    `Close` stores the first endpoint (and its attributes) again (`path.rs:537-539`).

`PathBatch` concatenates many paths into one pair of buffers plus (N+1) range arrays, the layout
`lyon_cuda_stroke` uploads (the reference's precedent is `PathBuffer`, `path_buffer.rs:15-27`).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

import math

import numpy as np

LINE_TO, QUADRATIC_TO, CUBIC_TO, BEGIN, CLOSE, END = 0, 1, 2, 3, 4, 5

_f32 = np.float32


def _pt(p) -> tuple:
    return (_f32(p[0]), _f32(p[1]))


@dataclass
class Path:
    points: np.ndarray  # (n, 2) float32
    verbs: np.ndarray  # (m,) uint8
    num_attributes: int = 0

    @staticmethod
    def builder() -> "PathBuilder":
        return PathBuilder(0)

    @staticmethod
    def builder_with_attributes(num_attributes: int) -> "PathBuilder":
        return PathBuilder(num_attributes)

    def iter(self):
        """Yield PathEvents like `Path::iter` (`path.rs:817-909`) as tuples:
        ("begin", at) / ("line", from, to) / ("quadratic", from, ctrl, to) /
        ("cubic", from, c1, c2, to) / ("end", last, first, close)."""
        stride = (self.num_attributes + 1) // 2
        pts = self.points
        cur = 0
        current = first = (0.0, 0.0)
        for v in self.verbs:
            if v == BEGIN:
                current = first = tuple(pts[cur])
                cur += 1 + stride
                yield ("begin", current)
            elif v == LINE_TO:
                frm, current = current, tuple(pts[cur])
                cur += 1 + stride
                yield ("line", frm, current)
            elif v == QUADRATIC_TO:
                frm, ctrl, current = current, tuple(pts[cur]), tuple(pts[cur + 1])
                cur += 2 + stride
                yield ("quadratic", frm, ctrl, current)
            elif v == CUBIC_TO:
                frm, c1, c2, current = current, tuple(pts[cur]), tuple(pts[cur + 1]), tuple(pts[cur + 2])
                cur += 3 + stride
                yield ("cubic", frm, c1, c2, current)
            elif v == CLOSE:
                last = current
                cur += 1 + stride
                yield ("end", last, first, True)
            elif v == END:
                last, current = current, first
                yield ("end", last, first, False)


    def _attrs_at(self, point_index: int):
        """Custom attributes stored after endpoint `point_index` (`path.rs:650-665`: two per point slot)."""
        n = self.num_attributes
        if n == 0:
            return ()
        stride = (n + 1) // 2
        return tuple(self.points[point_index + 1 : point_index + 1 + stride].reshape(-1)[:n])

    def id_iter(self):
        """`Path::id_iter` (`path.rs:960-1040`): the same events as `iter`, endpoints / control points as indices into
        `points` (EndpointId / ControlPointId both index the one point buffer, `path.rs:516-611`)."""
        stride = (self.num_attributes + 1) // 2
        cur = 0
        current = first = 0
        for v in self.verbs:
            if v == BEGIN:
                current = first = cur
                cur += 1 + stride
                yield ("begin", current)
            elif v == LINE_TO:
                frm, current = current, cur
                cur += 1 + stride
                yield ("line", frm, current)
            elif v == QUADRATIC_TO:
                frm, ctrl, current = current, cur, cur + 1
                cur += 2 + stride
                yield ("quadratic", frm, ctrl, current)
            elif v == CUBIC_TO:
                frm, c1, c2, current = current, cur, cur + 1, cur + 2
                cur += 3 + stride
                yield ("cubic", frm, c1, c2, current)
            elif v == CLOSE:
                last = current
                cur += 1 + stride
                yield ("end", last, first, True)
            elif v == END:
                last, current = current, first
                yield ("end", last, first, False)

    def iter_with_attributes(self):
        """`Path::iter_with_attributes` (`path.rs:911-958`): endpoints are (position, attributes) pairs, control points
        plain positions."""
        pts = self.points

        def ep(i):
            return (tuple(pts[i]), self._attrs_at(i))

        for ev in self.id_iter():
            k = ev[0]
            if k == "begin":
                yield ("begin", ep(ev[1]))
            elif k == "line":
                yield ("line", ep(ev[1]), ep(ev[2]))
            elif k == "quadratic":
                yield ("quadratic", ep(ev[1]), tuple(pts[ev[2]]), ep(ev[3]))
            elif k == "cubic":
                yield ("cubic", ep(ev[1]), tuple(pts[ev[2]]), tuple(pts[ev[3]]), ep(ev[4]))
            else:
                yield ("end", ep(ev[1]), ep(ev[2]), ev[3])

    def first_endpoint(self):
        """`PathSlice::first_endpoint` (`path.rs:344-346` region): None for an empty path."""
        if self.verbs.shape[0] == 0:
            return None
        return (tuple(self.points[0]), self._attrs_at(0))

    def last_endpoint(self):
        """`PathSlice::last_endpoint`: the last stored endpoint (a `Close` stores the first endpoint again)."""
        if self.verbs.shape[0] == 0:
            return None
        stride = (self.num_attributes + 1) // 2
        i = self.points.shape[0] - 1 - stride
        return (tuple(self.points[i]), self._attrs_at(i))


def from_polyline(points, closed: bool):
    """`FromPolyline::open / closed` (crates/path/src/iterator.rs:270-338): Begin, one Line per further point, End."""
    it = iter(points)
    try:
        first = current = _pt(next(it))
    except StopIteration:
        return
    yield ("begin", first)
    for p in it:
        nxt = _pt(p)
        yield ("line", current, nxt)
        current = nxt
    yield ("end", current, first, bool(closed))


class PathBuilder:
    """`PathBuilder` for `Path` (`path.rs:516-611` without attributes, `:650-807` with)."""

    def __init__(self, num_attributes: int = 0):
        self.num_attributes = int(num_attributes)
        self._pts: List[tuple] = []
        self._verbs: List[int] = []
        self._first = (_f32(0), _f32(0))
        self._first_attrs: Sequence[float] = ()

    def _push_attrs(self, attributes):
        a = [_f32(x) for x in (attributes or ())]
        assert len(a) == self.num_attributes, "wrong number of custom attributes"
        for i in range(0, len(a), 2):
            self._pts.append((a[i], a[i + 1] if i + 1 < len(a) else _f32(0)))

    def begin(self, at, attributes=None) -> int:
        eid = len(self._pts)
        self._first = _pt(at)
        self._first_attrs = list(attributes or ())
        self._pts.append(self._first)
        self._push_attrs(attributes)
        self._verbs.append(BEGIN)
        return eid

    def end(self, close: bool):
        if close:
            self._pts.append(self._first)
            self._push_attrs(self._first_attrs)
        self._verbs.append(CLOSE if close else END)

    def close(self):
        self.end(True)

    def line_to(self, to, attributes=None) -> int:
        eid = len(self._pts)
        self._pts.append(_pt(to))
        self._push_attrs(attributes)
        self._verbs.append(LINE_TO)
        return eid

    def quadratic_bezier_to(self, ctrl, to, attributes=None) -> int:
        self._pts.append(_pt(ctrl))
        eid = len(self._pts)
        self._pts.append(_pt(to))
        self._push_attrs(attributes)
        self._verbs.append(QUADRATIC_TO)
        return eid

    def cubic_bezier_to(self, ctrl1, ctrl2, to, attributes=None) -> int:
        self._pts.append(_pt(ctrl1))
        self._pts.append(_pt(ctrl2))
        eid = len(self._pts)
        self._pts.append(_pt(to))
        self._push_attrs(attributes)
        self._verbs.append(CUBIC_TO)
        return eid

    def extend_from_paths(self, paths):
        """`BuilderImpl::extend_from_paths` (`path.rs:468-493` region): append whole paths (same attribute count)."""
        for p in paths:
            assert p.num_attributes == self.num_attributes
            self._pts.extend((_f32(x), _f32(y)) for x, y in p.points.reshape(-1, 2))
            self._verbs.extend(int(v) for v in p.verbs)

    def with_svg(self) -> "SvgPathBuilder":
        return SvgPathBuilder(self)

    def build(self) -> Path:
        pts = np.array(self._pts, dtype=np.float32).reshape(-1, 2)
        return Path(pts, np.array(self._verbs, dtype=np.uint8), self.num_attributes)


class SvgPathBuilder:
    """`WithSvg<Builder>` (`builder.rs:1144-1506`): SVG-style commands (implicit move_to, relative and
    smooth variants) translated to begin/end pairs.  All coordinate arithmetic is f32."""

    def __init__(self, builder: PathBuilder):
        self.builder = builder
        z = (_f32(0), _f32(0))
        self.first_position = z
        self.current_position = z
        self.last_ctrl = z
        self.last_cmd = END
        self.need_moveto = True
        self.is_empty = True
        self.attribute_buffer = [0.0] * builder.num_attributes

    def _end_if_needed(self):
        if self.last_cmd <= BEGIN:  # builder.rs:1330-1334
            self.builder.end(False)

    def _begin_if_needed(self, default) -> bool:
        """Returns True when the command must be skipped (it became the move_to)."""
        if self.need_moveto:
            if self.is_empty:
                self.move_to(default)
                return True
            self.move_to(self.first_position)
        return False

    def move_to(self, to):
        self._end_if_needed()
        to = _pt(to)
        self.builder.begin(to, self.attribute_buffer)
        self.is_empty = False
        self.need_moveto = False
        self.first_position = to
        self.current_position = to
        self.last_cmd = BEGIN

    def line_to(self, to):
        to = _pt(to)
        if self._begin_if_needed(to):
            return
        self.current_position = to
        self.last_cmd = LINE_TO
        self.builder.line_to(to, self.attribute_buffer)

    def close(self):
        if self.need_moveto:
            return
        self.current_position = self.first_position
        self.need_moveto = True
        self.last_cmd = CLOSE
        self.builder.close()

    def quadratic_bezier_to(self, ctrl, to):
        ctrl, to = _pt(ctrl), _pt(to)
        if self._begin_if_needed(to):
            return
        self.current_position = to
        self.last_cmd = QUADRATIC_TO
        self.last_ctrl = ctrl
        self.builder.quadratic_bezier_to(ctrl, to, self.attribute_buffer)

    def cubic_bezier_to(self, ctrl1, ctrl2, to):
        ctrl1, ctrl2, to = _pt(ctrl1), _pt(ctrl2), _pt(to)
        if self._begin_if_needed(to):
            return
        self.current_position = to
        self.last_cmd = CUBIC_TO
        self.last_ctrl = ctrl2
        self.builder.cubic_bezier_to(ctrl1, ctrl2, to, self.attribute_buffer)

    def _rel(self, v):
        return (self.current_position[0] + _f32(v[0]), self.current_position[1] + _f32(v[1]))

    def _smooth(self, verb):
        if self.last_cmd == verb:
            c, l = self.current_position, self.last_ctrl
            return (c[0] + (c[0] - l[0]), c[1] + (c[1] - l[1]))
        return self.current_position

    def relative_move_to(self, v):
        self.move_to(self._rel(v))

    def relative_line_to(self, v):
        self.line_to(self._rel(v))

    def relative_quadratic_bezier_to(self, ctrl, to):
        self.quadratic_bezier_to(self._rel(ctrl), self._rel(to))

    def relative_cubic_bezier_to(self, c1, c2, to):
        self.cubic_bezier_to(self._rel(c1), self._rel(c2), self._rel(to))

    def smooth_cubic_bezier_to(self, ctrl2, to):
        self.cubic_bezier_to(self._smooth(CUBIC_TO), ctrl2, to)

    def smooth_relative_cubic_bezier_to(self, ctrl2, to):
        c1 = self._smooth(CUBIC_TO)
        self.cubic_bezier_to(c1, self._rel(ctrl2), self._rel(to))

    def smooth_quadratic_bezier_to(self, to):
        self.quadratic_bezier_to(self._smooth(QUADRATIC_TO), to)

    def smooth_relative_quadratic_bezier_to(self, to):
        c = self._smooth(QUADRATIC_TO)
        self.quadratic_bezier_to(c, self._rel(to))

    def horizontal_line_to(self, x):
        self.line_to((_f32(x), self.current_position[1]))

    def relative_horizontal_line_to(self, dx):
        self.line_to((self.current_position[0] + _f32(dx), self.current_position[1]))

    def vertical_line_to(self, y):
        self.line_to((self.current_position[0], _f32(y)))

    def relative_vertical_line_to(self, dy):
        self.line_to((self.current_position[0], self.current_position[1] + _f32(dy)))

    def arc(self, center, radii, sweep_angle, x_rotation):
        """`WithSvg::arc` (`builder.rs:1272-1312`): angles in radians; the arc becomes quadratics."""
        center = _pt(center)
        self.last_ctrl = self.current_position
        eps = np.finfo(np.float32).eps  # euclid ApproxEq for f32 points: |dx| < eps and |dy| < eps
        if abs(self.current_position[0] - center[0]) < eps and abs(self.current_position[1] - center[1]) < eps:
            return
        d = (_f32(self.current_position[0] - center[0]), _f32(self.current_position[1] - center[1]))
        start_angle = _f32(_fast_atan2(d[1], d[0]) - _f32(x_rotation))
        arc_start = arc_sample(center, radii, x_rotation, start_angle)
        if self.need_moveto:
            self.move_to(arc_start)
        else:
            dx, dy = _f32(arc_start[0] - self.current_position[0]), _f32(arc_start[1] - self.current_position[1])
            if _f32(_f32(dx * dx) + _f32(dy * dy)) > _f32(0.01):
                self.builder.line_to(arc_start, self.attribute_buffer)
        for ctrl, to, _t in arc_to_quadratics(center, radii, start_angle, sweep_angle, x_rotation):
            self.builder.quadratic_bezier_to(ctrl, to, self.attribute_buffer)
            self.current_position = to

    def arc_to(self, radii, x_rotation, large_arc: bool, sweep: bool, to):
        """`SvgPathBuilder::arc_to` for `WithSvg` (`builder.rs:1477-1495`); `x_rotation` in radians, default flags are
        both False (`ArcFlags::default`)."""
        to = _pt(to)
        if svg_arc_is_straight_line(self.current_position, to, radii):
            self.line_to(to)
        else:
            c, r, _start, sw, xr = svg_arc_to_arc(self.current_position, to, radii, x_rotation, large_arc, sweep)
            self.arc(c, r, sw, xr)

    def relative_arc_to(self, radii, x_rotation, large_arc: bool, sweep: bool, to):
        self.arc_to(radii, x_rotation, large_arc, sweep, self._rel(to))

    def build(self) -> Path:
        self._end_if_needed()
        return self.builder.build()


class ParseError(ValueError):
    """`ParseError` of crates/extra/src/parser.rs:13-33: `kind` is "number", "flag", "command" or "missing_move_to"; `line`
    and `column` are counted as `Source` does (parser.rs:51-118: a newline bumps the line and resets the column to -1)."""

    def __init__(self, kind: str, line: int, column: int, src=None, command=None):
        self.kind, self.line, self.column, self.src, self.command = kind, line, column, src, command
        what = {"number": f"Expected number, got {src!r}.", "flag": f"Expected flag (0/1), got {src!r}.",
                "command": f"Invalid command {command!r}.", "missing_move_to": f"Expected move-to command, got {command!r}."}[kind]
        super().__init__(f"Line {line} Column {column}: {what}")

    def __eq__(self, other):
        return isinstance(other, ParseError) and (self.kind, self.line, self.column, self.src, self.command) == (
            other.kind, other.line, other.column, other.src, other.command)

    __hash__ = ValueError.__hash__


class _Source:
    """`Source` (parser.rs:51-118): one character of look-ahead, line / column bookkeeping."""

    def __init__(self, src: str, line: int = 0, column: int = 0):
        self.it = iter(src)
        first = next(self.it, None)
        self.finished = first is None
        self.current = " " if first is None else first
        self.line = line + (1 if self.current == "\n" else 0)
        self.col = column

    def advance_one(self):
        if self.finished:
            return
        c = next(self.it, None)
        if c == "\n":
            self.current = "\n"
            self.line += 1
            self.col = -1
        elif c is not None:
            self.current = c
            self.col += 1
        else:
            self.current = "~"
            self.finished = True

    def skip_whitespace(self):
        while not self.finished and (self.current.isspace() or self.current == ","):
            self.advance_one()


def _parse_f32(text: str):
    """Rust `str::parse::<f32>` for what parse_number can collect ([-]digits[.digits][e[-]digits]): None when Rust fails
    (empty, lone sign / dot, missing exponent digits, a second sign ...)."""
    i, n = 0, len(text)
    if i < n and text[i] == "-":
        i += 1
    d0 = i
    while i < n and text[i].isdigit():
        i += 1
    digits = i - d0
    if i < n and text[i] == ".":
        i += 1
        f0 = i
        while i < n and text[i].isdigit():
            i += 1
        digits += i - f0
    if digits == 0:
        return None
    if i < n and text[i] in "eE":
        i += 1
        if i < n and text[i] == "-":
            i += 1
        e0 = i
        while i < n and text[i].isdigit():
            i += 1
        if i == e0:
            return None
    if i != n:
        return None
    try:
        # decimal -> f64 -> f32 equals Rust's correctly rounded decimal -> f32 except on double-rounding ties, which path
        # data written with a handful of digits does not hit
        with np.errstate(over="ignore"):
            return _f32(float(text))
    except ValueError:
        return None


def _cosf(a):
    return _f32(math.cos(float(a)))


def _sinf(a):
    return _f32(math.sin(float(a)))


def _fast_atan2(y, x):
    """euclid `Trig::fast_atan2` for f32 (euclid 0.22.9 src/trig.rs; the same polynomial as lyon_b200/csrc/device_math.cuh)."""
    f = _f32
    pi = f(math.pi)
    xa, ya = f(abs(x)), f(abs(y))
    with np.errstate(invalid="ignore", divide="ignore"):
        a = f(min(xa, ya) / max(xa, ya))
    s2 = f(a * a)
    r = f(f(f(f(f(f(f(-0.0464964749) * s2) + f(0.15931422)) * s2) - f(0.327622764)) * s2) * a)
    r = f(r + a)
    if ya > xa:
        r = f(f(1.57079637050628662) - r)
    if x < 0:
        r = f(pi - r)
    if y < 0:
        r = f(-r)
    return r


def svg_arc_to_arc(frm, to, radii, x_rotation_rad, large_arc: bool, sweep: bool):
    """`Arc::from_svg_arc` (crates/geom/src/arc.rs:58-148) in f32: (center, radii, start_angle, sweep_angle,
    x_rotation)."""
    f = _f32
    two, pi = f(2.0), f(math.pi)
    rx, ry = f(abs(radii[0])), f(abs(radii[1]))
    xr_rad = f(x_rotation_rad)
    xr = f(np.fmod(xr_rad, f(two * pi)))
    cos_phi, sin_phi = _cosf(xr), _sinf(xr)
    fx, fy, tx, ty = f(frm[0]), f(frm[1]), f(to[0]), f(to[1])
    hd_x, hd_y = f(f(fx - tx) / two), f(f(fy - ty) / two)
    hs_x, hs_y = f(f(fx + tx) / two), f(f(fy + ty) / two)
    px = f(f(cos_phi * hd_x) + f(sin_phi * hd_y))
    py = f(f(f(-sin_phi) * hd_x) + f(cos_phi * hd_y))
    rf = f(f(f(px * px) / f(rx * rx)) + f(f(py * py) / f(ry * ry)))
    if rf > f(1.0):
        scale = f(np.sqrt(rf))
        rx, ry = f(rx * scale), f(ry * scale)
    rxry, rxpy, rypx = f(rx * ry), f(rx * py), f(ry * px)
    sum_of_sq = f(f(rxpy * rxpy) + f(rypx * rypx))
    sign_coe = f(-1.0) if large_arc == sweep else f(1.0)
    coe = f(sign_coe * f(np.sqrt(f(abs(f(f(f(rxry * rxry) - sum_of_sq) / sum_of_sq))))))
    tcx = f(f(coe * rxpy) / ry)
    tcy = f(f(f(-coe) * rypx) / rx)
    cx = f(f(f(cos_phi * tcx) - f(sin_phi * tcy)) + hs_x)
    cy = f(f(f(sin_phi * tcx) + f(cos_phi * tcy)) + hs_y)
    svx, svy = f(f(px - tcx) / rx), f(f(py - tcy) / ry)
    evx, evy = f(f(f(-px) - tcx) / rx), f(f(f(-py) - tcy) / ry)
    two_pi = f(two * pi)
    start_angle = _fast_atan2(svy, svx)
    sweep_angle = f(np.fmod(f(_fast_atan2(evy, evx) - start_angle), two_pi))
    if sweep and sweep_angle < 0:
        sweep_angle = f(sweep_angle + two_pi)
    elif not sweep and sweep_angle > 0:
        sweep_angle = f(sweep_angle - two_pi)
    return (cx, cy), (rx, ry), start_angle, sweep_angle, xr_rad


def arc_sample(center, radii, x_rotation_rad, angle):
    """`Arc::sample` via `sample_ellipse` (arc.rs:705-712 region): center + rotate((rx cos a, ry sin a), x_rotation)."""
    f = _f32
    xcos, xsin = _cosf(x_rotation_rad), _sinf(x_rotation_rad)
    ex, ey = f(f(radii[0]) * _cosf(angle)), f(f(radii[1]) * _sinf(angle))
    return (f(f(center[0]) + f(f(ex * xcos) - f(ey * xsin))), f(f(center[1]) + f(f(ey * xcos) + f(ex * xsin))))


def arc_to_quadratics(center, radii, start_angle, sweep_angle, x_rotation_rad):
    """`Arc::for_each_quadratic_bezier_with_t` (crates/geom/src/arc.rs:668-703, `arc_to_quadratic_beziers_with_t`) in
    f32: the list of (ctrl, to, t_end).  cos / sin / tan are evaluated in double and rounded once (within one ulp of the
    platform f32 routines lyon calls, tests/test_libm_pinning.py)."""
    f = _f32
    two, pi = f(2.0), f(math.pi)
    two_pi = f(two * pi)
    rx, ry = f(radii[0]), f(radii[1])
    cx, cy = f(center[0]), f(center[1])
    start_angle, sweep_angle, xr_rad = f(start_angle), f(sweep_angle), f(x_rotation_rad)
    sign = f(np.copysign(f(1.0), sweep_angle))
    sa = f(min(f(abs(sweep_angle)), two_pi))
    with np.errstate(invalid="ignore", divide="ignore"):
        n_steps = f(np.ceil(f(sa / f(pi / f(4.0)))))
        step = f(f(sa / n_steps) * sign)
        dt = f(f(1.0) / n_steps)
    n = int(n_steps) if np.isfinite(n_steps) else 0
    if n == 0:
        return []
    alpha = f(math.tan(float(f(step * f(0.5)))))
    xcos, xsin = _cosf(xr_rad), _sinf(xr_rad)  # Rotation::new(x_rotation)

    def ellipse(angle):  # sample_ellipse: rotate (rx cos a, ry sin a) by x_rotation
        ex, ey = f(rx * _cosf(angle)), f(ry * _sinf(angle))
        return f(f(ex * xcos) - f(ey * xsin)), f(f(ey * xcos) + f(ex * xsin))

    def tangent(angle):  # Arc::tangent_at_angle
        a = f(angle)
        ex, ey = f(f(-rx) * _sinf(a)), f(ry * _cosf(a))
        return f(f(ex * xcos) - f(ey * xsin)), f(f(ey * xcos) + f(ex * xsin))

    out = []
    t0 = f(0.0)
    for i in range(n):
        a1 = f(start_angle + f(step * f(i)))
        a2 = f(start_angle + f(step * f(i + 1)))
        v1, v2 = ellipse(a1), ellipse(a2)
        p_from = (f(cx + v1[0]), f(cy + v1[1]))
        p_to = (f(cx + v2[0]), f(cy + v2[1]))
        tg = tangent(a1)
        ctrl = (f(p_from[0] + f(tg[0] * alpha)), f(p_from[1] + f(tg[1] * alpha)))
        t1 = f(1.0) if i + 1 == n else f(t0 + dt)
        out.append((ctrl, p_to, t1))
        t0 = t1
    return out


def svg_arc_is_straight_line(frm, to, radii) -> bool:
    """`SvgArc::is_straight_line` (arc.rs:763-767)."""
    eps = np.finfo(np.float32).eps
    return bool(abs(_f32(radii[0])) <= eps or abs(_f32(radii[1])) <= eps or (frm[0] == to[0] and frm[1] == to[1]))


def svg_arc_to_quadratics(frm, to, radii, x_rotation_deg, large_arc: bool, sweep: bool):
    """`SvgArc::for_each_quadratic_bezier_with_t` as the parser calls it (parser.rs:269-306): x_rotation in degrees."""
    xr_rad = _f32(_f32(x_rotation_deg) * _f32(_f32(math.pi) / _f32(180.0)))  # Angle::degrees
    return arc_to_quadratics(*svg_arc_to_arc(frm, to, radii, xr_rad, large_arc, sweep))


def parse_svg_path(src: str, num_attributes: int = 0, stop_at: Optional[str] = None, builder: "PathBuilder" = None) -> Path:
    """Extended SVG path syntax, statement for statement `PathParser::parse` / `parse_path` (crates/extra/src/parser.rs:
    154-360) feeding `Path::builder_with_attributes(n)` as the CLI does (cli/src/main.rs:366-397): implicit commands,
    relative coordinates, smooth curves, arcs (approximated by quadratics with interpolated attributes, parser.rs:269-
    306), `stop_at`, and the reference's errors with its line / column counting.  On an error the events parsed so far
    are still in `builder` when the caller passes one (as with the Rust API)."""
    b = builder if builder is not None else PathBuilder(num_attributes)
    s = _Source(src)
    state = {"cur": (_f32(0), _f32(0)), "need_end": False}
    attrs: List[float] = []

    def parse_number() -> np.float32:
        s.skip_whitespace()
        line, column = s.line, s.col
        buf = []
        if s.current == "-":
            buf.append("-")
            s.advance_one()
        while s.current.isnumeric():
            buf.append(s.current)
            s.advance_one()
        if s.current == ".":
            buf.append(".")
            s.advance_one()
            while s.current.isnumeric():
                buf.append(s.current)
                s.advance_one()
        if s.current in ("e", "E"):
            buf.append(s.current)
            s.advance_one()
            if s.current == "-":
                buf.append("-")
                s.advance_one()
            while s.current.isnumeric():
                buf.append(s.current)
                s.advance_one()
        text = "".join(buf)
        val = _parse_f32(text)
        if val is None:
            raise ParseError("number", line, column, src=text)
        return val

    def parse_flag() -> bool:
        s.skip_whitespace()
        if s.current in "01" and not s.finished:
            v = s.current == "1"
            s.advance_one()
            return v
        raise ParseError("flag", s.line, s.col, src=s.current)

    def parse_attributes():
        attrs.clear()
        for _ in range(num_attributes):
            attrs.append(parse_number())

    def parse_point(rel):
        x, y = parse_number(), parse_number()
        if rel:
            x, y = _f32(x + state["cur"][0]), _f32(y + state["cur"][1])
        return (x, y)

    def parse_endpoint(rel):
        pnt = parse_point(rel)
        state["cur"] = pnt
        parse_attributes()
        return pnt

    def smooth(prev):
        cur = state["cur"]
        if prev is not None:
            return (_f32(cur[0] + _f32(cur[0] - prev[0])), _f32(cur[1] + _f32(cur[1] - prev[1])))
        return cur

    def parse_path():
        state["cur"] = (_f32(0), _f32(0))
        first_position = state["cur"]
        need_start = False
        prev_cubic_ctrl = prev_quad_ctrl = None
        implicit_cmd = "M"
        s.skip_whitespace()
        while not s.finished:
            cmd, cmd_line, cmd_col = s.current, s.line, s.col
            if stop_at is not None and cmd == stop_at:
                break
            if cmd.isascii() and cmd.isalpha():
                s.advance_one()
            else:
                cmd = implicit_cmd
            if need_start and cmd not in "mM":
                raise ParseError("missing_move_to", cmd_line, cmd_col, command=cmd)
            rel = cmd.islower()
            c = cmd.lower()
            if c == "l":
                to = parse_endpoint(rel)
                b.line_to(to, attrs)
            elif c == "h":
                x = parse_number()
                if rel:
                    x = _f32(x + state["cur"][0])
                state["cur"] = (x, state["cur"][1])
                parse_attributes()
                b.line_to(state["cur"], attrs)
            elif c == "v":
                y = parse_number()
                if rel:
                    y = _f32(y + state["cur"][1])
                state["cur"] = (state["cur"][0], y)
                parse_attributes()
                b.line_to(state["cur"], attrs)
            elif c == "q":
                ctrl = parse_point(rel)
                to = parse_endpoint(rel)
                prev_quad_ctrl = ctrl
                b.quadratic_bezier_to(ctrl, to, attrs)
            elif c == "t":
                ctrl = smooth(prev_quad_ctrl)
                to = parse_endpoint(rel)
                prev_quad_ctrl = ctrl
                b.quadratic_bezier_to(ctrl, to, attrs)
            elif c == "c":
                c1 = parse_point(rel)
                c2 = parse_point(rel)
                to = parse_endpoint(rel)
                prev_cubic_ctrl = c2
                b.cubic_bezier_to(c1, c2, to, attrs)
            elif c == "s":
                c1 = smooth(prev_cubic_ctrl)
                c2 = parse_point(rel)
                to = parse_endpoint(rel)
                prev_cubic_ctrl = c2
                b.cubic_bezier_to(c1, c2, to, attrs)
            elif c == "a":
                prev_attributes = list(attrs)
                frm = state["cur"]
                rx, ry = parse_number(), parse_number()
                x_rotation = parse_number()
                large_arc, sweep = parse_flag(), parse_flag()
                to = parse_endpoint(rel)
                if svg_arc_is_straight_line(frm, to, (rx, ry)):
                    b.line_to(to, attrs)
                else:
                    for ctrl, q_to, t_end in svg_arc_to_quadratics(frm, to, (rx, ry), x_rotation, large_arc, sweep):
                        interp = [_f32(_f32(pa * _f32(_f32(1.0) - t_end)) + _f32(_f32(at) * t_end))
                                  for pa, at in zip(prev_attributes + [0.0] * (num_attributes - len(prev_attributes)), attrs)]
                        b.quadratic_bezier_to(ctrl, q_to, interp)
            elif c == "m":
                if state["need_end"]:
                    b.end(False)
                    state["need_end"] = False
                to = parse_endpoint(rel)
                first_position = to
                b.begin(to, attrs)
                state["need_end"] = True
                need_start = False
            elif c == "z":
                b.end(True)
                state["cur"] = first_position
                state["need_end"] = False
                need_start = True
            else:
                raise ParseError("command", cmd_line, cmd_col, command=cmd)
            if c in "cs":
                prev_quad_ctrl = None
            elif c in "qt":
                prev_cubic_ctrl = None
            else:
                prev_cubic_ctrl = None
                prev_quad_ctrl = None
            implicit_cmd = {"m": "l", "M": "L", "z": "m", "Z": "M"}.get(cmd, cmd)
            s.skip_whitespace()

    try:
        parse_path()
    finally:
        if state["need_end"]:  # PathParser::parse, parser.rs:169-171: also after an error
            b.end(False)
    return b.build()


@dataclass
class PathBatch:
    """Many paths in one verb/point buffer + (N+1) ranges: the SoA layout uploaded to the GPU."""

    verbs: np.ndarray  # uint8 [NV]
    points: np.ndarray  # float32 [NP, 2]
    path_verb_begin: np.ndarray  # uint32 [N+1]
    path_point_begin: np.ndarray  # uint32 [N+1]
    num_attributes: int = 0

    @property
    def n_paths(self) -> int:
        return int(self.path_verb_begin.shape[0] - 1)

    @staticmethod
    def from_paths(paths: Iterable[Path]) -> "PathBatch":
        paths = list(paths)
        na = paths[0].num_attributes if paths else 0
        assert all(p.num_attributes == na for p in paths), "all paths of a batch share num_attributes"
        vb = np.zeros(len(paths) + 1, dtype=np.uint64)
        pb = np.zeros(len(paths) + 1, dtype=np.uint64)
        for i, p in enumerate(paths):
            vb[i + 1] = vb[i] + p.verbs.shape[0]
            pb[i + 1] = pb[i] + p.points.shape[0]
        assert vb[-1] < 2**32 and pb[-1] < 2**32
        verbs = np.concatenate([p.verbs for p in paths]) if paths else np.zeros(0, np.uint8)
        points = np.concatenate([p.points.reshape(-1, 2) for p in paths]) if paths else np.zeros((0, 2), np.float32)
        return PathBatch(
            np.ascontiguousarray(verbs, np.uint8),
            np.ascontiguousarray(points, np.float32).reshape(-1, 2),
            vb.astype(np.uint32),
            pb.astype(np.uint32),
            na,
        )

    def path(self, i: int) -> Path:
        v0, v1 = int(self.path_verb_begin[i]), int(self.path_verb_begin[i + 1])
        p0, p1 = int(self.path_point_begin[i]), int(self.path_point_begin[i + 1])
        return Path(self.points[p0:p1].copy(), self.verbs[v0:v1].copy(), self.num_attributes)

    def slice(self, lo: int, hi: int) -> "PathBatch":
        """Paths [lo, hi) as a new batch (ranges rebased to 0)."""
        v0, v1 = int(self.path_verb_begin[lo]), int(self.path_verb_begin[hi])
        p0, p1 = int(self.path_point_begin[lo]), int(self.path_point_begin[hi])
        return PathBatch(
            self.verbs[v0:v1],
            self.points[p0:p1],
            (self.path_verb_begin[lo : hi + 1] - np.uint32(v0)).astype(np.uint32),
            (self.path_point_begin[lo : hi + 1] - np.uint32(p0)).astype(np.uint32),
            self.num_attributes,
        )
