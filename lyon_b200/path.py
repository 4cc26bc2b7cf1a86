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

    def build(self) -> Path:
        self._end_if_needed()
        return self.builder.build()


class ParseError(ValueError):
    pass


def parse_svg_path(src: str, num_attributes: int = 0) -> Path:
    """Extended SVG path syntax parser, mirroring `PathParser::parse`
    (crates/extra/src/parser.rs:154-360, 362-467) feeding `Path::builder_with_attributes(n)` as the CLI
    does (cli/src/main.rs:366-397).  Arcs ('a'/'A') are not supported (they are outside the hot path)."""
    b = PathBuilder(num_attributes)
    n = len(src)
    pos = 0
    cur = (_f32(0), _f32(0))
    first_position = cur
    need_start = False
    need_end = False
    prev_cubic_ctrl: Optional[tuple] = None
    prev_quad_ctrl: Optional[tuple] = None
    implicit_cmd = "M"
    attrs: List[float] = []

    def skip_ws():
        nonlocal pos
        while pos < n and (src[pos].isspace() or src[pos] == ","):
            pos += 1

    def number() -> np.float32:
        nonlocal pos
        skip_ws()
        s = pos
        if pos < n and src[pos] == "-":
            pos += 1
        while pos < n and src[pos].isdigit():
            pos += 1
        if pos < n and src[pos] == ".":
            pos += 1
            while pos < n and src[pos].isdigit():
                pos += 1
        if pos < n and src[pos] in "eE":
            pos += 1
            if pos < n and src[pos] == "-":
                pos += 1
            while pos < n and src[pos].isdigit():
                pos += 1
        try:
            # short decimals: f64 parse then f32 rounding equals Rust's correctly rounded str::parse::<f32>
            return _f32(float(src[s:pos]))
        except ValueError:
            raise ParseError(f"bad number {src[s:pos]!r} at {s}")

    def attributes():
        attrs.clear()
        for _ in range(num_attributes):
            attrs.append(number())

    def point(rel):
        x, y = number(), number()
        if rel:
            x, y = x + cur[0], y + cur[1]
        return (x, y)

    def endpoint(rel):
        nonlocal cur
        p = point(rel)
        cur = p
        attributes()
        return p

    def smooth(prev):
        if prev is not None:
            return (cur[0] + (cur[0] - prev[0]), cur[1] + (cur[1] - prev[1]))
        return cur

    skip_ws()
    while pos < n:
        cmd = src[pos]
        if cmd.isascii() and cmd.isalpha():
            pos += 1
        else:
            cmd = implicit_cmd
        if need_start and cmd not in "mM":
            raise ParseError(f"missing move-to before {cmd!r}")
        rel = cmd.islower()
        c = cmd.lower()
        if c == "l":
            to = endpoint(rel)
            b.line_to(to, attrs)
        elif c == "h":
            x = number()
            if rel:
                x = x + cur[0]
            cur = (x, cur[1])
            attributes()
            b.line_to(cur, attrs)
        elif c == "v":
            y = number()
            if rel:
                y = y + cur[1]
            cur = (cur[0], y)
            attributes()
            b.line_to(cur, attrs)
        elif c == "q":
            ctrl = point(rel)
            to = endpoint(rel)
            prev_quad_ctrl = ctrl
            b.quadratic_bezier_to(ctrl, to, attrs)
        elif c == "t":
            ctrl = smooth(prev_quad_ctrl)
            to = endpoint(rel)
            prev_quad_ctrl = ctrl
            b.quadratic_bezier_to(ctrl, to, attrs)
        elif c == "c":
            c1 = point(rel)
            c2 = point(rel)
            to = endpoint(rel)
            prev_cubic_ctrl = c2
            b.cubic_bezier_to(c1, c2, to, attrs)
        elif c == "s":
            c1 = smooth(prev_cubic_ctrl)
            c2 = point(rel)
            to = endpoint(rel)
            prev_cubic_ctrl = c2
            b.cubic_bezier_to(c1, c2, to, attrs)
        elif c == "m":
            if need_end:
                b.end(False)
                need_end = False
            to = endpoint(rel)
            first_position = to
            b.begin(to, attrs)
            need_end = True
            need_start = False
        elif c == "z":
            b.end(True)
            cur = first_position
            need_end = False
            need_start = True
        else:
            raise ParseError(f"unsupported command {cmd!r} at {pos}")
        if c in "cs":
            prev_quad_ctrl = None
        elif c in "qt":
            prev_cubic_ctrl = None
        else:
            prev_cubic_ctrl = None
            prev_quad_ctrl = None
        implicit_cmd = {"m": "l", "M": "L", "z": "m", "Z": "M"}.get(cmd, cmd)
        skip_ws()
    if need_end:
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
