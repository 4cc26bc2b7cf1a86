"""Shape helpers of lyon's builders and of `StrokeTessellator`, through the batch kernels.

* path generators with the reference's f32 arithmetic: `add_polygon` / `add_rectangle` / `add_circle` / `add_ellipse` /
  `add_rounded_rectangle` / `add_line_segment` (crates/path/src/builder.rs:555-700, 1506-1660);
* `StrokeTessellator.tessellate_polygon / tessellate_rectangle / tessellate_circle / tessellate_ellipse`
  (crates/tessellation/src/stroke.rs:255-307) for ONE shape or a whole batch of shapes, incl. the thin-rectangle
  fall-back `approximate_thin_rectangle` (stroke.rs:523-537, 2336-2369);
* `StrokeBuilder`: the streaming `StrokeTessellator::builder()` (stroke.rs:215-252, 312-398) -- commands are recorded
  and tessellated in batches on `build()`; options may change BETWEEN sub-paths (`set_line_join` and friends,
  stroke.rs:380-398), a change inside a sub-path raises (the kernels take one StrokeOptions per batch).

The streaming builder hands out sequential EndpointIds (`attrib_store.add`, stroke.rs:412-470) exactly like
`StrokeTessellator::tessellate` does for a plain event stream (stroke.rs:882-937), so stroking the generated Path gives
the vertices (incl. their `src` ids) the reference's builder produces."""
from __future__ import annotations

import math
from dataclasses import replace
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .geometry import VertexBuffers
from .options import LineCap, LineJoin, StrokeOptions
from .path import Path, PathBatch, PathBuilder, arc_sample, arc_to_quadratics

f32 = np.float32
POSITIVE, NEGATIVE = 1, -1
_CIRCLE_K = f32(0.55191505)  # https://spencermortensen.com/articles/bezier-circle/ (builder.rs:1521)


def _p(x, y):
    return (f32(x), f32(y))


def add_polygon(b: PathBuilder, points: Sequence, closed: bool, attributes=None):
    """builder.rs:560-573."""
    if len(points) == 0:
        return
    b.begin(points[0], attributes)
    for p in points[1:]:
        b.line_to(p, attributes)
    b.end(bool(closed))


def add_line_segment(b: PathBuilder, frm, to, attributes=None):
    """builder.rs:590-600."""
    b.begin(frm, attributes)
    b.line_to(to, attributes)
    b.end(False)


def add_rectangle(b: PathBuilder, rect_min, rect_max, winding: int = POSITIVE, attributes=None):
    """builder.rs:649-678."""
    (x0, y0), (x1, y1) = _p(*rect_min), _p(*rect_max)
    pts = [(x0, y0), (x1, y0), (x1, y1), (x0, y1)] if winding == POSITIVE else [(x0, y0), (x0, y1), (x1, y1), (x1, y0)]
    add_polygon(b, pts, True, attributes)


def add_circle(b: PathBuilder, center, radius, winding: int = POSITIVE, attributes=None):
    """builder.rs:1508-1546: four cubics."""
    cx, cy = _p(*center)
    r = f32(abs(f32(radius)))
    dirn = f32(1.0 if winding == POSITIVE else -1.0)
    d = f32(r * _CIRCLE_K)

    def at(dx, dy):
        return (f32(cx + f32(dx)), f32(cy + f32(dy)))

    b.begin(at(-r, f32(0.0)), attributes)
    b.cubic_bezier_to(at(-r, f32(f32(-d) * dirn)), at(-d, f32(f32(-r) * dirn)), at(f32(0.0), f32(f32(-r) * dirn)), attributes)
    b.cubic_bezier_to(at(d, f32(f32(-r) * dirn)), at(r, f32(f32(-d) * dirn)), at(r, f32(0.0)), attributes)
    b.cubic_bezier_to(at(r, f32(d * dirn)), at(d, f32(r * dirn)), at(f32(0.0), f32(r * dirn)), attributes)
    b.cubic_bezier_to(at(-d, f32(r * dirn)), at(-r, f32(d * dirn)), at(-r, f32(0.0)), attributes)
    b.end(True)


def add_ellipse(b: PathBuilder, center, radii, x_rotation: float, winding: int = POSITIVE, attributes=None):
    """builder.rs:606-636: the arc 0 .. +-2 pi as quadratics (crates/geom/src/arc.rs:668-703)."""
    dirn = f32(1.0 if winding == POSITIVE else -1.0)
    sweep = f32(f32(f32(2.0) * f32(math.pi)) * dirn)
    start = f32(0.0)
    b.begin(arc_sample(center, radii, f32(x_rotation), f32(start + f32(sweep * f32(0.0)))), attributes)
    for ctrl, to, _t in arc_to_quadratics(center, radii, start, sweep, f32(x_rotation)):
        b.quadratic_bezier_to(ctrl, to, attributes)
    b.end(True)


def add_rounded_rectangle(b: PathBuilder, rect_min, rect_max, radii: Sequence[float], winding: int = POSITIVE, attributes=None):
    """builder.rs:1551-1660; radii = (top_left, top_right, bottom_left, bottom_right)."""
    (x_min, y_min), (x_max, y_max) = _p(*rect_min), _p(*rect_max)
    w, h = f32(x_max - x_min), f32(y_max - y_min)
    min_wh = min(w, h)
    tl, tr, bl, br = (min(f32(abs(f32(r))), min_wh) for r in radii)
    if f32(tl + tr) > w:
        x = f32(f32(f32(tl + tr) - w) * f32(0.5))
        tl, tr = f32(tl - x), f32(tr - x)
    if f32(bl + br) > w:
        x = f32(f32(f32(bl + br) - w) * f32(0.5))
        bl, br = f32(bl - x), f32(br - x)
    if f32(tr + br) > h:
        x = f32(f32(f32(tr + br) - h) * f32(0.5))
        tr, br = f32(tr - x), f32(br - x)
    if f32(tl + bl) > h:
        x = f32(f32(f32(tl + bl) - h) * f32(0.5))
        tl, bl = f32(tl - x), f32(bl - x)
    tl_d, tr_d, br_d, bl_d = (f32(r * _CIRCLE_K) for r in (tl, tr, br, bl))
    z = f32(0.0)

    def add(c, dx, dy):
        return (f32(c[0] + f32(dx)), f32(c[1] + f32(dy)))

    tlc, trc, brc, blc = (x_min, y_min), (x_max, y_min), (x_max, y_max), (x_min, y_max)
    pts = [
        (x_min, f32(y_min + tl)), add(tlc, z, f32(tl - tl_d)), add(tlc, f32(tl - tl_d), z), add(tlc, tl, z),
        (f32(x_max - tr), y_min), add(trc, f32(f32(-tr) + tr_d), z), add(trc, z, f32(tr - tr_d)), add(trc, z, tr),
        (x_max, f32(y_max - br)), add(brc, z, f32(f32(-br) + br_d)), add(brc, f32(f32(-br) + br_d), z), add(brc, f32(-br), z),
        (f32(x_min + bl), y_max), add(blc, f32(bl - bl_d), z), add(blc, z, f32(f32(-bl) + bl_d)), add(blc, z, f32(-bl)),
    ]
    if winding == POSITIVE:
        b.begin(pts[0], attributes)
        if tl > 0:
            b.cubic_bezier_to(pts[1], pts[2], pts[3], attributes)
        b.line_to(pts[4], attributes)
        if tr > 0:
            b.cubic_bezier_to(pts[5], pts[6], pts[7], attributes)
        b.line_to(pts[8], attributes)
        if br > 0:
            b.cubic_bezier_to(pts[9], pts[10], pts[11], attributes)
        b.line_to(pts[12], attributes)
        if bl > 0:
            b.cubic_bezier_to(pts[13], pts[14], pts[15], attributes)
    else:
        b.begin(pts[15], attributes)
        if bl > 0:
            b.cubic_bezier_to(pts[14], pts[13], pts[12], attributes)
        b.line_to(pts[11], attributes)
        if br > 0:
            b.cubic_bezier_to(pts[10], pts[9], pts[8], attributes)
        b.line_to(pts[7], attributes)
        if tr > 0:
            b.cubic_bezier_to(pts[6], pts[5], pts[4], attributes)
        b.line_to(pts[3], attributes)
        if tl > 0:
            b.cubic_bezier_to(pts[2], pts[1], pts[0], attributes)
    b.end(True)


def thin_rectangle(rect_min, rect_max, options: StrokeOptions):
    """`StrokeBuilder::add_rectangle`'s choice (stroke.rs:523-537) and `approximate_thin_rectangle` (stroke.rs:2336-2369):
    None when the rectangle is stroked as a polygon, else (from, to, options') -- the line segment and the options
    (wider line, square or round caps) that stand in for it."""
    (x0, y0), (x1, y1) = _p(*rect_min), _p(*rect_max)
    lw = f32(options.line_width)
    threshold = f32(f32(1.0 if options.line_join == LineJoin.Miter else 0.05) * lw)
    w, h = f32(x1 - x0), f32(y1 - y0)
    if options.variable_line_width is not None or not (abs(w) < threshold or abs(h) < threshold):
        return None
    if w > h:
        d = f32(h * f32(0.5))
        y = f32(f32(y0 + y1) * f32(0.5))
        frm, to = (f32(x0 + d), y), (f32(x1 - d), y)
    else:
        d = f32(w * f32(0.5))
        x = f32(f32(x0 + x1) * f32(0.5))
        frm, to = (x, f32(y0 + d)), (x, f32(y1 - d))
    cap = LineCap.Round if options.line_join == LineJoin.Round else LineCap.Square
    return frm, to, replace(options, line_width=float(f32(lw + d)), start_cap=cap, end_cap=cap)


class StrokeTessellator:
    """Host mirror of `lyon_tessellation::StrokeTessellator` (stroke.rs:108-307) over a `DeviceTessellator`: every call
    appends to `output` (a `VertexBuffers`, indices offset by the vertices already there like `BuffersBuilder` does,
    geometry_builder.rs:325-349) or returns a fresh one."""

    def __init__(self, tessellator=None):
        from .tessellator import _default

        self.tz = tessellator or _default()

    # -- the general entry points -----------------------------------------------------------------------------------
    def tessellate_path(self, path: Path, options: StrokeOptions, output: Optional[VertexBuffers] = None) -> VertexBuffers:
        return self.tessellate_paths([path], options, output)

    tessellate = tessellate_path

    def tessellate_paths(self, paths: Sequence[Path], options: StrokeOptions, output: Optional[VertexBuffers] = None) -> VertexBuffers:
        from .tessellator import tessellate_paths

        vb = tessellate_paths(PathBatch.from_paths(list(paths)), options, self.tz)
        if output is None:
            return vb
        base = len(output.vertices)
        if base + len(vb.vertices) > 0xFFFFFFFF:
            from .geometry import TooManyVertices
            raise TooManyVertices(-2, "TooManyVertices")
        output.vertices = np.concatenate([output.vertices, vb.vertices])
        output.indices = np.concatenate([output.indices, (vb.indices + np.uint32(base)).astype(np.uint32)])
        return output

    # -- shapes (stroke.rs:255-307) ---------------------------------------------------------------------------------
    def tessellate_polygon(self, points, closed: bool, options, output=None):
        b = Path.builder()
        add_polygon(b, [_p(*q) for q in points], closed)
        return self.tessellate_path(b.build(), options, output)

    def tessellate_circle(self, center, radius, options, output=None):
        b = Path.builder()
        add_circle(b, center, radius, POSITIVE)
        return self.tessellate_path(b.build(), options, output)

    def tessellate_ellipse(self, center, radii, x_rotation, winding, options, output=None):
        b = Path.builder()
        add_ellipse(b, center, radii, x_rotation, winding)
        return self.tessellate_path(b.build(), options, output)

    def tessellate_rectangle(self, rect_min, rect_max, options, output=None):
        assert options.variable_line_width is None  # stroke.rs:272
        return self.tessellate_rectangles([(rect_min, rect_max)], options, output)

    def tessellate_rectangles(self, rects: Sequence[Tuple], options, output=None):
        """A batch of axis-aligned rectangles in order.  Runs of rectangles that share their effective options (all the
        polygon ones; thin ones of equal thickness) go to the device together."""
        out = output
        run: List[Path] = []
        run_opts = None

        def flush():
            nonlocal out, run, run_opts
            if run:
                out = self.tessellate_paths(run, run_opts, out)
                run, run_opts = [], None

        for rmin, rmax in rects:
            b = Path.builder()
            thin = thin_rectangle(rmin, rmax, options)
            if thin is None:
                add_rectangle(b, rmin, rmax, POSITIVE)
                o = options
            else:
                add_line_segment(b, thin[0], thin[1])
                o = thin[2]
            if run_opts is not None and o != run_opts:
                flush()
            run.append(b.build())
            run_opts = o
        flush()
        return out if out is not None else VertexBuffers.empty()

    def builder(self, options: StrokeOptions, output: Optional[VertexBuffers] = None) -> "StrokeBuilder":
        return StrokeBuilder(self, options, output)


class StrokeBuilder:
    """`StrokeTessellator::builder()` (stroke.rs:215-252): PathBuilder commands tessellated without an intermediate path in
    the reference; here the sub-paths are recorded and tessellated by `build()`, all the sub-paths entered under the same
    options as ONE path (EndpointIds and the advancement of open sub-paths run on from one sub-path to the next, stroke.rs:
    412-470, 1359).  A change of options -- or a thin rectangle, which the reference strokes with temporary options --
    starts a new path: ids and advancement restart there (the one difference from the streaming reference)."""

    def __init__(self, owner: StrokeTessellator, options: StrokeOptions, output: Optional[VertexBuffers]):
        self.owner, self.options, self.output = owner, options, output
        self._runs: List[Tuple[StrokeOptions, List[Path]]] = []
        self._b: Optional[PathBuilder] = None
        self._in_subpath = False

    # option changes (stroke.rs:380-398) -- between sub-paths only
    def _set(self, **kw):
        if self._in_subpath:
            raise NotImplementedError("stroke options cannot change inside a sub-path on the batch kernels")
        self._flush_builder()
        self.options = replace(self.options, **kw)

    def set_line_join(self, join: LineJoin):
        self._set(line_join=join)

    def set_start_cap(self, cap: LineCap):
        self._set(start_cap=cap)

    def set_end_cap(self, cap: LineCap):
        self._set(end_cap=cap)

    def set_miter_limit(self, limit: float):
        self._set(miter_limit=float(limit))

    def set_line_width(self, width: float):
        self._set(line_width=float(width))

    def _builder(self) -> PathBuilder:
        if self._b is None:
            self._b = Path.builder()
        return self._b

    def _flush_builder(self):
        if self._b is not None:
            path = self._b.build()
            self._b = None
            if path.verbs.shape[0]:
                if self._runs and self._runs[-1][0] == self.options:
                    self._runs[-1][1].append(path)
                else:
                    self._runs.append((self.options, [path]))

    # PathBuilder (stroke.rs:409-520)
    def begin(self, at):
        self._in_subpath = True
        return self._builder().begin(at)

    def line_to(self, to):
        return self._builder().line_to(to)

    def quadratic_bezier_to(self, ctrl, to):
        return self._builder().quadratic_bezier_to(ctrl, to)

    def cubic_bezier_to(self, c1, c2, to):
        return self._builder().cubic_bezier_to(c1, c2, to)

    def end(self, close: bool):
        self._builder().end(close)
        self._in_subpath = False

    def close(self):
        self.end(True)

    def add_polygon(self, points, closed: bool):
        add_polygon(self._builder(), [_p(*q) for q in points], closed)

    def add_circle(self, center, radius, winding: int = POSITIVE):
        add_circle(self._builder(), center, radius, winding)

    def add_ellipse(self, center, radii, x_rotation, winding: int = POSITIVE):
        add_ellipse(self._builder(), center, radii, x_rotation, winding)

    def add_rounded_rectangle(self, rect_min, rect_max, radii, winding: int = POSITIVE):
        add_rounded_rectangle(self._builder(), rect_min, rect_max, radii, winding)

    def add_rectangle(self, rect_min, rect_max, winding: int = POSITIVE):
        thin = thin_rectangle(rect_min, rect_max, self.options)
        if thin is None:
            add_rectangle(self._builder(), rect_min, rect_max, winding)
            return
        self._flush_builder()
        b = Path.builder()
        add_line_segment(b, thin[0], thin[1])
        self._runs.append((thin[2], [b.build()]))

    def build(self) -> VertexBuffers:
        """Tessellate everything recorded, in order."""
        self._flush_builder()
        out = self.output
        for opts, paths in self._runs:
            out = self.owner.tessellate_paths(paths, opts, out)
        self._runs = []
        return out if out is not None else VertexBuffers.empty()
