"""Synthetic workloads of BASELINE.json / SURVEY 8(d), generated as SoA batches with numpy (seeded)."""
from __future__ import annotations

import numpy as np

from lyon_b200.path import BEGIN, CLOSE, CUBIC_TO, END, LINE_TO, QUADRATIC_TO, Path, PathBatch, parse_svg_path


def c2_cubics(n: int, seed: int = 1234) -> np.ndarray:
    """C2: n random cubics, from ~ U[0,1000)^2, other control points = from + U[-4,4)^2.  (n, 8) f32."""
    rng = np.random.default_rng(seed)
    frm = rng.uniform(0, 1000, size=(n, 1, 2)).astype(np.float32)
    off = rng.uniform(-4, 4, size=(n, 3, 2)).astype(np.float32)
    return np.concatenate([frm, frm + off], axis=1).reshape(n, 8).astype(np.float32)


def c3_mixed_paths(n_paths: int, n_segments: int = 64, seed: int = 2024, closed_fraction: float = 0.0) -> PathBatch:
    """C3: n_paths open paths x n_segments mixed line/quad/cubic segments."""
    rng = np.random.default_rng(seed)
    kinds = rng.integers(0, 3, size=(n_paths, n_segments)).astype(np.uint8)  # 0 line 1 quad 2 cubic == verb codes
    begin = rng.uniform(0, 4096, size=(n_paths, 1, 2)).astype(np.float32)
    steps = rng.uniform(-16, 16, size=(n_paths, n_segments, 2)).astype(np.float32)
    ends = (begin + np.cumsum(steps, axis=1, dtype=np.float32)).astype(np.float32)  # endpoints
    starts = np.concatenate([begin, ends[:, :-1]], axis=1)
    jit = rng.uniform(-8, 8, size=(n_paths, n_segments, 2, 2)).astype(np.float32)
    l1 = (starts + (ends - starts) * np.float32(1 / 3)).astype(np.float32)
    l2 = (starts + (ends - starts) * np.float32(2 / 3)).astype(np.float32)
    mid = (starts + (ends - starts) * np.float32(0.5)).astype(np.float32)
    c1 = np.where((kinds == 1)[..., None], mid, l1) + jit[:, :, 0]
    c2 = l2 + jit[:, :, 1]
    closed = rng.uniform(size=n_paths) < closed_fraction
    npts_seg = kinds.astype(np.int64) + 1  # stored points per segment
    pts_per_path = 1 + npts_seg.sum(axis=1) + closed.astype(np.int64)
    verbs_per_path = n_segments + 2
    verbs = np.empty((n_paths, verbs_per_path), np.uint8)
    verbs[:, 0] = BEGIN
    verbs[:, 1:-1] = kinds
    verbs[:, -1] = np.where(closed, CLOSE, END)
    path_point_begin = np.zeros(n_paths + 1, np.int64)
    np.cumsum(pts_per_path, out=path_point_begin[1:])
    points = np.zeros((int(path_point_begin[-1]), 2), np.float32)
    seg_off = np.cumsum(npts_seg, axis=1) - npts_seg + 1 + path_point_begin[:-1, None]  # first point of each segment
    points[path_point_begin[:-1]] = begin[:, 0]
    is_c, is_q, is_l = kinds == 2, kinds == 1, kinds == 0
    points[seg_off[is_l]] = ends[is_l]
    points[seg_off[is_q]] = c1[is_q]
    points[seg_off[is_q] + 1] = ends[is_q]
    points[seg_off[is_c]] = c1[is_c]
    points[seg_off[is_c] + 1] = c2[is_c]
    points[seg_off[is_c] + 2] = ends[is_c]
    if closed.any():
        points[path_point_begin[1:][closed] - 1] = begin[closed, 0]
    return PathBatch(verbs.reshape(-1), points, (np.arange(n_paths + 1) * verbs_per_path).astype(np.uint32),
                     path_point_begin.astype(np.uint32), 0)


def c4_polylines(n_paths: int, n_segments: int = 1000, seed: int = 4242, with_attribute: bool = True) -> PathBatch:
    """C4: polylines, random walk U[-4,4)^2 steps with 10% sharp reversals, 1 custom attribute (width factor)."""
    rng = np.random.default_rng(seed)
    begin = rng.uniform(0, 4096, size=(n_paths, 1, 2)).astype(np.float32)
    steps = rng.uniform(-4, 4, size=(n_paths, n_segments, 2)).astype(np.float32)
    rev = rng.uniform(size=(n_paths, n_segments)) < 0.1
    prev = np.concatenate([np.zeros((n_paths, 1, 2), np.float32), steps[:, :-1]], axis=1)
    steps = np.where(rev[..., None], (-prev * np.float32(0.9) + steps * np.float32(0.05)).astype(np.float32), steps)
    pts = np.concatenate([begin, begin + np.cumsum(steps, axis=1, dtype=np.float32)], axis=1).astype(np.float32)
    n_pt = n_segments + 1
    if with_attribute:
        width = rng.uniform(0.5, 2.0, size=(n_paths, n_pt)).astype(np.float32)
        points = np.zeros((n_paths, n_pt, 2, 2), np.float32)
        points[:, :, 0, :] = pts
        points[:, :, 1, 0] = width
        points = points.reshape(-1, 2)
        stride = 2
    else:
        points = pts.reshape(-1, 2)
        stride = 1
    verbs = np.empty((n_paths, n_segments + 2), np.uint8)
    verbs[:, 0] = BEGIN
    verbs[:, 1:-1] = LINE_TO
    verbs[:, -1] = END
    return PathBatch(verbs.reshape(-1), np.ascontiguousarray(points),
                     (np.arange(n_paths + 1) * (n_segments + 2)).astype(np.uint32),
                     (np.arange(n_paths + 1) * n_pt * stride).astype(np.uint32), 1 if with_attribute else 0)


def c5_logo_copies(paths, n_copies: int, seed: int = 55) -> PathBatch:
    """C5: the fixture paths replicated with per-copy affine jitter (translate U[0,1e4)^2, scale U[0.5,2))."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_copies):
        p = paths[i % len(paths)]
        tr = rng.uniform(0, 1e4, size=2).astype(np.float32)
        sc = np.float32(rng.uniform(0.5, 2.0))
        out.append(Path((p.points * sc + tr).astype(np.float32), p.verbs, 0))
    return PathBatch.from_paths(out)


def random_small_paths(n_paths: int, seed: int, max_segments: int = 12, closed_fraction=0.4, degenerate=0.15) -> PathBatch:
    """Adversarial little paths for parity: mixed verbs, closed/open sub-paths, several sub-paths per path,
    duplicate points, tiny edges around the merge threshold, sharp reversals, empty sub-paths."""
    rng = np.random.default_rng(seed)
    paths = []
    for _ in range(n_paths):
        b = Path.builder()
        for _sp in range(int(rng.integers(1, 4))):
            cur = rng.uniform(-50, 50, size=2).astype(np.float32)
            b.begin(tuple(cur))
            nseg = int(rng.integers(0, max_segments + 1))
            last_step = np.array([1.0, 0.0], np.float32)
            for _s in range(nseg):
                r = rng.uniform()
                if r < degenerate * 0.4:
                    step = np.zeros(2, np.float32)  # exact duplicate
                elif r < degenerate * 0.8:
                    step = rng.uniform(-0.08, 0.08, size=2).astype(np.float32)  # around the merge distance
                elif r < degenerate:
                    step = (-last_step * np.float32(rng.uniform(0.5, 1.5))).astype(np.float32)  # reversal
                else:
                    step = rng.uniform(-20, 20, size=2).astype(np.float32)
                to = (cur + step).astype(np.float32)
                k = int(rng.integers(0, 3))
                if k == 0:
                    b.line_to(tuple(to))
                elif k == 1:
                    far = rng.uniform() < 0.2
                    ctrl = cur + rng.uniform(-30, 30, size=2).astype(np.float32) * np.float32(8.0 if far else 1.0)
                    b.quadratic_bezier_to(tuple(ctrl.astype(np.float32)), tuple(to))
                else:
                    c1 = cur + rng.uniform(-30, 30, size=2).astype(np.float32)
                    c2 = to + rng.uniform(-30, 30, size=2).astype(np.float32)
                    b.cubic_bezier_to(tuple(c1.astype(np.float32)), tuple(c2.astype(np.float32)), tuple(to))
                if np.any(step != 0):
                    last_step = step
                cur = to
            b.end(bool(rng.uniform() < closed_fraction))
        paths.append(b.build())
    return PathBatch.from_paths(paths)


def random_width_paths(n_paths: int, seed: int, num_attributes: int = 1, max_segments: int = 10, closed_fraction=0.4,
                       degenerate=0.15) -> PathBatch:
    """Adversarial little paths with custom attributes (attribute 0.. are width factors in [0.2, 3)) for the variable
    line width path: mixed verbs, closed/open sub-paths, duplicates, reversals, far control points."""
    rng = np.random.default_rng(seed)
    paths = []

    def attrs():
        return [float(np.float32(rng.uniform(0.2, 3.0))) for _ in range(num_attributes)]

    for _ in range(n_paths):
        b = Path.builder_with_attributes(num_attributes)
        for _sp in range(int(rng.integers(1, 3))):
            cur = rng.uniform(-50, 50, size=2).astype(np.float32)
            b.begin(tuple(cur), attrs())
            last_step = np.array([1.0, 0.0], np.float32)
            for _s in range(int(rng.integers(0, max_segments + 1))):
                r = rng.uniform()
                if r < degenerate * 0.4:
                    step = np.zeros(2, np.float32)
                elif r < degenerate * 0.8:
                    step = rng.uniform(-0.08, 0.08, size=2).astype(np.float32)
                elif r < degenerate:
                    step = (-last_step * np.float32(rng.uniform(0.5, 1.5))).astype(np.float32)
                else:
                    step = rng.uniform(-20, 20, size=2).astype(np.float32)
                to = (cur + step).astype(np.float32)
                k = int(rng.integers(0, 3))
                if k == 0:
                    b.line_to(tuple(to), attrs())
                elif k == 1:
                    far = rng.uniform() < 0.2
                    ctrl = cur + rng.uniform(-30, 30, size=2).astype(np.float32) * np.float32(8.0 if far else 1.0)
                    b.quadratic_bezier_to(tuple(ctrl.astype(np.float32)), tuple(to), attrs())
                else:
                    c1 = cur + rng.uniform(-30, 30, size=2).astype(np.float32)
                    c2 = to + rng.uniform(-30, 30, size=2).astype(np.float32)
                    b.cubic_bezier_to(tuple(c1.astype(np.float32)), tuple(c2.astype(np.float32)), tuple(to), attrs())
                if np.any(step != 0):
                    last_step = step
                cur = to
            b.end(bool(rng.uniform() < closed_fraction))
        paths.append(b.build())
    return PathBatch.from_paths(paths)


def fixture_paths(golden_dir):
    return [parse_svg_path((golden_dir / "logo.path").read_text()), parse_svg_path((golden_dir / "rust_logo.path").read_text())]


def concat_batches(batches) -> PathBatch:
    batches = list(batches)
    if len(batches) == 1:
        return batches[0]
    vb = [np.zeros(1, np.uint64)]
    pb = [np.zeros(1, np.uint64)]
    v0 = p0 = 0
    for b in batches:
        vb.append(b.path_verb_begin[1:].astype(np.uint64) + np.uint64(v0))
        pb.append(b.path_point_begin[1:].astype(np.uint64) + np.uint64(p0))
        v0 += int(b.path_verb_begin[-1])
        p0 += int(b.path_point_begin[-1])
    assert v0 < 2**32 and p0 < 2**32
    return PathBatch(np.concatenate([b.verbs for b in batches]), np.concatenate([b.points for b in batches]),
                     np.concatenate(vb).astype(np.uint32), np.concatenate(pb).astype(np.uint32), batches[0].num_attributes)


C3_CHUNK = 25000


def c3_batch_range(lo: int, hi: int, n_segments: int = 64, seed: int = 2024) -> PathBatch:
    """Paths [lo, hi) of the C3 workload.  The workload is defined chunk-wise (C3_CHUNK paths per chunk, chunk c
    seeded with seed + c) so that any rank can generate exactly its own slice of the same 1M-path batch."""
    assert lo % C3_CHUNK == 0 and (hi % C3_CHUNK == 0 or True)
    out = []
    c = lo // C3_CHUNK
    pos = lo
    while pos < hi:
        n = min(C3_CHUNK, hi - pos)
        b = c3_mixed_paths(C3_CHUNK, n_segments, seed=seed + c)
        out.append(b if n == C3_CHUNK else b.slice(0, n))
        pos += n
        c += 1
    return concat_batches(out)


from lyon_b200.sharding import partition_by_cost  # noqa: E402,F401  (the product owns the partitioner)
