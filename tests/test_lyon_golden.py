"""Golden dumps of the REAL lyon StrokeTessellator (tools/dump_lyon_golden, needs cargo: not available in this image).

When `tests/golden/*__<join>_<cap>_<tol>.bin` files are present, the oracle must reproduce them: vertex and index counts
and the index array exactly, positions / normals / advancement within 1e-5 x the path's extent (the tolerance the north
star states; the dumps come from platform libm, the oracle rounds double libm once -- tests/test_libm_pinning.py).  With
no dump present the test reports itself skipped: the oracle <-> lyon gap stays "partial" (DESIGN.md section 5)."""
import re
import struct
from pathlib import Path as FsPath

import numpy as np
import pytest

import oracle_ffi as O
import workloads as W
from lyon_b200 import LineCap, LineJoin, PathBatch, StrokeOptions, parse_svg_path
from lyon_b200.geometry import STROKE_VERTEX_DTYPE
from parity import bbox_extent

GOLDEN = FsPath(__file__).resolve().parent / "golden"
JOINS = {"miter": LineJoin.Miter, "miterclip": LineJoin.MiterClip, "round": LineJoin.Round, "bevel": LineJoin.Bevel}
CAPS = {"butt": LineCap.Butt, "square": LineCap.Square, "round": LineCap.Round}


def read_dump(path: FsPath):
    raw = path.read_bytes()
    magic, nv, ni = struct.unpack_from("<III", raw, 0)
    assert magic == 0x4C594F4E, "not a dump_lyon_golden file"
    v = np.frombuffer(raw, dtype=STROKE_VERTEX_DTYPE, count=nv, offset=12)
    i = np.frombuffer(raw, dtype=np.uint32, count=ni, offset=12 + 32 * nv)
    return v, i


def test_dump_format_round_trips(tmp_path):
    """The reader above against a file written the way tools/dump_lyon_golden/src/main.rs writes it."""
    batch = PathBatch.from_paths([W.fixture_paths(GOLDEN)[0]])
    m = O.stroke_batch(batch, StrokeOptions(tolerance=0.05))
    f = tmp_path / "logo__miter_butt_0.05.bin"
    f.write_bytes(struct.pack("<III", 0x4C594F4E, len(m.vertices), len(m.indices)) + m.vertices.tobytes() + m.indices.tobytes())
    v, i = read_dump(f)
    assert v.tobytes() == m.vertices.tobytes() and np.array_equal(i, m.indices)


def test_oracle_reproduces_lyon_dumps():
    dumps = sorted(GOLDEN.glob("*__*_*_*.bin"))
    if not dumps:
        pytest.skip("no lyon dumps in tests/golden (run tools/dump_lyon_golden with cargo to create them)")
    for f in dumps:
        m = re.match(r"(.+)__([a-z]+)_([a-z]+)_([0-9.]+)\\.bin$", f.name)
        assert m, f.name
        stem, join, cap, tol = m.group(1), JOINS[m.group(2)], CAPS[m.group(3)], float(m.group(4))
        path = parse_svg_path((GOLDEN / f"{stem}.path").read_text())
        batch = PathBatch.from_paths([path])
        o = StrokeOptions(tolerance=tol, line_width=1.0, line_join=join).with_line_cap(cap)
        want_v, want_i = read_dump(f)
        got = O.stroke_batch(batch, o)
        assert len(got.vertices) == len(want_v) and len(got.indices) == len(want_i), f.name
        assert np.array_equal(got.indices, want_i), f.name
        tol_abs = 1e-5 * bbox_extent(batch)
        for field in ("position_on_path", "normal", "advancement"):
            assert np.allclose(got.vertices[field], want_v[field], atol=tol_abs, rtol=0, equal_nan=True), (f.name, field)
        assert np.array_equal(got.vertices["src"], want_v["src"]) and np.array_equal(got.vertices["t"], want_v["t"]), f.name
