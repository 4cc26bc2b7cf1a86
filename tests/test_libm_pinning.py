"""How far is the oracle's libm convention from what rustc-built lyon calls on THIS host?

lyon's `f32::{sin, cos, acos, asin, log2}` reach the platform libm (`sinf`, `cosf`, ...).  The oracle and the device both
evaluate them in double precision and round once ((float)cos((double)x)), because a GPU cannot run glibc's f32 routines bit
for bit.  glibc's f32 routines are not correctly rounded, so the two conventions differ by one ulp on a fraction of all
inputs (measured exhaustively on glibc 2.39: cos 0.26 %, sin 0.55 %, acos 1.4 %, asin 1.2 % of the floats in the ranges
used).  What matters is where a one-ulp difference can change TOPOLOGY (vertex / index counts):

  acos    circle_flattening_step (stroke.rs:2749-2753): the arc step divides every round join / cap angle
  log2    `num_segments.log2().round()` (stroke.rs:2003-2004, 2489-2490): the subdivision level of an arc
  sincos  find_sharp_turn's rotation (stroke.rs:3178-3212) -> t_split -> Wang counts of the two halves
  asin    the side-line tilt of variable-width joins (stroke.rs:1817): geometry only

This test records every libm call the oracle makes on the parity configurations (oracle_libm_trace_*), evaluates both
conventions on exactly those arguments and asserts that the count-deciding ones agree; the position-only ones are
reported (they stay within the 1e-5 x extent tolerance of north_star by construction: one ulp of a unit normal).
"""
import ctypes
import ctypes.util

import numpy as np
import pytest

import oracle_ffi as O
import workloads as W
from lyon_b200 import LineCap, LineJoin, PathBatch, StrokeOptions

GOLDEN = O.ROOT / "tests" / "golden"
NAMES = {0: "cos (arc normals)", 1: "sin (arc normals)", 2: "acos (arc step)", 3: "asin (vw tilt)", 4: "log2 (arc level)",
         5: "cos (sharp turn)", 6: "sin (sharp turn)"}
LIBM = ctypes.CDLL(ctypes.util.find_library("m"))
for _n in ("cosf", "sinf", "acosf", "asinf", "log2f"):
    getattr(LIBM, _n).restype = ctypes.c_float
    getattr(LIBM, _n).argtypes = [ctypes.c_float]
F32 = {0: LIBM.cosf, 1: LIBM.sinf, 2: LIBM.acosf, 3: LIBM.asinf, 4: LIBM.log2f, 5: LIBM.cosf, 6: LIBM.sinf}
F64 = {0: np.cos, 1: np.sin, 2: np.arccos, 3: np.arcsin, 4: np.log2, 5: np.cos, 6: np.sin}


def traced(batch, opts):
    L = O.lib()
    L.oracle_libm_trace_end.restype = ctypes.c_uint64
    L.oracle_libm_trace_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.oracle_libm_trace_begin()
    O.stroke_batch(batch, opts)
    n = L.oracle_libm_trace_end()
    fn = np.zeros(n, np.uint8)
    arg = np.zeros(n, np.float32)
    L.oracle_libm_trace_copy(fn.ctypes.data, arg.ctypes.data)
    return fn, arg


def mismatches(fn, arg):
    """per function code: (calls, distinct arguments, distinct arguments where glibc f32 != double rounded once)"""
    out = {}
    for code in np.unique(fn):
        a = np.unique(arg[fn == code])
        with np.errstate(invalid="ignore", divide="ignore"):
            dbl = F64[int(code)](a.astype(np.float64)).astype(np.float32)
        f32 = np.array([F32[int(code)](float(x)) for x in a], np.float32)
        diff = ~((dbl == f32) | (np.isnan(dbl) & np.isnan(f32)))
        out[int(code)] = (int((fn == code).sum()), len(a), int(diff.sum()), a[diff])
    return out


def configs():
    rnd = StrokeOptions(line_width=2.0, line_join=LineJoin.Round, tolerance=0.1).with_line_cap(LineCap.Round)
    yield "C1 logo, round", PathBatch.from_paths(W.fixture_paths(GOLDEN)), StrokeOptions(tolerance=0.05, line_join=LineJoin.Round).with_line_cap(LineCap.Round)
    yield "C3 20k paths", W.c3_mixed_paths(20000, 64, seed=2024), rnd
    yield "C4 300 polylines, variable width", W.c4_polylines(300, 1000), StrokeOptions(
        line_join=LineJoin.Miter, miter_limit=4.0).with_variable_line_width(0)
    yield "adversarial small paths", W.random_small_paths(400, 123), rnd


@pytest.mark.parametrize("name,batch,opts", list(configs()), ids=[c[0] for c in configs()])
def test_count_deciding_libm_calls_agree_with_platform_f32(name, batch, opts):
    fn, arg = traced(batch, opts)
    assert len(fn) > 0
    res = mismatches(fn, arg)
    report = []
    for code, (calls, distinct, bad, which) in sorted(res.items()):
        report.append(f"{NAMES[code]:20s} calls {calls:9d} distinct args {distinct:8d} glibc-f32 != double-rounded {bad:7d}"
                      f" ({100.0 * bad / max(distinct, 1):.3f} %)")
    print(f"\n{name}\n  " + "\n  ".join(report))
    # count-deciding: acos and log2 must agree exactly on every argument that occurs
    for code in (2, 4):
        if code in res:
            calls, distinct, bad, which = res[code]
            if code == 4:  # only round(log2(n)) is used: compare after rounding
                a = np.unique(arg[fn == code])
                with np.errstate(invalid="ignore", divide="ignore"):
                    r32 = np.round(np.array([LIBM.log2f(float(x)) for x in a], np.float32))
                    r64 = np.round(np.log2(a.astype(np.float64)).astype(np.float32))
                bad = int((~((r32 == r64) | (np.isnan(r32) & np.isnan(r64)))).sum())
            assert bad == 0, f"{NAMES[code]}: {bad} arguments where the platform f32 routine would change a vertex count: {which[:5]}"


@pytest.mark.parametrize("name,batch,opts", list(configs()), ids=[c[0] for c in configs()])
def test_whole_oracle_under_platform_f32_libm_has_the_same_topology(name, batch, opts):
    """The oracle run twice -- transcendental calls in double rounded once (the convention the device shares) and through
    the platform's f32 libm (what rustc-built lyon calls here): identical vertex count, index count and index array;
    positions / normals differ by at most a few ulp of a unit normal (well inside north_star's 1e-5 x extent)."""
    L = O.lib()
    want = O.stroke_batch(batch, opts)
    L.oracle_set_libm_mode(1)
    try:
        got = O.stroke_batch(batch, opts)
    finally:
        L.oracle_set_libm_mode(0)
    assert len(got.vertices) == len(want.vertices) and len(got.indices) == len(want.indices), name
    assert np.array_equal(got.indices, want.indices), name
    ext = float(np.ptp(batch.points[:, :2], axis=0).max())
    dpos = float(np.nanmax(np.abs(got.positions() - want.positions()))) if len(want.vertices) else 0.0
    nbytes = int((got.vertices.view(np.uint8).reshape(len(got.vertices), -1) != want.vertices.view(np.uint8).reshape(len(want.vertices), -1)).any(axis=1).sum())
    print(f"\n{name}: {len(want.vertices)} vertices, {nbytes} differ in some byte, max |dpos| {dpos:.3e} (tolerance {1e-5 * ext:.3e})")
    assert dpos <= 1e-5 * ext


def test_round_log2_is_libm_independent_below_2_pow_16():
    """`num_segments.log2().round()` over every integer-valued argument the stroker can produce below 2^16: the platform
    log2f, log2 in double rounded once, and the device's integer formulation (round_log2_segments, exercised by the
    emulation tests) all agree."""
    n = np.arange(1, 65536, dtype=np.float32)
    r64 = np.round(np.log2(n.astype(np.float64)).astype(np.float32))
    r32 = np.array([round(float(LIBM.log2f(float(x)))) for x in n], np.float32)
    assert np.array_equal(r64, r32)


def test_arc_step_of_the_parity_configurations_is_libm_independent():
    """circle_flattening_step for every (half width, tolerance) the test-suite and the benchmark use."""
    for hw in (0.15, 0.25, 0.5, 1.0, 1.5, 2.0):
        for tol in (1.0, 0.5, 0.3, 0.1, 0.05, 0.03, 0.02, 0.01, 0.003, 0.001):
            t = np.float32(min(tol, hw))
            a = np.float32((np.float32(hw) - t) / np.float32(hw))
            assert np.float32(np.arccos(np.float64(a))) == np.float32(LIBM.acosf(float(a))), (hw, tol)
