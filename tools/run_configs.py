#!/usr/bin/env python
"""The parity-case configs of BASELINE.json timed on the device (not bench lines): C1 single-path latency (logo.path and the
Rust logo of bench/tess), C5 = the two logos x 1000 with per-copy jitter, tolerance sweep 1.0 .. 0.001."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import oracle_ffi as O  # noqa: E402
import workloads as W  # noqa: E402
from lyon_b200 import DeviceTessellator, LineJoin, PathBatch, StrokeOptions  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
tz = DeviceTessellator(0)
fix = W.fixture_paths(GOLDEN)


def timed(d, o, reps=20):
    for _ in range(3):
        v, i = tz.stroke_fused(d, o)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        v, i = tz.stroke_fused(d, o)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, int(v.shape[0]), int(i.shape[0])


print("C1 (single path, wall clock per call incl. launches and the size read-back)")
for name, path, o in (("assets/logo.path w1 tol0.05 miter", fix[0], StrokeOptions(tolerance=0.05)),
                      ("rust logo default (stroke_01)", fix[1], StrokeOptions()),
                      ("rust logo bevel (stroke_02)", fix[1], StrokeOptions(line_join=LineJoin.Bevel)),
                      ("rust logo round (stroke_03)", fix[1], StrokeOptions(line_join=LineJoin.Round))):
    b = PathBatch.from_paths([path])
    ms, nv, ni = timed(tz.upload(b), o, 50)
    t0 = time.perf_counter()
    for _ in range(20):
        O.stroke_batch(b, o)
    cpu_ms = (time.perf_counter() - t0) / 20 * 1e3
    print(f"  {name:36s} {nv:6d} v {ni:6d} i   GPU {ms * 1e3:7.1f} us   CPU oracle (1 thread, via ctypes) {cpu_ms * 1e3:7.1f} us")

print("C5 (logo.path + Rust logo, 1000 copies each-alternating, jittered), stroke width 1 miter butt")
batch = W.c5_logo_copies(fix, 1000)
d = tz.upload(batch)
for tol in (1.0, 0.3, 0.1, 0.03, 0.01, 0.003, 0.001):
    o = StrokeOptions(tolerance=tol)
    ms, nv, ni = timed(d, o, 5)
    gb = (32 * nv + 4 * ni + d.input_bytes()) / 1e9
    print(f"  tol {tol:6.3f}: {nv:10d} v {ni:10d} i  {ms:8.3f} ms  {nv / ms / 1e6:6.2f} Gvert/s  {gb / ms * 1e3:6.0f} GB/s")
tz.check()
