#!/usr/bin/env python
"""C5 at one tolerance (profiling target): the two logo fixtures x n copies with jitter, stroke width 1 miter butt."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import workloads as W  # noqa: E402
from lyon_b200 import DeviceTessellator, StrokeOptions  # noqa: E402

tol = float(sys.argv[1]) if len(sys.argv) > 1 else 0.003
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
tz = DeviceTessellator(0)
batch = W.c5_logo_copies(W.fixture_paths(ROOT / "tests" / "golden"), n)
d = tz.upload(batch)
o = StrokeOptions(tolerance=tol)
for _ in range(3):
    v, i = tz.stroke_fused(d, o)
torch.cuda.synchronize()
l0 = tz.launch_count()
t0 = time.perf_counter()
v, i = tz.stroke_fused(d, o)
torch.cuda.synchronize()
print(f"tol {tol}: {v.shape[0]} vertices, {(time.perf_counter() - t0) * 1e3:.3f} ms, {tz.launch_count() - l0} launches")
