#!/usr/bin/env python
"""Run the fused stroke pipeline on a C3 sample a few times (profiling target: ncu -k regex:k_stroke_tile ...)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import workloads as W  # noqa: E402
from lyon_b200 import DeviceTessellator, LineCap, LineJoin, StrokeOptions  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tz = DeviceTessellator(0)
batch = W.c3_batch_range(0, n, 64)
opts = StrokeOptions(line_width=2.0, line_join=LineJoin.Round, tolerance=0.1).with_line_cap(LineCap.Round)
d = tz.upload(batch)
tz.enable_timing(True)
for _ in range(reps):
    v, i = tz.stroke_fused(d, opts)
torch.cuda.synchronize()
print(f"paths {n} vertices {v.shape[0]} indices {i.shape[0]} kernel {tz.last_kernel_ms():.3f} ms")
