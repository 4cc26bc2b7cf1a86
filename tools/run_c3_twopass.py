#!/usr/bin/env python
"""C3 sample through the two-pass route: the tile kernel in COUNT mode (look-back, no emission) and in EMIT mode (bases read
back, no look-back), each timed with the library's CUDA events, next to the single-pass FUSED kernel.  Separates what the
in-order look-back costs from what the tile geometry costs:  LYON_TILE=nt,cap,0 LYON_CUDA_LIB=... python tools/run_c3_twopass.py [n]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import workloads as W  # noqa: E402
from lyon_b200 import DeviceTessellator, LineCap, LineJoin, StrokeOptions  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
tz = DeviceTessellator(0)
batch = W.c3_batch_range(0, n, 64)
opts = StrokeOptions(line_width=2.0, line_join=LineJoin.Round, tolerance=0.1).with_line_cap(LineCap.Round)
d = tz.upload(batch)
tz.enable_timing(True)
cnt, emt, fus = [], [], []
for _ in range(4):
    nv, ni = tz.stroke_plan(d, opts)
    cnt.append(tz.last_kernel_ms())
    v, i = tz.stroke_emit(nv, ni)
    torch.cuda.synchronize()
    emt.append(tz.last_kernel_ms())
    v, i = tz.stroke_fused(d, opts, vertices=v, indices=i)
    torch.cuda.synchronize()
    fus.append(tz.last_kernel_ms())
print(f"paths {n} vertices {nv} count {min(cnt):.3f} ms emit {min(emt):.3f} ms fused {min(fus):.3f} ms")
