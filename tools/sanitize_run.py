#!/usr/bin/env python
"""A small adversarial workload for compute-sanitizer (memcheck / racecheck / synccheck): every entry point once.

  compute-sanitizer --tool racecheck python tools/sanitize_run.py
The tile kernel aliases shared-memory regions across phases (TMA-staged inputs -> offsets / lists / join codes) and mixes
async-proxy (cp.async.bulk) and generic-proxy accesses; racecheck watches exactly those hand-overs."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import workloads as W  # noqa: E402
from lyon_b200 import DeviceTessellator, LineCap, LineJoin, StrokeOptions, device_mesh_to_host  # noqa: E402

tz = DeviceTessellator(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
batches = [W.random_small_paths(n, seed=3), W.c3_mixed_paths(max(n // 4, 16), 64, seed=5, closed_fraction=0.3)]
total = 0
for batch in batches:
    d = tz.upload(batch)
    for o in (StrokeOptions(line_width=2.0, line_join=LineJoin.Round, tolerance=0.1).with_line_cap(LineCap.Round),
              StrokeOptions(line_width=1.0, line_join=LineJoin.MiterClip, tolerance=0.02).with_line_cap(LineCap.Square)):
        v, i = tz.stroke_fused(d, o)
        v2, i2 = tz.stroke(d, o)
        torch.cuda.synchronize()
        tz.check()
        assert np.array_equal(device_mesh_to_host(v, i).indices, device_mesh_to_host(v2, i2).indices)
        total += int(v.shape[0])
    tz.flatten_paths(d, 0.05)
    tz.measure_paths(d, 0.05)
wb = W.random_width_paths(max(n // 4, 16), seed=7, num_attributes=1)
v, i = tz.stroke(tz.upload(wb), StrokeOptions(line_join=LineJoin.Miter).with_variable_line_width(0))
ctrl = torch.from_numpy(W.c2_cubics(5000)).cuda()
tz.flatten_curves_fused(ctrl, 4, 0.01)
torch.cuda.synchronize()
tz.check()
print(f"sanitize_run: ok, {total} vertices")
