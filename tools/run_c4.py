#!/usr/bin/env python
"""C4 (BASELINE.json configs[3]): 100k polylines x 1k segments, miter joins (limit 4), butt caps -- fixed width through the
tile pipeline and variable width (attribute 0) through the thread-per-path kernels.  Prints device times."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import workloads as W  # noqa: E402
from lyon_b200 import DeviceTessellator, LineJoin, StrokeOptions  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
tz = DeviceTessellator(0)
o = StrokeOptions(line_join=LineJoin.Miter, miter_limit=4.0)
for vw in (False, True):
    batch = W.c4_polylines(n, 1000, with_attribute=vw)
    oo = o.with_variable_line_width(0) if vw else o
    d = tz.upload(batch)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for rep in range(3):
        ev[0].record()
        v, i = tz.stroke_fused(d, oo)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1])
    tz.check()
    nv, ni = int(v.shape[0]), int(i.shape[0])
    gb = (32 * nv + 4 * ni + d.input_bytes()) / 1e9
    print(f"C4 {'variable' if vw else 'fixed'} width: {n} polylines x 1000, vertices {nv} indices {ni}  {ms:.2f} ms  "
          f"{nv / ms / 1e6:.2f} Gvert/s  {gb / ms * 1e3:.0f} GB/s algorithmic")
    del d, v, i, batch
    torch.cuda.empty_cache()
