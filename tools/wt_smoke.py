#!/usr/bin/env python
"""Small batches through the automatic pipeline (warp tiles) against the oracle; prints what differs.  For
compute-sanitizer runs and first checks of a new kernel build:  LYON_CUDA_LIB=... python tools/wt_smoke.py [n_paths]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import oracle_ffi as O  # noqa: E402
import workloads as W  # noqa: E402
from lyon_b200 import DeviceTessellator, LineCap, LineJoin, StrokeOptions, device_mesh_to_host  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
verbose = len(sys.argv) > 2
tz = DeviceTessellator(0)
bad = 0
cases = [("c3", W.c3_mixed_paths(n, 64, seed=7, closed_fraction=0.3)), ("small", W.random_small_paths(4 * n, 3)),
         ("poly", W.c4_polylines(max(n // 10, 4), 200, with_attribute=False))]
for name, batch in cases:
    for join in (LineJoin.Round, LineJoin.Miter, LineJoin.MiterClip, LineJoin.Bevel):
        for cap in (LineCap.Round, LineCap.Butt, LineCap.Square):
            o = StrokeOptions(line_width=2.0, line_join=join, tolerance=0.1).with_line_cap(cap)
            d = tz.upload(batch)
            for fused in (True, False):
                try:
                    v, i = tz.stroke_fused(d, o) if fused else tz.stroke(d, o)
                    torch.cuda.synchronize()
                    tz.check()
                except Exception as e:  # noqa: BLE001
                    print(name, join, cap, "fused" if fused else "plan+emit", "ERROR", e)
                    bad += 1
                    continue
                got = device_mesh_to_host(v, i)
                want = O.stroke_batch(batch, o)
                ok = (len(got.vertices) == len(want.vertices) and len(got.indices) == len(want.indices)
                      and np.array_equal(got.indices, want.indices) and got.vertices.tobytes() == want.vertices.tobytes())
                if not ok:
                    bad += 1
                    msg = f"nv {len(got.vertices)}/{len(want.vertices)} ni {len(got.indices)}/{len(want.indices)}"
                    if len(got.indices) == len(want.indices):
                        d_i = np.nonzero(got.indices != want.indices)[0]
                        msg += f" first index diff at {d_i[:3]}" if len(d_i) else " indices equal"
                    if len(got.vertices) == len(want.vertices):
                        gb = got.vertices.view(np.uint8).reshape(len(got.vertices), -1)
                        wb = want.vertices.view(np.uint8).reshape(len(want.vertices), -1)
                        d_v = np.nonzero((gb != wb).any(axis=1))[0]
                        msg += f" vertex diffs {len(d_v)} first {d_v[:5]}"
                        gf = got.vertices.view(np.uint32).reshape(len(got.vertices), 8)
                        wf = want.vertices.view(np.uint32).reshape(len(want.vertices), 8)
                        msg += " per-field " + str((gf != wf).sum(axis=0).tolist())
                        if verbose:
                            for q in d_v[:4]:
                                print("   v", q, "got", got.vertices[q], "want", want.vertices[q])
                    print(name, join, cap, "fused" if fused else "plan+emit", "MISMATCH", msg)
print("wt_smoke:", "OK" if bad == 0 else f"{bad} failures", "launches", tz.launch_count())
sys.exit(1 if bad else 0)
