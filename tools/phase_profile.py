#!/usr/bin/env python
"""Per-phase cycle breakdown of the stroke tile kernel on a C3 sample (needs a -DLYON_PHASE_TIMING build, dev_build.sh).

  LYON_CUDA_LIB=lyon_b200/liblyon_cuda_phase.so python tools/phase_profile.py [n_paths]
  LYON_PIPE=ctile ...   the CTA-tile kernel (k_stroke_tile) instead of the warp-tile kernel (k_stroke_wtile)
"""
import ctypes
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import workloads as W  # noqa: E402
from lyon_b200 import DeviceTessellator, LineCap, LineJoin, StrokeOptions  # noqa: E402

NAMES_CTILE = ["0 header", "1a verbs + split", "1c merge filter", "-", "-", "3 classify", "3 dense count", "4 scan + chains", "5 lookback",
               "6 fast emit", "6 full/end emit", "0 stage (TMA wait)"]
NAMES_WTILE = ["0 ticket+header", "0 stage (TMA)", "1 setup (verbs)", "2 gen (slots)", "3 walk", "4 classify", "4 dense count",
               "5 scan", "6 lookback", "7 fast emit", "7 full emit", "7 end records"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    names = NAMES_CTILE if os.environ.get("LYON_PIPE") == "ctile" else NAMES_WTILE
    tz = DeviceTessellator(0)
    batch = W.c3_batch_range(0, n, 64)
    opts = StrokeOptions(line_width=2.0, line_join=LineJoin.Round, tolerance=0.1).with_line_cap(LineCap.Round)
    d = tz.upload(batch)
    tz.enable_timing(True)
    for _ in range(2):
        v, i = tz.stroke_fused(d, opts)
    torch.cuda.synchronize()
    out = (ctypes.c_uint64 * 16)()
    tz.L.lyon_cuda_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
    tz.L.lyon_cuda_debug_phase_cycles(tz.ctx, out)
    v, i = tz.stroke_fused(d, opts)
    torch.cuda.synchronize()
    ms = tz.last_kernel_ms()
    tz.L.lyon_cuda_debug_phase_cycles(tz.ctx, out)
    tot = sum(out[:12])
    print(f"paths {n} vertices {v.shape[0]} kernel {ms:.3f} ms")
    for k, name in enumerate(names):
        print(f"  {name:18s} {out[k] / tot * 100:6.2f} %   {out[k]:>14d} cycles")
    if names is NAMES_CTILE:
        print("  tiles by number of non-fast joins: <=480 %d, <=512 %d, <=544 %d, >544 %d" % tuple(out[12:16]))


if __name__ == "__main__":
    main()
