#!/usr/bin/env python
"""Run the fused curve flattener on C2 (10M random cubics, tol 0.01) a few times (profiling target: ncu -k regex:k_flat_emit)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402
import workloads as W  # noqa: E402
from lyon_b200 import DeviceTessellator  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
tz = DeviceTessellator(0)
d = torch.from_numpy(W.c2_cubics(n)).cuda()
tz.enable_timing(True)
for _ in range(4):
    off, pts = tz.flatten_curves_fused(d, 4, 0.01)
torch.cuda.synchronize()
print(f"curves {n} points {pts.shape[0]} kernel {tz.last_kernel_ms():.3f} ms")
