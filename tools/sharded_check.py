#!/usr/bin/env python
"""Multi-GPU parity of the sharded C entry point (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 tools/sharded_check.py

Every rank strokes its slice with lyon_cuda_stroke_fused_sharded (the library's own NCCL communicator); rank 0 gathers the
per-rank buffers and compares their concatenation with the CPU oracle's mesh of the whole batch."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import workloads as W  # noqa: E402
from lyon_b200 import DeviceTessellator, LineCap, LineJoin, StrokeOptions, device_mesh_to_host  # noqa: E402
from lyon_b200.sharding import partition_by_cost  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # only carries the NCCL id and the results; the data path uses the library's communicator
    batch = W.c3_mixed_paths(3000, 48, seed=77, closed_fraction=0.25)
    o = StrokeOptions(line_width=2.0, line_join=LineJoin.Round, tolerance=0.1).with_line_cap(LineCap.Round)
    cuts = partition_by_cost(batch.path_verb_begin, world)
    mine = batch.slice(int(cuts[rank]), int(cuts[rank + 1]))
    tz = DeviceTessellator(local)
    ids = [tz.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    tz.comm_init(ids[0], rank, world)
    for _ in range(2):  # the second call reuses the buffers
        v, i, base, totals = tz.stroke_fused_sharded(tz.upload(mine), o)
    torch.cuda.synchronize()
    tz.check()
    mesh = device_mesh_to_host(v, i)
    parts = [None] * world
    dist.all_gather_object(parts, (mesh.vertices.tobytes(), mesh.indices.tobytes(), base, totals.tolist()))
    ok = True
    if rank == 0:
        import oracle_ffi as O
        want = O.stroke_batch(batch, o)
        vb = b"".join(p[0] for p in parts)
        ib = np.concatenate([np.frombuffer(p[1], np.uint32) for p in parts])
        ok = vb == want.vertices.tobytes() and np.array_equal(ib, want.indices)
        bases = [p[2] for p in parts]
        ok = ok and bases == [int(sum(t[0] for t in parts[0][3][:q])) for q in range(world)]
        print(f"sharded_check: {'OK' if ok else 'MISMATCH'} world {world} vertices {len(want.vertices)} indices {len(want.indices)} "
              f"bases {bases}")
    tz.comm_destroy()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
