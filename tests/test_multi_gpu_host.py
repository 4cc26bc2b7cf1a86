"""Multi-rank host logic on CPU: two `gloo` ranks shard a batch by cumulative segment count, each runs the stroke on
its own slice, the ranks all-gather {V_r, I_r} and rebase their indices; the concatenation must be bit-for-bit the
single-rank result.  The per-slice compute is the CPU emulation of the device code (tests/emul,
test infrastructure) -- the thing under test here is lyon_b200.sharding, which is what bench.py and a multi-GPU caller
run around the CUDA library."""
import os
import socket
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from lyon_b200 import LineCap, LineJoin, StrokeOptions  # noqa: E402
from lyon_b200.sharding import exchange_totals, partition_by_cost  # noqa: E402


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch():
    import workloads as W
    a = W.c3_mixed_paths(40, 16, seed=5)
    b = W.random_small_paths(60, seed=9)
    return W.concat_batches([a, b, W.c3_mixed_paths(7, 64, seed=6)])


def _opts():
    return StrokeOptions(line_width=2.0, line_join=LineJoin.Round, tolerance=0.1).with_line_cap(LineCap.Round)


def _worker(rank: int, world: int, port: int, out_dir: str):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import emul_ffi as E
        batch = _batch()
        cuts = partition_by_cost(batch.path_verb_begin, world)
        mine = batch.slice(int(cuts[rank]), int(cuts[rank + 1]))
        # count pass (sizes), the one collective, emit with the global vertex base
        sizes = E.stroke_batch(mine, _opts(), pipeline="tile")
        vbase, ibase, totals = exchange_totals(len(sizes.vertices), len(sizes.indices))
        assert totals.shape == (world, 2)
        got = E.stroke_batch(mine, _opts(), vertex_base=vbase, pipeline="tile")
        # the single-pass flow of bench.py at N>1: local ids, then the rank's indices rebased by its vertex base
        assert np.array_equal(sizes.indices + np.uint32(vbase), got.indices)
        assert sizes.vertices.tobytes() == got.vertices.tobytes()
        np.save(os.path.join(out_dir, f"v{rank}.npy"), got.vertices)
        np.save(os.path.join(out_dir, f"i{rank}.npy"), got.indices)
        np.save(os.path.join(out_dir, f"b{rank}.npy"), np.array([vbase, ibase, cuts[rank], cuts[rank + 1]]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_stroke_matches_single_rank(world):
    import torch.multiprocessing as mp
    import emul_ffi as E
    E.lib()  # build once before forking the ranks
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d), nprocs=world, join=True)
        vs = [np.load(os.path.join(d, f"v{r}.npy")) for r in range(world)]
        is_ = [np.load(os.path.join(d, f"i{r}.npy")) for r in range(world)]
        bs = [np.load(os.path.join(d, f"b{r}.npy")) for r in range(world)]
    whole = E.stroke_batch(_batch(), _opts(), pipeline="tile")
    assert np.concatenate(vs).tobytes() == whole.vertices.tobytes()
    assert (np.concatenate(is_) == whole.indices).all()
    # bases are the exclusive prefix sums of the per-rank totals, ranges tile the batch
    assert bs[0][0] == 0 and bs[0][1] == 0 and bs[0][2] == 0 and bs[-1][3] == _batch().n_paths
    for r in range(1, world):
        assert bs[r][0] == bs[r - 1][0] + len(vs[r - 1])
        assert bs[r][1] == bs[r - 1][1] + len(is_[r - 1])
        assert bs[r][2] == bs[r - 1][3]


def test_partition_by_cost_balances_segments():
    import workloads as W
    batch = W.concat_batches([W.c3_mixed_paths(50, 64, seed=1), W.c3_mixed_paths(400, 8, seed=2)])
    for world in (1, 2, 4, 8):
        cuts = partition_by_cost(batch.path_verb_begin, world)
        assert cuts[0] == 0 and cuts[-1] == batch.n_paths and (np.diff(cuts) >= 0).all()
        cost = np.diff(batch.path_verb_begin[cuts].astype(np.int64))
        assert cost.sum() == batch.path_verb_begin[-1]
        longest = int(np.diff(batch.path_verb_begin.astype(np.int64)).max())
        assert cost.max() - cost.min() <= 2 * longest  # never off by more than one path either way


def test_partition_more_ranks_than_paths():
    import workloads as W
    batch = W.c3_mixed_paths(3, 4, seed=3)
    cuts = partition_by_cost(batch.path_verb_begin, 8)
    assert cuts[0] == 0 and cuts[-1] == 3 and (np.diff(cuts) >= 0).all() and len(cuts) == 9


def test_exchange_totals_without_process_group():
    vb, ib, totals = exchange_totals(10, 30)
    assert (vb, ib) == (0, 0) and totals.tolist() == [[10, 30]]
