"""Multi-GPU sharding of a path batch.

Paths are independent units of work -- the reference creates a fresh ``StrokeBuilderImpl`` per ``tessellate*`` call
(``crates/tessellation/src/stroke.rs:133-134``) and only the sub-paths of ONE path share state (the advancement carry,
``stroke.rs:1359``) -- so a batch is cut into ``world`` contiguous path ranges of equal cumulative segment count and
every rank tessellates its own range.  The single exchange on the path is an all-gather of the per-rank output totals
``{V_r, I_r}`` (16 bytes per rank; NCCL over NVLink on the GPUs, gloo in the CPU tests): rank ``r`` emits (or, after a single-pass emit, rebases) its indices
by ``sum(V_q for q < r)`` -- what ``BuffersBuilder::with_vertex_offset`` does in the reference
(``crates/tessellation/src/geometry_builder.rs:325-329``) -- so that the concatenation of the per-rank buffers is
bit-for-bit the single-GPU result.  Outputs stay resident per GPU.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def partition_by_cost(path_verb_begin: np.ndarray, n_ranks: int) -> np.ndarray:
    """Cut a batch into ``n_ranks`` contiguous path ranges of ~equal cumulative segment (verb) count.
    Returns ``n_ranks + 1`` non-decreasing path indices; rank r owns paths ``[cuts[r], cuts[r + 1])``."""
    path_verb_begin = np.asarray(path_verb_begin)
    n_paths = len(path_verb_begin) - 1
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    total = int(path_verb_begin[-1]) - int(path_verb_begin[0])
    targets = int(path_verb_begin[0]) + (np.arange(1, n_ranks, dtype=np.int64) * total) // n_ranks
    cuts = np.searchsorted(path_verb_begin, targets, side="left")
    cuts = np.minimum(cuts, n_paths)
    return np.concatenate([[0], cuts, [n_paths]]).astype(np.int64)


def exchange_totals(n_vertices: int, n_indices: int, group=None, device=None) -> Tuple[int, int, np.ndarray]:
    """All-gather ``{V_r, I_r}`` over the process group.  Returns ``(vertex_base, index_base, totals[world, 2])`` for
    this rank: the number of vertices / indices emitted by the ranks before it.  Without an initialised process
    group (single GPU) the bases are 0."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 0, 0, np.array([[n_vertices, n_indices]], dtype=np.int64)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = torch.tensor([n_vertices, n_indices], dtype=torch.int64, device=device)
    gathered = torch.zeros(2 * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    g = gathered.cpu().numpy().reshape(world, 2)
    return int(g[:rank, 0].sum()), int(g[:rank, 1].sum()), g


def check_global_vertex_count(totals: np.ndarray) -> None:
    """``TooManyVertices`` for the job as a whole: the last global vertex id must stay <= u32::MAX - 1
    (``geometry_builder.rs:578-581``), whichever rank holds it."""
    from .geometry import TooManyVertices

    if int(np.asarray(totals)[:, 0].sum()) > 0xFFFFFFFF:
        raise TooManyVertices(-2, "TooManyVertices: the batch emits more than u32::MAX vertices across all ranks")


class ShardedStroker:
    """One rank's side of the sharded stroke: ``plan`` (count pass on the local slice) -> ``exchange_totals`` ->
    ``emit`` with the global vertex base.  ``tessellator`` is this rank's ``DeviceTessellator``."""

    def __init__(self, tessellator, group=None, device=None):
        self.tz = tessellator
        self.group = group
        self.device = device
        self.totals: Optional[np.ndarray] = None
        self.vertex_base = 0
        self.index_base = 0

    def run(self, dbatch, options, vertices=None, indices=None):
        nv, ni = self.tz.stroke_plan(dbatch, options)
        self.vertex_base, self.index_base, self.totals = exchange_totals(nv, ni, self.group, self.device)
        check_global_vertex_count(self.totals)
        return self.tz.stroke_emit(nv, ni, vertex_base=self.vertex_base, vertices=vertices, indices=indices)

    def run_single_pass(self, dbatch, options, vertices=None, indices=None):
        """Same result with ONE pass over the slice: the fused count+scan+emit kernel with local ids, the all-gather of
        ``{V_r, I_r}``, then a streaming rebase of this rank's indices (``lyon_cuda_rebase_indices``)."""
        v, i = self.tz.stroke_fused(dbatch, options, vertices=vertices, indices=indices)
        self.vertex_base, self.index_base, self.totals = exchange_totals(int(v.shape[0]), int(i.shape[0]), self.group, self.device)
        check_global_vertex_count(self.totals)
        self.tz.rebase_indices(i, int(v.shape[0]), self.vertex_base)
        return v, i
