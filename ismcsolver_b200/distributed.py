"""Multi-GPU root-parallel plumbing: one process per GPU, trees sharded over ranks, root arrays merged.

Root-parallel trees are independent (reference: one tree per std::async worker, execution.h:193-204),
so the global tree range [0, trees_total) of every position is split contiguously over the ranks; tree
ids are global, which makes the Philox streams — and therefore the merged result — independent of the
number of GPUs.  The only exchange step is the sum of the per-GPU root arrays (visits, 2*score,
availability per move): RootParallel::compileVisitCounts (execution.h:219-231) across devices, one
all-reduce per search call (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_trees(trees_total: int, world: int, rank: int) -> tuple[int, int]:
    """(tree_base, trees) of `rank`: a contiguous, near-equal split of [0, trees_total)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    if trees_total < world:
        raise ValueError("fewer trees than ranks")
    base, extra = divmod(trees_total, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def merge_root_arrays(arrays: torch.Tensor, group=None) -> torch.Tensor:
    """Sums the root arrays (any integer tensor; int64 on device) over all ranks, in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(arrays, op=dist.ReduceOp.SUM, group=group)
    return arrays


def best_moves(visits: torch.Tensor) -> torch.Tensor:
    """RootParallel::bestMove (execution.h:209-216): first maximum of the merged visits in ascending
    move order, per position; -1 where the root is terminal (no visits)."""
    best = torch.argmax(visits, dim=-1)              # argmax returns the first maximal index
    best[visits.sum(dim=-1) == 0] = -1
    return best
