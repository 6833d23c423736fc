"""Sharding of independent lineage trees across GPUs (one process per GPU).

Trees are independent units: roots are partitioned over ranks by greedy bin packing on
their cell-timepoint count; every descendant stays with its root.  The only cross-tree quantities, the
init_cells_f/r statistics (moma_input.h:675-735), are computed on the whole data set before sharding.
The per-rank log-likelihoods are combined with one all-reduce of n_vec doubles (NCCL on GPUs, gloo in the
CPU tests); prediction outputs stay sharded by ctp range.
"""
import heapq

import numpy as np

from .forest import LineageData


def tree_sizes(data: LineageData):
    """cell-timepoints per root (descendants included)."""
    n = np.diff(data.cell_offset).astype(np.int64)
    root_of = np.arange(data.n_cells)
    # parents precede or follow daughters arbitrarily: resolve by pointer jumping
    par = data.parent.astype(np.int64)
    cur = np.where(par >= 0, par, root_of)
    while True:
        nxt = np.where(par[cur] >= 0, par[cur], cur)
        if np.array_equal(nxt, cur):
            break
        cur = nxt
    roots = data.roots()
    size = np.zeros(data.n_cells, dtype=np.int64)
    np.add.at(size, cur, n)
    return roots, size[roots]


def partition_roots(data: LineageData, world_size: int):
    """list (one entry per rank) of root indices; largest-first greedy packing, ties by file order."""
    roots, size = tree_sizes(data)
    order = np.argsort(-size, kind="stable")
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    parts = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        parts[r].append(int(roots[i]))
        heapq.heappush(heap, (load + int(size[i]), r))
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def shard(data: LineageData, rank: int, world_size: int):
    """(LineageData of this rank's trees, cell indices, ctp indices into the full data set)."""
    if world_size == 1:
        return data, np.arange(data.n_cells), np.arange(data.n_ctp)
    parts = partition_roots(data, world_size)
    return data.subset(parts[rank])


def allreduce_loglik(local_ll, group=None):
    """sum of per-rank log-likelihood tensors (torch.distributed; NCCL for CUDA tensors, gloo for CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local_ll, op=dist.ReduceOp.SUM, group=group)
    return local_ll
