"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed).  No collective sits on the data
path: queries are independent.

* replicated index (<= 100 M points): every rank loads the whole index and searches its own batch;
* sharded graph (1 B points): rank r keeps rows with id % G == r; the 64-byte CUDA IPC handles of the shards
  are exchanged once (all_gather_object) and imported, after which the traversal kernel reads peers' rows
  with P2P loads over NVLink (replaces BANG_Base's host-RAM graph, bang_search.cu:709-845).
"""
from __future__ import annotations

import numpy as np


def rank_batch(n_queries_per_rank: int, rank: int) -> slice:
    """Weak scaling: rank r searches queries [r*q, (r+1)*q) of a file holding world*q queries."""
    return slice(rank * n_queries_per_rank, (rank + 1) * n_queries_per_rank)


def split_batch(n_queries: int, rank: int, world: int) -> slice:
    """Strong scaling: contiguous, balanced split of one batch (the first n % world ranks get one more)."""
    base, extra = divmod(n_queries, world)
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


def max_over_ranks(values, device=None):
    """Element-wise max over ranks of a list of floats (timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def gather_rows(a: np.ndarray) -> np.ndarray:
    """Concatenate per-rank result arrays in rank order on every rank (final top-k gather)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return a
    outs = [None] * dist.get_world_size()
    dist.all_gather_object(outs, a)
    return np.concatenate(outs, 0)


def exchange_shards(search, rank: int, world: int) -> None:
    """After bang_load on every rank with set_sharding(rank, world): import every peer's graph shard."""
    import torch.distributed as dist
    if world == 1:
        return
    handles = [None] * world
    dist.all_gather_object(handles, search.export_shard())
    for r, h in enumerate(handles):
        if r != rank:
            search.import_shard(r, h)
