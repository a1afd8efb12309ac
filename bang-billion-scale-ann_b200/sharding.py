"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed).  No collective sits on the data
path: queries are independent.

* replicated index (<= 100 M points): every rank loads the whole index and searches its own batch;
* sharded graph (1 B points): rank r keeps rows with id % G == r; the shards (CUDA virtual-memory allocations) are
  exchanged once as file descriptors and mapped, after which the traversal kernel reads peers' rows with P2P
  loads over NVLink (replaces BANG_Base's host-RAM graph, bang_search.cu:709-845).
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


def exchange_fds(fd: int, payload: bytes, rank: int, world: int, tag: str | None = None) -> list:
    """All-to-all of one open file descriptor per rank between the processes of one box.  Returns a list with, for
    every peer r != rank, (fd_in_this_process, payload_of_r), and None at index `rank`.  Descriptors cannot travel
    through torch.distributed; they are sent as SCM_RIGHTS ancillary data over Unix-domain sockets (abstract names,
    nothing on disk).  Ranks must already be in a process group (used for the barriers and to agree on the names)."""
    import os
    import socket
    import torch.distributed as dist
    if tag is None:
        box = [None]
        if rank == 0:
            box[0] = f"bang_b200_{os.getpid()}_{os.urandom(4).hex()}"
        dist.broadcast_object_list(box, src=0)
        tag = box[0]
    name = lambda r: "\0" + f"{tag}_{r}"
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(name(rank))
    srv.listen(world)
    dist.barrier()                      # every rank listens before anyone connects
    out = [None] * world
    try:
        for r in range(world):          # send mine to every peer ...
            if r == rank:
                continue
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(name(r))
            hdr = rank.to_bytes(4, "little") + len(payload).to_bytes(4, "little") + payload
            socket.send_fds(c, [hdr], [fd])
            c.close()
        for _ in range(world - 1):      # ... and receive one from each (connections queue in the listen backlog)
            conn, _addr = srv.accept()
            data, fds, _flags, _a = socket.recv_fds(conn, 8 + 4096, 1)
            while len(data) < 8 or len(data) < 8 + int.from_bytes(data[4:8], "little"):
                more = conn.recv(4096)
                if not more:
                    break
                data += more
            conn.close()
            src = int.from_bytes(data[:4], "little")
            n = int.from_bytes(data[4:8], "little")
            out[src] = (fds[0], data[8:8 + n])
    finally:
        dist.barrier()
        srv.close()
    return out


def exchange_shards(search, rank: int, world: int) -> None:
    """After bang_load on every rank with set_sharding(rank, world): import every peer's graph shard.
    Default: the rows are CUDA virtual-memory (VMM) allocations, shared as file descriptors (exchange_fds) — peer
    reads of rows mapped this way run at the speed of local ones, while cudaMalloc + CUDA-IPC mappings have a slow
    mode at some sizes (9 M points on 2 GPUs: 14.5 ms against 8.4 ms, profiles/r2_c5.md).  BANG_B200_SHARD_VMM=0 (set
    before the load) selects the CUDA-IPC scheme: 64-byte handles through all_gather_object."""
    import os
    import torch.distributed as dist
    if world == 1:
        return
    if os.environ.get("BANG_B200_SHARD_VMM", "1") != "0":
        fd, nbytes = search.export_shard_fd()
        got = exchange_fds(fd, nbytes.to_bytes(8, "little"), rank, world)
        os.close(fd)
        for r, item in enumerate(got):
            if r != rank:
                pfd, payload = item
                search.import_shard_fd(r, pfd, int.from_bytes(payload, "little"))
                os.close(pfd)
        return
    handles = [None] * world
    dist.all_gather_object(handles, search.export_shard())
    for r, h in enumerate(handles):
        if r != rank:
            search.import_shard(r, h)
