"""World-size-2 checks of the multi-GPU host logic on CPU (gloo, 127.0.0.1): batch split, slowest-rank timing,
result gather, and the shard-handle exchange protocol (with a stand-in for the CUDA IPC calls)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bang_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeSearch:
    def __init__(self, rank):
        self.rank = rank
        self.imported = {}
        self.imported_fd = {}

    def export_shard(self):
        return bytes([self.rank]) * 64

    def import_shard(self, shard, handle):
        assert len(handle) == 64
        self.imported[shard] = handle

    # VMM scheme (the default): the shard travels as an open file descriptor + its size
    def export_shard_fd(self):
        import tempfile
        self._f = tempfile.TemporaryFile()
        self._f.write(bytes([self.rank]) * 64)
        self._f.flush()
        return os.dup(self._f.fileno()), 4096 + self.rank

    def import_shard_fd(self, shard, fd, nbytes):
        assert nbytes == 4096 + shard
        self.imported_fd[shard] = os.pread(fd, 64, 0)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q = 7
        sl = sharding.rank_batch(q, rank)
        mine = np.arange(world * q)[sl].reshape(-1, 1).astype(np.uint64)
        allr = sharding.gather_rows(mine)
        t = sharding.max_over_ranks([1.0 + rank, 5.0 - rank])
        fs = _FakeSearch(rank)
        sharding.exchange_shards(fs, rank, world)            # default scheme: VMM shards as file descriptors
        assert sorted(fs.imported_fd) == [1 - rank] and fs.imported_fd[1 - rank] == bytes([1 - rank]) * 64 and not fs.imported
        os.environ["BANG_B200_SHARD_VMM"] = "0"              # CUDA-IPC scheme: 64-byte handles
        sharding.exchange_shards(fs, rank, world)
        del os.environ["BANG_B200_SHARD_VMM"]
        # file descriptors between the ranks (the transport of VMM shards): every rank shares an open temp file
        import tempfile
        with tempfile.TemporaryFile() as f:
            f.write(b"shard of rank %d" % rank)
            f.flush()
            got = sharding.exchange_fds(f.fileno(), (1000 + rank).to_bytes(8, "little"), rank, world)
        seen = {}
        for r, item in enumerate(got):
            if r == rank:
                assert item is None
                continue
            pfd, payload = item
            seen[r] = (os.pread(pfd, 64, 0).decode(), int.from_bytes(payload, "little"))
            os.close(pfd)
        out[rank] = (allr.ravel().tolist(), t, sorted(fs.imported), [fs.imported[k][0] for k in sorted(fs.imported)], seen)
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for rank in range(world):
        allr, t, imp, first_bytes, seen = out[rank]
        assert seen == {1 - rank: (f"shard of rank {1 - rank}", 1000 + 1 - rank)}   # the peer's descriptor arrived and is readable
        assert allr == list(range(14))              # rank-ordered concatenation of the per-rank batches
        assert t == [2.0, 5.0]                      # slowest rank
        assert imp == [1 - rank] and first_bytes == [1 - rank]   # imported exactly the peer's shard


def test_split_batch_is_balanced_and_covers():
    for n, w in [(10000, 8), (10, 3), (7, 8), (1, 2)]:
        sl = [sharding.split_batch(n, r, w) for r in range(w)]
        sizes = [s.stop - s.start for s in sl]
        assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
        assert sl[0].start == 0 and all(sl[i].stop == sl[i + 1].start for i in range(w - 1))
