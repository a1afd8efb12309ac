"""End-to-end dry run of the sharded (C5) builder's host logic on CPU: two gloo ranks, the host Vamana builder in place
of the GPU one, a recording stand-in for the search library.  The two ranks' rows are then assembled into one index and
searched with the oracle: recall against the builder's own ground truth proves that node ids, row ownership
(id mod G / id div G), the adjacency exchange + merge, the PQ code placement, the medoid and the ground truth all
speak the same numbering — for generation-order ids ("mod") and for partition ownership ("partition")."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N, D, M, Q, NGT = 24_000, 32, 8, 64, 64


class _Recorder:
    """collects what build_and_load hands to bang_b200_load_device_* (CPU pointers in the dry run)"""

    def __init__(self):
        self.rows = {}
        self.codes = None

    @staticmethod
    def _view(ptr, nbytes):
        return np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(ptr)).copy()

    def load_device_begin(self, N_, D_, medoid, piv, cen, offs):
        self.N, self.D, self.medoid, self.piv, self.cen, self.offs = N_, D_, medoid, piv.copy(), cen.copy(), np.asarray(offs).copy()
        self.codes = np.zeros((N_, len(offs) - 1), np.uint8)

    def load_device_rows(self, first, n, vptr, aptr):
        vec = self._view(vptr, n * self.D).reshape(n, self.D)
        adj = self._view(aptr, n * 64 * 4).view(np.int32).reshape(n, 64)
        for i in range(n):
            self.rows[first + i] = (vec[i], adj[i])

    def load_device_codes(self, first_id, n, cptr):
        self.codes[first_id:first_id + n] = self._view(cptr, n * self.codes.shape[1]).reshape(n, -1)

    def load_device_codes_at(self, iptr, n, cptr):
        ids = self._view(iptr, n * 4).view(np.int32)
        self.codes[ids] = self._view(cptr, n * self.codes.shape[1]).reshape(n, -1)

    def load_device_end(self):
        pass


def _cpu_build(vec, entry, L, passes, seed):
    from bang_b200 import builder
    deg, nbrs, _ = builder.build_vamana_cpu(vec.numpy(), L=max(L, 64), alpha=1.2, seed=seed, nthreads=2, passes=passes)
    out = nbrs.astype(np.int64)
    out[np.arange(64)[None, :] >= deg[:, None]] = -1
    return torch.from_numpy(out.astype(np.int32))


def _worker(rank, world, port, ownership, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bang_b200  # noqa: F401
        from bang_b200 import build_sharded
        rec = _Recorder()
        my_q, gt_ids, gt_d, medoid, T = build_sharded.build_and_load(rec, N, D, Q // world, NGT, m=M, P_per_rank=2, chunk=10_000, L_build=48,
                                                                     passes=1, ownership=ownership, device="cpu", build_fn=_cpu_build, shard_slack=3.0)
        out[rank] = dict(rows=rec.rows, codes=rec.codes, n_id=rec.N, medoid=medoid, piv=rec.piv, cen=rec.cen, offs=rec.offs, my_q=my_q,
                         my_idx=T["my_idx"], gt_ids=gt_ids, gt_d=gt_d, home=T.get("home"))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("ownership", ["mod", "partition"])
def test_sharded_builder_dry_run(ownership):
    import oracle as O
    from bang_b200 import formats, recall
    world = 2
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ownership, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    n_id = r0["n_id"]
    assert n_id == r1["n_id"] and r0["medoid"] == r1["medoid"] and np.array_equal(r0["codes"], r1["codes"])   # replicated parts agree
    assert n_id == N if ownership == "mod" else N <= n_id <= 2 * N
    # one index from the two ranks' rows: node id = local_row * world + rank
    base = np.zeros((n_id, D), np.uint8)
    nbrs = np.zeros((n_id, 64), np.uint32)
    deg = np.zeros(n_id, np.uint32)
    real = np.zeros(n_id, bool)
    for rank, r in ((0, r0), (1, r1)):
        for row, (vec, adj) in r["rows"].items():
            nid = row * world + rank
            if nid >= n_id:
                continue
            k = int((adj >= 0).sum())
            assert (adj[:k] >= 0).all() and (adj[k:] < 0).all()                # used slots first
            base[nid], deg[nid], nbrs[nid, :k] = vec, k, adj[:k]
            real[nid] |= k > 0
    assert real.sum() == N                                                     # every generated point has a row with edges; holes have none
    used = nbrs[np.arange(64)[None, :] < deg[:, None]]
    assert real[used].all()                                                    # no edge points at a hole
    assert float(deg[real].mean()) > 20
    ox = O.OracleIndex(formats.pack_disk_bin(base, deg, nbrs), "uint8", D, 64, int(r0["medoid"]), r0["codes"], r0["piv"], r0["cen"], r0["offs"])
    # the global query batch in its original order, from the ranks' shares
    queries = np.zeros((Q, D), np.uint8)
    for r in (r0, r1):
        queries[r["my_idx"]] = r["my_q"]
    assert sorted(np.concatenate([r0["my_idx"], r1["my_idx"]]).tolist()) == list(range(Q))
    ids, _ = ox.search(queries[:NGT], 10, 64, mode=O.MODE_EXACT)
    rec = recall.calculate_recall(r0["gt_ids"][:NGT], r0["gt_d"][:NGT], ids, 10)
    assert rec >= 95.0, rec
    # PQ codes sit at the nodes' ids: re-encoding the assembled vectors reproduces them
    from bang_b200 import synth
    want = synth.encode_pq(torch.from_numpy(base), r0["piv"], r0["cen"], r0["offs"]).numpy()
    assert np.array_equal(want[real], r0["codes"][real])
    ids_pq, _ = ox.search(queries[:NGT], 10, 64, mode=O.MODE_INMEMORY)
    assert recall.calculate_recall(r0["gt_ids"][:NGT], r0["gt_d"][:NGT], ids_pq, 10) >= 60.0
    if ownership == "partition":   # most of a query's 10 nearest neighbours live on its home GPU
        home = r0["home"].numpy()[:NGT]
        local = np.mean([(r0["gt_ids"][i, :10] % world == home[i]).mean() for i in range(NGT)])
        assert local > 0.8, local
