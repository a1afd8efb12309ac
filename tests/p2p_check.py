"""Run under torchrun on G >= 2 GPUs of one box: graph rows sharded id % G across the ranks' HBM, peers' rows
read with P2P loads inside the traversal kernel (replaces BANG_Base's host-RAM graph).  Every rank
searches all queries and must reproduce the oracle bit for bit.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/p2p_check.py
"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bang_b200  # noqa: E402,F401
from bang_b200 import api, formats, sharding  # noqa: E402
import oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    z = np.load(os.path.join(ROOT, "tests", "golden", "fx_u8.npz"))
    tmp = tempfile.mkdtemp(prefix=f"p2p_r{rank}_")
    prefix = os.path.join(tmp, "fx_u8")
    formats.write_index(prefix, z["base"], z["deg"], z["nbrs"], int(z["medoid"]), z["pivots"], z["centroid"],
                        z["chunk_offsets"], z["codes"])
    ox = O.OracleIndex(formats.pack_disk_bin(z["base"], z["deg"], z["nbrs"]), "uint8", z["base"].shape[1], 64,
                       int(z["medoid"]), z["codes"], z["pivots"], z["centroid"], z["chunk_offsets"])
    ok = True
    for mode, om in (("inmemory", O.MODE_INMEMORY), ("base", O.MODE_BASE), ("exact", O.MODE_EXACT)):
        s = api.BANGSearch("uint8", mode, device=local)
        s.set_sharding(rank, world)
        assert s.bang_load(prefix), s.last_error
        full = z["base"].shape[0] * 160  # bytes of the unsharded rows (256 + 32 -> 288 B/row here)
        sharding.exchange_shards(s, rank, world)
        s.set_dists_layout(api.DISTS_QUERY_MAJOR)
        s.bang_set_searchparams(10, 40)
        Q = len(z["queries"])
        s.bang_alloc(Q)
        s.bang_init(Q)
        ids, d = s.bang_query(z["queries"])
        oids, od = ox.search(z["queries"], 10, 40, mode=om)
        same = bool(np.array_equal(ids, oids) and np.array_equal(d, od))
        ok = ok and same
        print(f"rank {rank} mode {mode}: shard {s.info().device_bytes} B in HBM, ids == oracle: {same}", flush=True)
        dist.barrier()  # keep peers' memory alive until every rank is done with it
        s.bang_free()
        s.bang_unload()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
