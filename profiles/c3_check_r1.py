import sys, time, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, torch, bang_b200
from bang_b200 import builder, formats, api, recall
import oracle as O
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
prefix = f"/tmp/bc/gist_{n}"
t0 = time.time()
info = builder.make_fixture_auto(prefix, n, 960, "float", 10000, None, device=dev, L_build=64)
print("built", round(time.time()-t0,1), info, flush=True)
q = formats.read_bin(prefix + "_query.bin", np.float32)
gi, gd = formats.read_truthset(prefix + "_gt.bin")
s = api.BANGSearch("float", "exact")
t0 = time.time(); assert s.bang_load(prefix), s.last_error; print("load", round(time.time()-t0,2), s.info().device_bytes>>20, "MiB")
s.set_dists_layout(1)
for L in (10, 16, 24, 32, 48, 64, 96):
    s.bang_set_searchparams(10, L); s.bang_alloc(len(q))
    best = 1e9
    for rep in range(2):
        s.bang_init(len(q)); ids, d = s.bang_query(q); best = min(best, s.last_timing().kernel_ms)
    st = s.last_stats(len(q)); tm = s.last_timing()
    bq = api.algorithmic_bytes(st, "exact", 960, 4, 0, 10).mean()
    print(f"  L={L:4d} recall {recall.calculate_recall(gi, gd, ids, 10):6.2f} kernel {best:8.3f} ms QPS {len(q)/best*1e3:9.0f} hops {st['hops'].mean():6.1f} cand {st['n_cand'].mean():7.1f} B/q {bq:9.0f} GB/s {bq*len(q)/best/1e6:7.1f} grid {tm.grid} block {tm.block} smem {tm.smem_bytes}", flush=True)
    s.bang_free()
# parity on a sample vs oracle at L=32
ox = O.OracleIndex.from_files(prefix, with_pq=False)
s.bang_set_searchparams(10, 32); s.bang_alloc(256); s.bang_init(256)
ids, d = s.bang_query(q[:256])
oi, od = ox.search(q[:256], 10, 32, mode=O.MODE_EXACT)
print("parity vs oracle on 256 queries:", np.array_equal(ids, oi), np.array_equal(d, od))
