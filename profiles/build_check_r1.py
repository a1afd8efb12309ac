import sys, time, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, torch, bang_b200
from bang_b200 import builder, formats, api, recall, synth
dev = torch.device("cuda", 0)
for n, L_build in ((100_000, 64), (1_000_000, 64)):
    prefix = f"/tmp/bc/u8_{n}"
    t0 = time.time()
    info = builder.make_fixture_auto(prefix, n, 128, "uint8", 10000, 32, device=dev, L_build=L_build)
    print("N", n, "built", round(time.time()-t0,1), info, flush=True)
    q = formats.read_bin(prefix + "_query.bin", np.uint8)
    gi, gd = formats.read_truthset(prefix + "_gt.bin")
    s = api.BANGSearch("uint8", "inmemory")
    t0 = time.time(); assert s.bang_load(prefix), s.last_error; print("load", round(time.time()-t0,2))
    s.set_dists_layout(1)
    for L in (10, 16, 24, 32, 48, 64, 96, 128, 152):
        s.bang_set_searchparams(10, L); s.bang_alloc(len(q))
        best = 1e9
        for rep in range(3):
            s.bang_init(len(q)); ids, d = s.bang_query(q); best = min(best, s.last_timing().kernel_ms)
        st = s.last_stats(len(q)); tm = s.last_timing()
        bq = api.algorithmic_bytes(st, "inmemory", 128, 1, 32, 10).mean()
        print(f"  L={L:4d} recall {recall.calculate_recall(gi, gd, ids, 10):6.2f} kernel {best:8.3f} ms  QPS {len(q)/best*1e3:10.0f} hops {st['hops'].mean():6.1f} cand {st['n_cand'].mean():7.1f} B/q {bq:9.0f} GB/s {bq*len(q)/best/1e6:7.1f} ctas/sm {tm.ctas_per_sm} smem {tm.smem_bytes}", flush=True)
        s.bang_free()
    s.bang_unload()
