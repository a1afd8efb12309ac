import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, torch
import bang_b200
from bang_b200 import synth, formats, api
import oracle as O
rng = np.random.default_rng(77)
N, D, R, m = 60_000, 32, 64, 8
base = rng.integers(0, 256, size=(N, D), dtype=np.uint8)
r = rng.integers(0, N - 1, size=(N, R))
while True:
    srt = np.sort(r, axis=1)
    dup = np.nonzero((srt[:, 1:] == srt[:, :-1]).any(1))[0]
    if len(dup) == 0:
        break
    r[dup] = rng.integers(0, N - 1, size=(len(dup), R))
nbrs = ((np.arange(N)[:, None] + 1 + r) % N).astype(np.uint32)
deg = np.full(N, R, dtype=np.uint32)
piv, cen, offs = synth.train_pq(torch.from_numpy(base), m, iters=4)
codes = synth.encode_pq(torch.from_numpy(base), piv, cen, offs).numpy()
os.makedirs("/tmp/rnd", exist_ok=True)
prefix = "/tmp/rnd/rnd"
formats.write_index(prefix, base, deg, nbrs, 123, piv, cen, offs, codes)
queries = rng.integers(0, 256, size=(24, D), dtype=np.uint8)
ox = O.OracleIndex(formats.pack_disk_bin(base, deg, nbrs), "uint8", D, R, 123, codes, piv, cen, offs)
MO = {"base": O.MODE_BASE, "inmemory": O.MODE_INMEMORY, "exact": O.MODE_EXACT}
for mode in ["inmemory", "base", "exact"]:
    s = api.BANGSearch("uint8", mode)
    assert s.bang_load(prefix)
    s.set_dists_layout(api.DISTS_QUERY_MAJOR)
    for L in [16, 40, 64, 100, 128, 200, 300, 512]:
        s.bang_set_searchparams(10, L)
        s.bang_alloc(len(queries)); s.bang_init(len(queries))
        ids, dists = s.bang_query(queries)
        st = s.last_stats(len(queries))
        s.bang_free()
        oids, od, ost = ox.search(queries, 10, L, mode=MO[mode], order=O.ORDER_GPU, stats=True)
        bad = (ids != oids).any(1)
        print(mode, L, "differ", int(bad.sum()), "n_cand gpu/oracle", st["n_cand"].mean(), ost["n_cand"].mean(), "hops", st["hops"].mean(), ost["hops"].mean(),
              "first bad", (int(np.argmax(bad)), st["n_cand"][np.argmax(bad)], ost["n_cand"][np.argmax(bad)]) if bad.any() else None, flush=True)
    s.bang_unload()
