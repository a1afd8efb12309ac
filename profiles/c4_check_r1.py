"""C4 DEEP100M-shape (N = 1e8, D = 96 fp32, R = 64, PQ m = 32, Inmemory mode) on one B200: build on the GPU, write the
reference's files, load through bang_load, sweep L.   usage: python profiles/c4_check_r1.py [N]"""
import sys, time, os, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, bang_b200
from bang_b200 import builder, formats, api, recall
dev = torch.device("cuda", 0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
d_ = "/tmp/bang_c4"
os.makedirs(d_, exist_ok=True)
prefix = f"{d_}/deep_{n}"
t0 = time.time()
info = builder.make_fixture_auto(prefix, n, 96, "float", 10000, 32, device=dev, L_build=64)
print("built", round(time.time() - t0, 1), info, flush=True)
torch.cuda.empty_cache()
q = formats.read_bin(prefix + "_query.bin", np.float32)
gi, gd = formats.read_truthset(prefix + "_gt.bin")
s = api.BANGSearch("float", "inmemory")
t0 = time.time(); assert s.bang_load(prefix), s.last_error
print("bang_load", round(time.time() - t0, 1), "s;", s.info().device_bytes >> 20, "MiB in HBM", flush=True)
s.set_dists_layout(1)
for L in (16, 32, 48, 64, 96, 128, 176, 256):
    s.bang_set_searchparams(10, L); s.bang_alloc(len(q))
    best = 1e9
    for rep in range(3):
        s.bang_init(len(q)); ids, dd = s.bang_query(q); best = min(best, s.last_timing().kernel_ms)
    st = s.last_stats(len(q))
    bq = api.algorithmic_bytes(st, "inmemory", 96, 4, 32, 10).mean()
    r = recall.calculate_recall(gi, gd, ids, 10)
    print(f"  L={L:4d} recall {r:6.2f} kernel {best:8.3f} ms QPS {len(q)/best*1e3:9.0f} hops {st['hops'].mean():6.1f} cand {st['n_cand'].mean():7.1f} B/q {bq:9.0f} GB/s {bq*len(q)/best/1e6:7.1f}", flush=True)
    s.bang_free()
    if r >= 99.0: break
s.bang_unload()
for f in os.listdir(d_):
    os.remove(os.path.join(d_, f))
