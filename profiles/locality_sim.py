"""How local would the hops be with partition-based row ownership?  (CPU-only estimate.)
Builds a clustered index with the host builder, runs the oracle with a trace of the expanded nodes, and compares two
ownership rules for G = 8 GPUs: rows owned by id mod G (today) vs rows owned by the GPU of their nearest of P = 32
partition centres (k-means), with every query sent to the GPU that owns ITS nearest centre.
  python profiles/locality_sim.py [N=200000] [D=64] [Q=1000] [L=100]"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import bang_b200
from bang_b200 import builder, formats
import oracle as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 64
Q = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
L = int(sys.argv[4]) if len(sys.argv) > 4 else 100
G, P = 8, 32
with tempfile.TemporaryDirectory() as tmp:
    prefix = os.path.join(tmp, "idx")
    t = time.time()
    info = builder.make_fixture_auto(prefix, N, D, "uint8", Q, 16, k_gt=10, device="cpu")
    print("index:", info, f"{time.time() - t:.0f}s", flush=True)
    ox = O.OracleIndex.from_files(prefix)
    q = formats.read_bin(prefix + "_query.bin", np.uint8)
    base = formats.read_bin(prefix + "_base.bin", np.uint8) if os.path.exists(prefix + "_base.bin") else None
    meta = formats.read_disk_metadata(prefix + "_disk_metadata.bin")
    vec, deg, nbrs = formats.read_disk_bin(prefix + "_disk.bin", meta)
    ids, d, st = ox.search(q, 10, L, mode=O.MODE_INMEMORY, stats=True, trace_len=L + 121)
x = torch.from_numpy(vec.astype(np.float32))
g = torch.Generator().manual_seed(1)
cen = x[torch.randperm(N, generator=g)[:P]].clone()
for _ in range(10):
    lab = torch.cdist(x, cen).argmin(1)
    for p in range(P):
        m = lab == p
        if m.any():
            cen[p] = x[m].mean(0)
lab = torch.cdist(x, cen).argmin(1).numpy()
# partitions -> GPUs: balance by size (largest first onto the lightest GPU)
sizes = np.bincount(lab, minlength=P)
gpu_of_part = np.zeros(P, np.int64); load = np.zeros(G)
for p in np.argsort(-sizes):
    gsel = int(np.argmin(load)); gpu_of_part[p] = gsel; load[gsel] += sizes[p]
owner_part = gpu_of_part[lab]
home = gpu_of_part[torch.cdist(torch.from_numpy(q.astype(np.float32)), cen).argmin(1).numpy()]
tr, hops = st["trace"], st["hops"]
loc_mod = loc_part = tot = 0
per_q = []
for i in range(len(q)):
    nodes = tr[i, :min(hops[i], tr.shape[1])].astype(np.int64)
    tot += len(nodes)
    loc_mod += int(((nodes % G) == (i % G)).sum())
    lp = int((owner_part[nodes] == home[i]).sum())
    loc_part += lp
    per_q.append(lp / max(1, len(nodes)))
print(f"N={N} D={D} Q={len(q)} L={L}: hops/query {tot / len(q):.1f}; GPU loads {np.round(load / N, 3).tolist()}")
print(f"local fraction of expanded rows: id mod {G}: {loc_mod / tot:.3f}   partition-owned + query routing: {loc_part / tot:.3f} "
      f"(median per query {np.median(per_q):.3f}, 10th percentile {np.percentile(per_q, 10):.3f})")
print(f"queries per GPU under routing: {np.bincount(home, minlength=G).tolist()}")
