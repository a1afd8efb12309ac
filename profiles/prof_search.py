"""Short driver for ncu captures: builds (or reuses) a synthetic index in /tmp and runs a few searches.
usage: python profiles/prof_search.py [L] [mode] [runs] [shape]     shape: sift1m (default, C2) | deep10m (C4's shape at 10^7 points)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bang_b200
from bang_b200 import builder, formats, api, recall
L = int(sys.argv[1]) if len(sys.argv) > 1 else 152
mode = sys.argv[2] if len(sys.argv) > 2 else "inmemory"
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
shape = sys.argv[4] if len(sys.argv) > 4 else "sift1m"
n, D, dt, npdt, name = {"sift1m": (1_000_000, 128, "uint8", np.uint8, "u8_1m"),
                        "deep10m": (10_000_000, 96, "float", np.float32, "f32_10m")}[shape]
prefix = "/tmp/bang_prof/" + name
if not os.path.exists(prefix + "_gt.bin"):
    print(builder.make_fixture_auto(prefix, n, D, dt, 10000, 32, device=torch.device("cuda", 0)))
q = formats.read_bin(prefix + "_query.bin", npdt)
gi, gd = formats.read_truthset(prefix + "_gt.bin")
s = api.BANGSearch(dt, mode)
assert s.bang_load(prefix)
s.set_dists_layout(1)
s.bang_set_searchparams(10, L)
s.bang_alloc(len(q))
for r in range(runs):
    s.bang_init(len(q))
    ids, d = s.bang_query(q)
    print("run", r, "kernel ms", s.last_timing().kernel_ms, "recall", recall.calculate_recall(gi, gd, ids, 10))
