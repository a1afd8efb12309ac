"""Short driver for ncu captures: builds (or reuses) the C2 SIFT1M-shape index in /tmp and runs a few searches.
usage: python profiles/prof_search.py [L] [mode] [n_queries_runs]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bang_b200
from bang_b200 import builder, formats, api, recall
L = int(sys.argv[1]) if len(sys.argv) > 1 else 152
mode = sys.argv[2] if len(sys.argv) > 2 else "inmemory"
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
prefix = "/tmp/bang_prof/u8_1m"
if not os.path.exists(prefix + "_gt.bin"):
    print(builder.make_fixture_auto(prefix, 1_000_000, 128, "uint8", 10000, 32, device=torch.device("cuda", 0)))
q = formats.read_bin(prefix + "_query.bin", np.uint8)
gi, gd = formats.read_truthset(prefix + "_gt.bin")
s = api.BANGSearch("uint8", mode)
assert s.bang_load(prefix)
s.set_dists_layout(1)
s.bang_set_searchparams(10, L)
s.bang_alloc(len(q))
for r in range(runs):
    s.bang_init(len(q))
    ids, d = s.bang_query(q)
    print("run", r, "kernel ms", s.last_timing().kernel_ms, "recall", recall.calculate_recall(gi, gd, ids, 10))
