"""Per-phase SM-clock breakdown of the fused search kernel (debug build with -DBANG_PHASE_TIMERS).
usage: python profiles/phase_clocks.py [L] [mode]     (needs bang-billion-scale-ann_b200/libbang_b200_prof.so)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bang_b200
from bang_b200 import builder, formats, api, recall
L = int(sys.argv[1]) if len(sys.argv) > 1 else 152
mode = sys.argv[2] if len(sys.argv) > 2 else "inmemory"
prefix = "/tmp/bang_prof/u8_1m"
if not os.path.exists(prefix + "_gt.bin"):
    print(builder.make_fixture_auto(prefix, 1_000_000, 128, "uint8", 10000, 32, device=torch.device("cuda", 0)))
q = formats.read_bin(prefix + "_query.bin", np.uint8)
lib_path = os.path.join(ROOT, "bang-billion-scale-ann_b200", "libbang_b200_prof.so")
s = api.BANGSearch("uint8", mode, lib_path=lib_path)
assert s.bang_load(prefix)
s.set_dists_layout(1); s.bang_set_searchparams(10, L); s.bang_alloc(len(q))
for r in range(3):
    s.bang_init(len(q)); ids, d = s.bang_query(q)
print("kernel ms", s.last_timing().kernel_ms)
fn = s._lib.bang_b200_debug_phase_clocks
fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
out = np.zeros((len(q), 16), dtype=np.int64)
assert fn(s._h, out.ctypes.data) == 0
names = ["setup", "adjwait", "hash", "bloom", "compact", "codewait", "lut", "scan", "decide", "merge", "unvis", "rerank", "topk", "hops", "merges"]
tot = out[:, :13].sum(1).mean()
hops = out[:, 13].mean(); merges = out[:, 14].mean()
print(f"mean clocks per query {tot:.0f}; hops {hops:.1f}; merging hops {merges:.1f}")
for i, nm in enumerate(names[:13]):
    m = out[:, i].mean()
    per = m / hops if nm not in ("setup", "rerank", "topk") else m
    print(f"  {nm:9s} {m:10.0f} clk/query  {m / tot * 100:5.1f}%   {per:8.0f} clk/{'hop' if nm not in ('setup','rerank','topk') else 'query'}")
