"""C5 SIFT1B-shape: graph rows sharded id % G over the GPUs' HBM, PQ codes replicated, P2P neighbour fetch in the
traversal kernel (replaces BANG_Base's host-RAM graph + PCIe fetch).  Index built on the GPUs (build_sharded.py).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29551 profiles/c5_run.py N [Q] [Ls]
Prints one JSON line per worklist length from rank 0 and writes them to gpurun_out/c5_<N>.jsonl"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bang_b200
from bang_b200 import api, build_sharded, recall, sharding

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
Ls = sys.argv[3].split(",") if len(sys.argv) > 3 else ["64", "128", "176", "256"]   # "L" or "L:warps_per_sm"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t_all = time.time()
s = api.BANGSearch("uint8", "inmemory", device=local)
s.set_sharding(rank, world)
n_gt = min(1000, Q)
my_q, gt_ids, gt_d, medoid, T = build_sharded.build_and_load(s, N, 128, Q, n_gt, P_per_rank=int(os.environ.get("C5_SHARDS_PER_RANK", "4")),
                                                             passes=int(os.environ.get("C5_PASSES", "2")),
                                                             ownership=os.environ.get("C5_OWNERSHIP", "mod"))
my_idx = T.pop("my_idx")
T.pop("home", None)
Qall = Q * world          # the global batch; under partition ownership the ranks' shares differ in size
Q = len(my_q)
sharding.exchange_shards(s, rank, world)
info = s.info()
if rank == 0:
    print(f"[c5] N={N} built+loaded in {time.time() - t_all:.1f}s {T}; per-GPU HBM {info.device_bytes / 2**30:.1f} GiB; medoid {medoid}", flush=True)
s.set_dists_layout(api.DISTS_QUERY_MAJOR)
out = []
try:   # SM clock + throttle reasons sampled right after each timed batch (same process, NVML)
    import pynvml
    pynvml.nvmlInit()
    _nv = pynvml.nvmlDeviceGetHandleByIndex(local)
    def sample_clock():
        return (pynvml.nvmlDeviceGetClockInfo(_nv, pynvml.NVML_CLOCK_SM), int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(_nv)))
except Exception:
    def sample_clock():
        return (0, 0)
for cfg in Ls:
    solo = cfg.endswith("s")          # "176s": ranks > 0 stay idle (their shards are still served)
    L, _, wps = cfg.rstrip("s").partition(":")
    L = int(L)
    if wps:
        os.environ["BANG_B200_WARPS_PER_SM"] = wps
    else:
        os.environ.pop("BANG_B200_WARPS_PER_SM", None)
    s.bang_set_searchparams(10, L)
    s.bang_alloc(Q)
    ms, e2e, clk = [], [], []
    for r in range(7):
        s.bang_init(Q)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if solo and rank != 0 and r > 0:
            ms.append(0.0); e2e.append(0.0)
            dist.barrier()
            continue
        ids, d = s.bang_query(my_q)
        if solo:
            dist.barrier()
        e2e.append((time.perf_counter() - t0) * 1e3)
        clk.append(sample_clock())
        ms.append(s.last_timing().kernel_ms)
    st = s.last_stats(Q)
    if os.environ.get("C5_PHASES") and rank == 0:
        import ctypes
        fn = s._lib.bang_b200_debug_phase_clocks
        fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        ph = np.zeros((Q, 16), dtype=np.int64)
        if fn(s._h, ph.ctypes.data) == 0:
            names = ["setup", "adjwait", "hash", "bloom", "compact", "codewait", "lut", "scan", "decide", "merge", "unvis", "rerank"]
            hops = ph[:, 13].mean()
            cyc = ph[:, :12].sum(1); ns = ph[:, 15]; t0 = ph[:, 12] - ph[:, 12].min()
            print(f"[c5] per-query cycles mean {cyc.mean():.0f} p99 {np.percentile(cyc, 99):.0f} max {cyc.max()}; wall us mean {ns.mean() / 1e3:.0f} p99 "
                  f"{np.percentile(ns, 99) / 1e3:.0f} max {ns.max() / 1e3:.0f}; SM clock GHz {cyc.sum() / ns.sum():.3f}; last start {t0.max() / 1e6:.2f} ms, "
                  f"last end {(t0 + ns).max() / 1e6:.2f} ms", flush=True)
            for qi in np.argsort(-ns)[:4]:
                print(f"[c5] slow query {qi}: wall {ns[qi] / 1e3:.0f} us start {t0[qi] / 1e6:.2f} ms hops {ph[qi, 13]} merges {ph[qi, 14]} n_cand {st['n_cand'][qi]} sum_deg {st['sum_deg'][qi]} kcycles",
                      {n: int(ph[qi, i] // 1000) for i, n in enumerate(names)}, flush=True)
            print("[c5] phase clocks per hop:", {n: int(ph[:, i].mean() / (hops if n not in ("setup", "rerank") else 1)) for i, n in enumerate(names)}, flush=True)
    per_rank = [None] * world
    dist.all_gather_object(per_rank, round(float(np.mean(ms[2:])), 2))
    clk_rank = [None] * world
    dist.all_gather_object(clk_rank, [int(np.median([c[0] for c in clk] or [0])), int(np.bitwise_or.reduce([c[1] for c in clk] or [0]))])
    k_ms = sharding.max_over_ranks([float(np.mean(ms[2:]))], device=torch.device("cuda", local))[0]
    e_ms = sharding.max_over_ranks([float(np.mean(e2e[2:]))], device=torch.device("cuda", local))[0]
    esz = 1
    bq = api.algorithmic_bytes(st, "inmemory", 128, esz, 32, 10)
    adj_vec = 4 * st["hops"].astype(np.int64) + 4 * st["sum_deg"].astype(np.int64) + st["hops"].astype(np.int64) * 128
    nvlink = adj_vec * (world - 1) / world
    parts = [None] * world   # results of the ground-truth queries, wherever they were searched
    keep = my_idx < n_gt
    dist.all_gather_object(parts, (my_idx[keep], ids[keep]))
    q_per_rank = [None] * world
    dist.all_gather_object(q_per_rank, Q)
    if rank == 0:
        got = np.zeros((n_gt, 10), dtype=np.uint64)
        for idx_r, ids_r in parts:
            got[idx_r] = ids_r
        rec = recall.calculate_recall(gt_ids[:n_gt], gt_d[:n_gt], got, 10)
        line = {"metric": "QPS at recall@10 (batched greedy Vamana search, SIFT1B-shape, graph sharded over HBM)", "n_gpus": world,
                "N": N, "L": L, "warps_per_sm": int(wps) if wps else 16, "solo": solo, "kernel_ms_per_rank": per_rank, "sm_mhz_and_event_reasons_per_rank": clk_rank, "kernel_ms_runs_rank0": [round(x, 2) for x in ms], "recall_at_10": round(rec, 2), "recall_queries": n_gt, "kernel_ms_max_over_ranks": k_ms,
                "value": Qall / (k_ms * 1e-3), "e2e": Qall / (e_ms * 1e-3), "unit": "QPS",
                "ownership": os.environ.get("C5_OWNERSHIP", "mod"), "queries_per_gpu": q_per_rank, "hops_per_query": float(st["hops"].mean()), "candidates_per_query": float(st["n_cand"].mean()),
                "bytes_per_query": float(bq.mean()), "nvlink_bytes_per_query": float(nvlink.mean()),
                "per_gpu_hbm_gib": info.device_bytes / 2**30, "build_seconds": T}
        print(json.dumps(line), flush=True)
        out.append(line)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"c5_{N}.jsonl"), "w") as f:   # rewritten after every L: a cut-off run keeps what it measured
            for l in out:
                f.write(json.dumps(l) + "\n")
    s.bang_free()
dist.barrier()
s.bang_unload()
dist.barrier()
dist.destroy_process_group()
