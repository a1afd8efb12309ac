#!/usr/bin/env python
"""bench.py — QPS of the batched greedy Vamana search at recall@10 >= 0.90 / 0.95 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference ...                     # host-core baseline (oracle port, all cores)

A "step" is one pass of the hot path over one batch of Q = 10 000 queries.  Default workload = BASELINE.json
configs[1]: SIFT1M-shape synthetic (N = 10^6, D = 128, uint8, R = 64, PQ 32 B/vector), BANG_Inmemory
semantics, k = 10.  The index (data, Vamana graph, PQ, ground truth) is generated on the box by the
committed builder and written in the reference's file formats; the search loads those files through
bang_load.  The worklist length L is the smallest of a sweep reaching the recall target (found before the
timed region, as the reference's driver sweeps L, test_driver.cpp:388-420).

JSON line (one, from rank 0): see the contract in the task statement; `value` = whole-job QPS at
recall@10 >= 0.90 with queries already resident in HBM (device-side CUDA events, max over ranks);
`e2e` = the same through the host-facing bang_query call (host query buffer in, host ids/dists out);
`at_recall_95` repeats both at the >= 0.95 operating point.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: dict(n, d, dtype, m, q, mode)
    "sift10k": dict(n=10_000, d=128, dtype="uint8", m=32, q=100, mode="inmemory", label="C1 SIFT10K-shape"),
    "sift1m": dict(n=1_000_000, d=128, dtype="uint8", m=32, q=10_000, mode="inmemory", label="C2 SIFT1M-shape"),
    "gist1m": dict(n=1_000_000, d=960, dtype="float", m=None, q=10_000, mode="exact", label="C3 GIST1M-shape"),
    "deep100m": dict(n=100_000_000, d=96, dtype="float", m=32, q=10_000, mode="inmemory", label="C4 DEEP100M-shape"),
}
K = 10
L_SWEEP = (10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 56, 64, 80, 96, 112, 128, 152, 176, 200, 256, 320, 400, 512)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(workload: str, mode: str, L: int, q: int):
    """DRAM bytes (read + write) of one launch of the search kernel from the committed ncu capture, if that
    capture was taken on this very workload / mode / L / batch; else None (B200 bench contract: traffic or null)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1e_traffic.json")) as f:
            t = json.load(f)
        if (t["workload"], t["mode"], t["L"], t["queries"]) == (workload, mode, L, q):
            return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------------------------------
# index generation
# ----------------------------------------------------------------------------------------------------
def make_index(wl: dict, workdir: str, device, builder: str, rank: int, world: int):
    """Rank 0 generates the dataset + index files; the other ranks wait for them (same box)."""
    import bang_b200  # noqa: F401
    from bang_b200 import builder as B
    nq = wl["q"] * max(1, world)  # weak scaling: every rank searches its own batch of wl["q"] queries
    prefix = os.path.join(workdir, f"{wl['dtype']}_{wl['n']}_{wl['d']}_q{nq}")
    done = prefix + ".done"
    if rank == 0 and not os.path.exists(done):
        t0 = time.time()
        info = B.make_fixture_auto(prefix, wl["n"], wl["d"], wl["dtype"], nq, wl["m"], k_gt=100, device=device,
                                   builder=builder)
        log(f"[bench] index built in {time.time() - t0:.1f}s ({info})")
        with open(done, "w") as f:
            f.write("ok")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    return prefix


def pick_L(search, queries, gt_ids, gt_d, targets=(90.0, 95.0)):
    """Smallest L of the sweep reaching each recall target (outside the timed region)."""
    from bang_b200 import recall
    found = {}
    curve = []
    Q = len(queries)
    for L in L_SWEEP:
        if L < K:
            continue
        search.bang_set_searchparams(K, L)
        search.bang_alloc(Q)
        search.bang_init(Q)
        ids, _ = search.bang_query(queries)
        search.bang_free()
        r = recall.calculate_recall(gt_ids, gt_d, ids, K)
        curve.append((L, round(r, 2)))
        for t in targets:
            if t not in found and r >= t:
                found[t] = (L, r)
        if len(found) == len(targets):
            break
    return found, curve


# ----------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------
def time_config(search, queries_np, L, steps, warmup, device, world):
    """Returns dict with device-resident QPS, e2e QPS, stats for one worklist length."""
    import torch
    from bang_b200 import api
    Q = len(queries_np)
    search.bang_set_searchparams(K, L)
    search.bang_alloc(Q)
    search.bang_init(Q)
    stream = torch.cuda.current_stream(device)
    d_q = torch.from_numpy(queries_np).to(device)
    d_ids = torch.empty((Q, K), dtype=torch.int64, device=device)
    d_d = torch.empty((Q, K), dtype=torch.float32, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize(device)

    # ---- device-resident: inputs already in HBM; CUDA events on the launching stream ----
    for _ in range(warmup):
        search.query_device(d_q.data_ptr(), Q, d_ids.data_ptr(), d_d.data_ptr(), stream.cuda_stream)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for s in range(steps):
        flush.zero_()  # L2 flush between timed iterations (the index itself, 384 MB + 32 MB, also exceeds L2)
        ev[s][0].record(stream)
        search.query_device(d_q.data_ptr(), Q, d_ids.data_ptr(), d_d.data_ptr(), stream.cuda_stream)
        ev[s][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    stats = search.last_stats(Q)
    ids_dev = d_ids.cpu().numpy().astype(np.uint64)

    # ---- end to end: host query buffer -> bang_query -> host ids/dists (H2D + D2H inside) ----
    for _ in range(max(1, warmup // 2)):
        search.bang_init(Q)
        search.bang_query(queries_np)
    e2e_ms = []
    barrier()
    for s in range(steps):
        flush.zero_()
        torch.cuda.synchronize(device)
        search.bang_init(Q)
        t0 = time.perf_counter()
        ids_host, _ = search.bang_query(queries_np)
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
    tm = search.last_timing()
    assert np.array_equal(ids_host, ids_dev), "device-resident and host paths disagree"
    search.bang_free()
    return dict(kern_ms=kern_ms, e2e_ms=e2e_ms, stats=stats, ids=ids_host, timing=tm, wall_s=t_wall)


def cpu_baseline_sample(prefix, wl, queries, L, target_s=12.0):
    """Oracle port timed on the host cores on a bounded sample of the same workload (rank 0, N=1 only)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    ox = O.OracleIndex.from_files(prefix, with_pq=wl["mode"] != "exact")
    mode = {"base": O.MODE_BASE, "inmemory": O.MODE_INMEMORY, "exact": O.MODE_EXACT}[wl["mode"]]
    cores = os.cpu_count() or 1
    n = min(len(queries), max(64, 4 * cores))
    t0 = time.perf_counter()
    ox.search(queries[:n], K, L, mode=mode, nthreads=cores)
    dt = time.perf_counter() - t0
    # scale the sample to ~target_s of CPU work
    n2 = int(min(len(queries), max(n, n * target_s / max(dt, 1e-3))))
    t0 = time.perf_counter()
    ox.search(queries[:n2], K, L, mode=mode, nthreads=cores)
    dt = time.perf_counter() - t0
    return dict(value=n2 / dt, unit="QPS", cores=cores, kind="port",
                sample=f"{n2} of {len(queries)} queries at L={L}, oracle/bang_oracle.c with OpenMP over queries")


def reference_cuda_run(prefix, wl, paths, Q, L, gt_ids, gt_d):
    """The UNMODIFIED reference (BANG_Base built from /root/reference into oracle/_ref, sm_100a) on this GPU over the
    same index files: its own BANGSearch<T> API through oracle/ref_driver.cpp, 4 runs, first discarded (the
    authors' protocol, BANG_Inmemory/parANN.h:30-31).  Comparison only; PQ modes only."""
    from bang_b200 import recall
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(exe) or wl["mode"] == "exact":
        return None
    out = os.path.join(os.path.dirname(prefix), "ref_ids.bin")
    dt = {"uint8": "uint8", "int8": "int8", "float": "float"}[wl["dtype"]]
    try:
        r = subprocess.run([exe, prefix, paths.query, str(Q), str(K), str(L), dt, out, "4"], capture_output=True, text=True,
                           timeout=900)
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}
    if r.returncode != 0:
        return {"error": (r.stdout[-300:] + r.stderr[-300:])}
    ms = [float(l.split()[2]) for l in r.stdout.splitlines() if l.startswith("RUN ")]
    ids = np.fromfile(out, dtype=np.uint64).reshape(Q, K)
    rec = recall.calculate_recall(gt_ids[:Q], gt_d[:Q], ids, K)
    use = ms[1:] if len(ms) > 1 else ms
    gm = float(np.exp(np.mean(np.log(use))))
    return {"qps": Q / (gm * 1e-3), "ms": gm, "runs_ms": ms, "L": L, "recall_at_10": round(rec, 2),
            "what": "unmodified BANG_Base (oracle/_ref/libbang.so, nvcc -arch sm_100a) via its BANGSearch<T> API, wall clock around bang_query"}


def run_b200(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the sm_100a kernels are the only search path (no CPU fallback)")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    import bang_b200  # noqa: F401
    from bang_b200 import api, formats, recall

    wl = dict(WORKLOADS[args.workload])
    if args.n:
        wl["n"] = args.n
    if args.q:
        wl["q"] = args.q
    workdir = args.workdir or os.path.join(tempfile.gettempdir(), "bang_b200_bench")
    os.makedirs(workdir, exist_ok=True)
    prefix = make_index(wl, workdir, device, args.builder, rank, world)
    paths = formats.IndexPaths(prefix)
    queries = formats.read_bin(paths.query, api.NP[wl["dtype"]])[: wl["q"] * world]
    gt_ids, gt_d = formats.read_truthset(paths.truth)

    # replicated index; every rank searches its own batch of wl["q"] queries (weak scaling), no collective on
    # the data path
    per = wl["q"]
    lo, hi = rank * per, (rank + 1) * per
    my_q = np.ascontiguousarray(queries[lo:hi])

    search = api.BANGSearch(wl["dtype"], wl["mode"], device=local)
    t0 = time.time()
    if not search.bang_load(prefix):
        raise SystemExit("bang_load failed: " + search.last_error)
    log(f"[bench r{rank}] bang_load {time.time() - t0:.1f}s, index {search.info().device_bytes / 2**20:.0f} MiB in HBM")
    search.set_dists_layout(api.DISTS_QUERY_MAJOR)

    if args.L:
        found = {90.0: (args.L, float("nan")), 95.0: (args.L95 or args.L, float("nan"))}
        curve = []
    else:
        found, curve = pick_L(search, queries[:per], gt_ids[:per], gt_d[:per])  # every rank: same deterministic sweep on batch 0
        if 90.0 not in found or 95.0 not in found:
            raise SystemExit(f"recall targets not reached in the L sweep: {curve}")
    if rank == 0:
        log(f"[bench] recall curve {curve}; operating points {found}")

    sampler = ClockSampler(local)
    sampler.start()
    res = {}
    for tgt in (90.0, 95.0):
        L = found[tgt][0]
        res[tgt] = time_config(search, my_q, L, args.steps, args.warmup, device, world)
    clocks = sampler.stop()

    # gather results: ids for the recall check, timings for max-over-ranks
    from bang_b200 import sharding

    def allmax(xs):
        return sharding.max_over_ranks(xs, device=device)

    gather_np = sharding.gather_rows

    out = {}
    for tgt in (90.0, 95.0):
        r = res[tgt]
        kern = allmax(r["kern_ms"])
        e2e = allmax(r["e2e_ms"])
        ids_all = gather_np(r["ids"])
        stats_all = {k_: gather_np(v) for k_, v in r["stats"].items()}
        rec = recall.calculate_recall(gt_ids[: len(ids_all)], gt_d[: len(ids_all)], ids_all, K)
        esz = 4 if wl["dtype"] == "float" else 1
        bq = api.algorithmic_bytes(stats_all, wl["mode"], wl["d"], esz, wl["m"] or 0, K)
        ms = float(np.mean(kern))
        out[tgt] = dict(L=found[tgt][0], recall=rec, ms=ms, qps=len(ids_all) / (ms * 1e-3),
                        e2e_ms=float(np.mean(e2e)), e2e_qps=len(ids_all) / (np.mean(e2e) * 1e-3),
                        bytes_per_query=float(bq.mean()), bytes_total=float(bq.sum()),
                        hops=float(stats_all["hops"].mean()), n_cand=float(stats_all["n_cand"].mean()),
                        timing=r["timing"], my_bytes=float(api.algorithmic_bytes(r["stats"], wl["mode"], wl["d"], esz, wl["m"] or 0, K).sum()),
                        my_ms=float(np.mean(r["kern_ms"])))

    if rank != 0:
        return
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    p90, p95 = out[90.0], out[95.0]
    ach = p90["my_bytes"] / (p90["my_ms"] * 1e-3) / 1e9  # rank-0 kernel: algorithmic bytes per launch / its duration
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(prefix, wl, queries, p90["L"])
    tm = p90["timing"]
    Qtot = len(queries)
    esz = 4 if wl["dtype"] == "float" else 1
    info = search.info()
    line = {
        "metric": "QPS at recall@10 >= 0.90 (batched greedy Vamana search)",
        "value": p90["qps"], "unit": "QPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": p90["ms"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8 codes / fp32 ADC sums" if wl["mode"] != "exact" else "fp32",
        "data": "synthetic clustered (Gaussian mixture), index built on the box by the committed builder",
        "config": {"workload": f"{wl['label']}: N={wl['n']} D={wl['d']} {wl['dtype']} R=64 "
                               + (f"PQ m={wl['m']}" if wl["m"] else "no PQ") + f", Q={Qtot}, k={K}, mode={wl['mode']}",
                   "L_at_recall_90": p90["L"], "recall_at_10": round(p90["recall"], 2),
                   "parallelism": f"index replicated on {world} GPU(s), one batch of {wl['q']} queries per GPU, no collective",
                   "l2": "256 MiB buffer written between timed iterations; index (rows+codes) larger than L2",
                   "builder": args.builder},
        "e2e": {"value": p90["e2e_qps"], "unit": "QPS", "ms_per_step": p90["e2e_ms"],
                "h2d_bytes_per_step": int(Qtot * wl["d"] * esz), "d2h_bytes_per_step": int(Qtot * K * 12)},
        "at_recall_95": {"L": p95["L"], "recall_at_10": round(p95["recall"], 2), "value": p95["qps"], "ms_per_step": p95["ms"],
                         "e2e": p95["e2e_qps"], "bytes_per_query": p95["bytes_per_query"]},
        "gpu_launches": args.steps * 1,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                     "traffic": ncu_traffic(args.workload, wl["mode"], p90["L"], wl["q"]),
                     "kernel": "bang_search_kernel (fused traversal, 1 launch per step)",
                     "bytes_per_query": p90["bytes_per_query"], "hops_per_query": p90["hops"],
                     "candidates_per_query": p90["n_cand"], "peak_source": peak_src,
                     "grid": tm.grid, "block": tm.block, "smem_bytes": tm.smem_bytes, "ctas_per_sm": tm.ctas_per_sm},
        "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if world == 1 and not args.no_ref_cuda:
        search.bang_unload()
        ref = reference_cuda_run(prefix, wl, paths, wl["q"], p90["L"], gt_ids, gt_d)
        if ref:
            line["reference_cuda_b200"] = ref
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference's search logic on the host cores (oracle port), all threads
# ----------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bang_b200  # noqa: F401
    from bang_b200 import api, formats, recall
    import oracle as O
    wl = dict(WORKLOADS[args.workload])
    if args.n:
        wl["n"] = args.n
    if args.q:
        wl["q"] = args.q
    workdir = args.workdir or os.path.join(tempfile.gettempdir(), "bang_b200_bench")
    os.makedirs(workdir, exist_ok=True)
    device = "cpu"
    try:
        import torch
        if torch.cuda.is_available():
            device = torch.device("cuda", 0)
    except Exception:
        pass
    prefix = make_index(wl, workdir, device, args.builder, 0, 1)
    paths = formats.IndexPaths(prefix)
    queries = formats.read_bin(paths.query, api.NP[wl["dtype"]])[: wl["q"]]
    gt_ids, gt_d = formats.read_truthset(paths.truth)
    ox = O.OracleIndex.from_files(prefix, with_pq=wl["mode"] != "exact")
    mode = {"base": O.MODE_BASE, "inmemory": O.MODE_INMEMORY, "exact": O.MODE_EXACT}[wl["mode"]]
    cores = os.cpu_count() or 1
    # operating point: same sweep, on a bounded sample
    ns = min(len(queries), 512)
    L90 = args.L
    if not L90:
        for L in L_SWEEP:
            ids, _ = ox.search(queries[:ns], K, L, mode=mode, nthreads=cores)
            if recall.calculate_recall(gt_ids[:ns], gt_d[:ns], ids, K) >= 90.0:
                L90 = L
                break
    if not L90:
        L90 = L_SWEEP[-1]
    # bounded sample per step: ~ (60 s total) / (steps + warmup)
    t0 = time.perf_counter()
    ox.search(queries[:ns], K, L90, mode=mode, nthreads=cores)
    per_q = (time.perf_counter() - t0) / ns
    budget = 60.0 / max(1, args.steps + args.warmup)
    n = int(min(len(queries), max(ns, budget / per_q)))
    for _ in range(args.warmup):
        ox.search(queries[:n], K, L90, mode=mode, nthreads=cores)
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        ids, _ = ox.search(queries[:n], K, L90, mode=mode, nthreads=cores)
        ts.append(time.perf_counter() - t0)
    rec = recall.calculate_recall(gt_ids[:n], gt_d[:n], ids, K)
    qps = n / float(np.mean(ts))
    Qtot = len(queries)
    sample = f"{n} of {Qtot} queries per step at L={L90}; oracle/bang_oracle.c (CPU restatement of the reference's search), OpenMP over queries"
    line = {"impl": "reference", "metric": "QPS at recall@10 >= 0.90 (batched greedy Vamana search)", "value": qps, "unit": "QPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(ts)) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic clustered",
            "config": {"workload": f"{wl['label']}: N={wl['n']} D={wl['d']} {wl['dtype']} R=64 "
                                   + (f"PQ m={wl['m']}" if wl["m"] else "no PQ") + f", Q={Qtot}, k={K}, mode={wl['mode']}",
                       "L_at_recall_90": L90, "recall_at_10": round(rec, 2),
                       "note": "the reference is GPU-only; its search logic runs here on the host cores (kind=port)"},
            "cpu_baseline": {"value": qps, "unit": "QPS", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "QPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sift1m", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="override the number of base points (parity/debug runs only)")
    ap.add_argument("--q", type=int, default=0)
    ap.add_argument("--L", type=int, default=0, help="fix the worklist length instead of sweeping")
    ap.add_argument("--L95", type=int, default=0)
    ap.add_argument("--builder", default="auto", choices=["auto", "gpu", "cpu"])
    ap.add_argument("--workdir", default="")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
