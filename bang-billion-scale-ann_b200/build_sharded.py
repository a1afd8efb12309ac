"""Distributed index construction + device-resident load for the SIFT1B-shape configuration (C5).

Run under torchrun with G ranks (one per GPU of one box).  Nothing here touches the disk: a 10^9-point `_disk.bin`
is 388 GB and the box has 80 GB, so the index is built on the GPUs and handed to the search library through
`bang_b200_load_device_*` (same HBM layout and the same kernel as the file path).

Method = DiskANN's own recipe for billion-point builds (build_merged_vamana_index): partition the points into P
overlapping shards (every point goes to its 2 nearest of P partition centres), build one Vamana graph per shard
(csrc/builder.cu, one shard at a time per GPU), take the union of the out-neighbours of the two copies of every
node and truncate it to R = 64 by a pseudo-random rule (DiskANN truncates randomly; here: the 64 smallest
hash(node, neighbour), which makes the merge order-independent).

The synthetic dataset (Gaussian mixture, seeded per chunk) is generated redundantly on every GPU, chunk by chunk,
so no vector ever crosses NVLink: a rank keeps the members of its 4 shards and the rows it owns (id % G == rank).
The per-node adjacency lists travel once, with all_to_all, from the rank that built the shard to the rank that owns
the row.  PQ codes are encoded by the chunk's owner and broadcast (they are replicated on every GPU).
"""
from __future__ import annotations

import sys
import time

import numpy as np
import torch
import torch.distributed as dist

from . import synth
from .builder import build_vamana_gpu

SIGMA_U8 = 24.0


def log(rank, *a):
    if rank == 0:
        print("[c5]", *a, file=sys.stderr, flush=True)   # (stdout belongs to the caller: bench.py prints one JSON line there)


def mixture_centers(n_clusters: int, d: int, device) -> torch.Tensor:
    g = torch.Generator(device=device)
    g.manual_seed(synth.BASE_SEED ^ 0xC0)
    return torch.rand(n_clusters, d, generator=g, device=device) * 192.0 + 32.0


def gen_chunk(centers: torch.Tensor, chunk: int, n: int, seed: int) -> torch.Tensor:
    """uint8 [n][D]; identical on every rank (same seed, same device type)."""
    g = torch.Generator(device=centers.device)
    g.manual_seed((seed << 20) ^ chunk)
    assign = torch.randint(0, centers.shape[0], (n,), generator=g, device=centers.device)
    pts = centers[assign] + SIGMA_U8 * torch.randn(n, centers.shape[1], generator=g, device=centers.device)
    return pts.round_().clamp_(0, 255).to(torch.uint8)


def _hash32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """cheap 2-input integer mix (int64 tensors in, values in [0, 2^31))"""
    x = (a * 0x9E3779B1 + b * 0x85EBCA77) & 0xFFFFFFFF
    x = (x ^ (x >> 15)) * 0x2C1B3C6D & 0xFFFFFFFF
    x = (x ^ (x >> 12)) * 0x297A2D39 & 0xFFFFFFFF
    return (x ^ (x >> 15)) & 0x7FFFFFFF


def balance_partitions(sizes, G: int) -> torch.Tensor:
    """Partitions -> GPUs, largest first onto the lightest GPU.  Returns int64 [P] (on the CPU)."""
    sizes = [float(x) for x in sizes]
    load = [0.0] * G
    out = [0] * len(sizes)
    for p in sorted(range(len(sizes)), key=lambda i: (-sizes[i], i)):
        g = min(range(G), key=lambda i: (load[i], i))
        out[p] = g
        load[g] += sizes[p]
    return torch.tensor(out, dtype=torch.int64)


def assign_ids(owner: torch.Tensor, counts: list, G: int) -> torch.Tensor:
    """Node ids under partition ownership: the i-th point (in generation order) owned by GPU g gets id i * G + g, so
    the search library's ownership rule (id mod G, local row id div G) holds without the rows being interleaved by
    generation order.  owner: int64 [n] for one chunk; counts: running per-GPU totals, updated in place."""
    ids = torch.empty_like(owner)
    for g in range(G):
        msk = owner == g
        k = int(msk.sum())
        if k:
            ids[msk] = (torch.arange(k, device=owner.device, dtype=torch.int64) + counts[g]) * G + g
        counts[g] += k
    return ids


def merge_lists(adj_own: torch.Tensor, rows: torch.Tensor, new: torch.Tensor, gid_of_row_mul: int, gid_of_row_add: int) -> None:
    """adj_own[rows] <- the 64 smallest-hash members of (adj_own[rows] ∪ new); rows unique within the call.
    Rows that are still empty (the first of a node's two copies) just take the list."""
    rows = rows.long()
    empty = adj_own[rows, 0] < 0
    if bool(empty.any()):
        adj_own[rows[empty]] = new[empty]
    if bool(empty.all()):
        return
    rows, new = rows[~empty], new[~empty]
    step = 1 << 20
    for s in range(0, rows.numel(), step):
        r = rows[s:s + step]
        cat = torch.cat([adj_own[r], new[s:s + step]], 1)                    # int32 [m,128], -1 = empty
        cat, _ = torch.sort(cat, dim=1)
        dup = torch.zeros_like(cat, dtype=torch.bool)
        dup[:, 1:] = cat[:, 1:] == cat[:, :-1]
        v = (r * gid_of_row_mul + gid_of_row_add)[:, None]
        catl = cat.long()
        h = _hash32(v.expand_as(catl), catl)
        h = torch.where((cat < 0) | dup | (catl == v), torch.full_like(h, 1 << 40), h)
        hs, idx = torch.topk(h, 64, dim=1, largest=False)
        out = torch.gather(cat, 1, idx)
        out = torch.where(hs >= (1 << 40), torch.full_like(out, -1), out)
        adj_own[r] = out
        del cat, catl, dup, h, hs, idx, out


def build_and_load(search, N: int, D: int, n_queries_per_rank: int, n_gt_queries: int, m: int = 32, P_per_rank: int = 4,
                   chunk: int = 1 << 22, L_build: int = 64, passes: int = 2, seed: int = synth.BASE_SEED,
                   ownership: str = "mod", device=None, build_fn=None, shard_slack: float = 1.30, gt_fn=None):
    """Builds the sharded index on the GPUs and loads this rank's shard into `search` (an api.BANGSearch with
    set_sharding(rank, world) already called).  Returns (queries_for_this_rank u8 [q][D], gt_ids [n_gt][100] or None
    on ranks != 0, gt_dists, medoid, timings dict).

    ownership = "mod": node ids are generation order, rows owned by id mod G (every hop is remote with probability
    (G-1)/G).  ownership = "partition" (experimental, profiles/locality_sim.py): a row is owned by the GPU of its
    nearest partition centre and ids are assigned so that id mod G is that GPU (assign_ids); the caller should then
    send every query to `query_home(...)`.  The returned dict T carries "home" (int64 [G*q], the home GPU of every
    query of the global batch) and "n_virtual" (the id space, G x the largest per-GPU row count) in that mode.

    device / build_fn exist for the CPU dry run of this logic (tests/test_build_sharded.py: gloo, host Vamana builder):
    build_fn(vectors [n][D], entry, L, passes, seed) -> int32 [n][64] local neighbour ids, -1 = unused."""
    rank, G = dist.get_rank(), dist.get_world_size()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    on_gpu = dev.type == "cuda"
    mem_gib = (lambda: torch.cuda.memory_allocated() >> 30) if on_gpu else (lambda: 0)
    if build_fn is None:
        build_fn = lambda vec, entry, L, passes, seed: build_vamana_gpu(vec, entry, L=L, passes=passes, seed=seed, device_out=True)
    P = P_per_rank * G
    T = {}
    t0 = time.time()
    n_clusters = max(16, N // 1000)
    centers = mixture_centers(n_clusters, D, dev)
    n_chunks = (N + chunk - 1) // chunk
    # queries (identical on every rank), ground-truth subset = the first n_gt_queries of rank 0's batch
    qg = torch.Generator(device=dev)
    qg.manual_seed(synth.QUERY_SEED ^ seed)
    qa = torch.randint(0, n_clusters, (G * n_queries_per_rank,), generator=qg, device=dev)
    queries = (centers[qa] + SIGMA_U8 * torch.randn(qa.numel(), D, generator=qg, device=dev)).round_().clamp_(0, 255).to(torch.uint8)
    gt_q_raw = queries[:n_gt_queries].contiguous()
    if gt_fn is None and dev.type != "cuda":   # CPU dry run: plain torch (uint8 distances are exact in fp32)
        def gt_fn(x, q, k):
            d2 = torch.cdist(q.float(), x.float()) ** 2
            dd, ii = torch.topk(d2.round(), k, dim=1, largest=False)
            return dd, ii
    # partition centres: Lloyd on the mixture centres, rank 0 decides
    pc = centers[torch.randperm(n_clusters, device=dev)[:P]].clone()
    if rank == 0:
        for _ in range(8):
            lab = torch.cdist(centers, pc).argmin(1)
            sums = torch.zeros_like(pc).index_add_(0, lab, centers)
            cnt = torch.zeros(P, device=dev).index_add_(0, lab, torch.ones(n_clusters, device=dev))
            pc = torch.where(cnt[:, None] > 0, sums / cnt.clamp(min=1)[:, None], pc)
    dist.broadcast(pc, 0)
    pcn = (pc * pc).sum(1)
    by_part = ownership == "partition"
    gpu_of_part = None
    if by_part:
        lab_c = (pcn[None, :] - 2.0 * (centers @ pc.T)).argmin(1)           # expected partition sizes ~ mixture centres per partition
        gpu_of_part = balance_partitions(torch.bincount(lab_c, minlength=P).tolist(), G).to(dev)
        T["home"] = gpu_of_part[(pcn[None, :] - 2.0 * (queries.float() @ pc.T)).argmin(1)].cpu()
    # PQ: rank 0 trains on chunk 0, everyone receives pivots / centroid
    offs = synth.chunk_offsets_even(D, m)
    piv = torch.zeros(256, D, device=dev)
    cen = torch.zeros(D, device=dev)
    if rank == 0:
        x0 = gen_chunk(centers, 0, min(chunk, N), seed)
        p_np, c_np, _ = synth.train_pq(x0, m)
        piv.copy_(torch.from_numpy(p_np)); cen.copy_(torch.from_numpy(c_np))
        del x0
    dist.broadcast(piv, 0); dist.broadcast(cen, 0)
    mean = centers.mean(0)
    T["setup"] = time.time() - t0

    # ---- pass 1: every rank generates every chunk; keeps its shards' members, its own rows, its share of the GT ----
    t0 = time.time()
    my_shards = [rank + G * j for j in range(P_per_rank)]
    cap = int(2 * N / P * shard_slack) + chunk // 8   # shard buffers: every point lands in 2 of the P shards
    sh_vec = [torch.empty((cap, D), dtype=torch.uint8, device=dev) for _ in my_shards]
    sh_gid = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in my_shards]
    sh_n = [0] * P_per_rank
    n_own = (N - rank + G - 1) // G if not by_part else int(N / G * max(1.25, shard_slack)) + chunk // 4   # capacity; the count is known after pass 1
    own_vec = torch.empty((n_own, D), dtype=torch.uint8, device=dev)
    own_cnt = [0] * G
    gid_all = torch.empty(N, dtype=torch.int32, device=dev) if by_part else None       # new id of every generated point
    best_d = torch.full((n_gt_queries, 100), float("inf"), device=dev)
    best_i = torch.zeros((n_gt_queries, 100), dtype=torch.int64, device=dev)
    med_d, med_i = float("inf"), 0
    for c in range(n_chunks):
        n = min(chunk, N - c * chunk)
        x = gen_chunk(centers, c, n, seed)
        xf = x.float()
        d2p = pcn[None, :] - 2.0 * (xf @ pc.T)
        top2 = torch.topk(d2p, 2, dim=1, largest=False)[1]
        if by_part:   # id = (running index among the points of the owning GPU) * G + owning GPU
            own = gpu_of_part[top2[:, 0]]
            gid = assign_ids(own, own_cnt, G)
            gid_all[c * chunk:c * chunk + n] = gid.to(torch.int32)
        else:
            gid = torch.arange(c * chunk, c * chunk + n, device=dev, dtype=torch.int64)
        for j, s in enumerate(my_shards):
            sel = ((top2[:, 0] == s) | (top2[:, 1] == s)).nonzero().squeeze(1)
            k = sel.numel()
            if sh_n[j] + k > cap:
                raise RuntimeError(f"shard {s} exceeds its buffer ({sh_n[j] + k} > {cap})")
            sh_vec[j][sh_n[j]:sh_n[j] + k] = x[sel]
            sh_gid[j][sh_n[j]:sh_n[j] + k] = gid[sel].to(torch.int32)
            sh_n[j] += k
        if by_part:
            msk = own == rank
            rows_here = gid[msk] // G
            if rows_here.numel() and int(rows_here.max()) >= n_own:
                raise RuntimeError(f"rank {rank} owns more rows than its buffer holds ({int(rows_here.max()) + 1} > {n_own})")
            own_vec[rows_here] = x[msk]
        else:
            first = (rank - c * chunk) % G   # first index of this chunk with global id % G == rank
            mine = x[first::G]
            lo = (c * chunk + first) // G
            own_vec[lo:lo + mine.shape[0]] = mine
        dm = ((xf - mean) ** 2).sum(1)
        v, i = dm.min(0)
        if float(v) < med_d:
            med_d, med_i = float(v), int(gid[int(i)])
        if c % G == rank and n_gt_queries:
            if gt_fn is not None:   # (CPU dry runs of the host logic: tests/test_build_sharded_dryrun.py)
                dd, ii = gt_fn(x, gt_q_raw, min(100, n))
            else:                   # exact kNN of the chunk by the library's brute-force kernel (csrc/prep_kernels.cu)
                from . import api
                ii, dd = api.bruteforce_gt(x, gt_q_raw, min(100, n))
            cat_d = torch.cat([best_d, dd], 1); cat_i = torch.cat([best_i, gid[ii]], 1)
            sel = torch.topk(cat_d, 100, dim=1, largest=False)[1]
            best_d = torch.gather(cat_d, 1, sel); best_i = torch.gather(cat_i, 1, sel)
        del x, xf, d2p, top2, gid
    medoid = med_i
    n_virtual = N
    if by_part:   # every rank processed every chunk, so own_cnt is the same everywhere
        rows_alloc = max(own_cnt)
        if rows_alloc > n_own:
            raise RuntimeError(f"a GPU owns {rows_alloc} rows, more than the {n_own}-row buffers")
        n_own = rows_alloc                      # rows past this rank's own count are holes (no edge points at them)
        own_vec = own_vec[:n_own]
        n_virtual = G * rows_alloc
        T["n_virtual"] = n_virtual
        log(rank, f"partition ownership: rows per GPU {own_cnt} (id space {n_virtual})")
    T["generate"] = time.time() - t0
    log(rank, f"pass 1 done in {T['generate']:.1f}s; shard sizes {sh_n}; medoid {medoid}; mem {mem_gib()} GiB")

    # ground truth: merge the ranks' partial top-100
    gt_ids = gt_d = None
    if n_gt_queries:
        all_d = [torch.empty_like(best_d) for _ in range(G)]
        all_i = [torch.empty_like(best_i) for _ in range(G)]
        dist.all_gather(all_d, best_d); dist.all_gather(all_i, best_i)
        cd, ci = torch.cat(all_d, 1), torch.cat(all_i, 1)
        key = cd.double() * (1 << 31) + ci.double()           # order by (dist, id); distances are integers < 2^24
        sel = torch.argsort(key, dim=1)[:, :100]
        gt_d = torch.gather(cd, 1, sel).cpu().numpy().astype(np.float32)
        gt_ids = torch.gather(ci, 1, sel).cpu().numpy().astype(np.uint32)
        del all_d, all_i, cd, ci, key

    # ---- per shard: build, map to global ids, ship every node's list to the rank that owns its row, merge ----
    adj_own = torch.full((n_own, 64), -1, dtype=torch.int32, device=dev)
    T["build"] = 0.0; T["exchange"] = 0.0
    slice_n = 1 << 21
    for j, s in enumerate(my_shards):
        t0 = time.time()
        ns = sh_n[j]
        vec = sh_vec[j][:ns]
        stride = max(1, ns // 65536)
        sample = vec[::stride].float()
        loc_med = int(((sample - sample.mean(0)) ** 2).sum(1).argmin()) * stride   # entry point of the shard graph
        del sample
        nb = build_fn(vec, min(loc_med, ns - 1), L_build, passes, s + 1)   # int32 [ns][64] local ids
        gidt = sh_gid[j][:ns]
        sh_vec[j] = None
        del vec
        if on_gpu:
            torch.cuda.synchronize()
        T["build"] += time.time() - t0
        t0 = time.time()
        owner = (gidt % G).long()
        rounds = torch.tensor([(ns + slice_n - 1) // slice_n], device=dev)
        dist.all_reduce(rounds, op=dist.ReduceOp.MAX)
        for r_ in range(int(rounds)):
            a, b = min(ns, r_ * slice_n), min(ns, (r_ + 1) * slice_n)
            ow = owner[a:b]
            order = torch.argsort(ow, stable=True)
            send_rows = (gidt[a:b][order] // G).to(torch.int32).contiguous()
            nbs = nb[a:b][order]
            send_list = torch.where(nbs >= 0, gidt[nbs.clamp(min=0).long()], torch.full_like(nbs, -1)).contiguous()  # local -> global ids
            counts = torch.bincount(ow, minlength=G)
            rcounts = torch.empty_like(counts)
            dist.all_to_all_single(rcounts, counts)
            nrecv = int(rcounts.sum())
            recv_rows = torch.empty(nrecv, dtype=torch.int32, device=dev)
            recv_list = torch.empty((nrecv, 64), dtype=torch.int32, device=dev)
            dist.all_to_all_single(recv_rows, send_rows, rcounts.tolist(), counts.tolist())
            dist.all_to_all_single(recv_list, send_list, rcounts.tolist(), counts.tolist())
            # a node has two copies; both may arrive in one slice: merge first occurrences, then second ones
            sr, si = torch.sort(recv_rows.long(), stable=True)
            second = torch.zeros_like(sr, dtype=torch.bool)
            second[1:] = sr[1:] == sr[:-1]
            for mask in (~second, second):
                idx = si[mask]
                if idx.numel():
                    merge_lists(adj_own, recv_rows[idx], recv_list[idx], G, rank)
            del send_rows, send_list, recv_rows, recv_list, order
        del nb, owner
        sh_gid[j] = None
        if on_gpu:
            torch.cuda.empty_cache()
        T["exchange"] += time.time() - t0
        log(rank, f"shard {s}: {ns} points, build {T['build']:.1f}s exchange {T['exchange']:.1f}s (cumulative); mem {mem_gib()} GiB")
    deg = (adj_own >= 0).sum(1)
    log(rank, f"merged graph: mean degree {float(deg.float().mean()):.2f}, min {int(deg.min())}, full rows {float((deg == 64).float().mean()):.3f}")

    # ---- hand the rows and the codes to the search library ----
    t0 = time.time()
    del deg
    if on_gpu:
        torch.cuda.empty_cache()   # the library allocates with cudaMalloc: hand torch's cached blocks back first (1e9 points: 80 GB needed)
    search.load_device_begin(n_virtual, D, medoid, piv.cpu().numpy(), cen.cpu().numpy(), offs)
    step = 1 << 22
    for a in range(0, n_own, step):
        b = min(n_own, a + step)
        search.load_device_rows(a, b - a, own_vec[a:b].data_ptr(), adj_own[a:b].data_ptr())
    del own_vec, adj_own
    if on_gpu:
        torch.cuda.empty_cache()
    piv_np, cen_np = piv.cpu().numpy(), cen.cpu().numpy()
    for c0 in range(0, n_chunks, G):   # G chunks at a time: every rank encodes one of them, then G broadcasts
        mine = None
        c_me = c0 + rank
        if c_me < n_chunks:
            n_me = min(chunk, N - c_me * chunk)
            mine = synth.encode_pq(gen_chunk(centers, c_me, n_me, seed), piv_np, cen_np, offs)
        for c in range(c0, min(n_chunks, c0 + G)):
            n = min(chunk, N - c * chunk)
            codes = mine if c == c_me else torch.empty((n, m), dtype=torch.uint8, device=dev)
            dist.broadcast(codes, c % G)
            if by_part:   # the chunk's points carry scattered ids
                search.load_device_codes_at(gid_all[c * chunk:c * chunk + n].data_ptr(), n, codes.data_ptr())
            else:
                search.load_device_codes(c * chunk, n, codes.data_ptr())
        if on_gpu:
            torch.cuda.synchronize()
        del mine, codes
    search.load_device_end()
    T["load"] = time.time() - t0
    if by_part:   # every query goes to the GPU that owns the partition it falls into
        mine = (T["home"] == rank).nonzero().squeeze(1)
        T["my_idx"] = mine.numpy()
        my_q = queries[mine.to(dev)].cpu().numpy()
    else:
        T["my_idx"] = np.arange(rank * n_queries_per_rank, (rank + 1) * n_queries_per_rank)
        my_q = queries[rank * n_queries_per_rank:(rank + 1) * n_queries_per_rank].cpu().numpy()
    return my_q, gt_ids, gt_d, medoid, T
