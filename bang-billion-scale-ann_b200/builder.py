"""Fixture index builder: synthetic data -> files in the reference's formats.

Small shapes (tests, C1) use the host C++ Vamana builder (csrc/fixture_builder.cpp).  The reference
consumes DiskANN output and has no builder of its own (README.md:46-58); see SURVEY.md §8(f) rank 1.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import build as _build
from . import formats, synth

_DT = {"int8": 0, "uint8": 1, "float": 2}


def build_vamana_cpu(vectors: np.ndarray, R: int = 64, L: int = 100, alpha: float = 1.2, seed: int = 1,
                     nthreads: int = 0, passes: int = 2):
    """Returns (degrees u32[N], nbrs u32[N][R] ascending, medoid)."""
    lib = ctypes.CDLL(_build.build_fixture())
    fn = lib.bang_fixture_build_vamana
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                   ctypes.c_float, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.c_void_p]
    vectors = np.ascontiguousarray(vectors)
    N, D = vectors.shape
    deg = np.zeros(N, dtype=np.uint32)
    nbrs = np.zeros((N, R), dtype=np.uint32)
    med = ctypes.c_uint64(0)
    rc = fn(vectors.ctypes.data, _DT[formats.dtype_name(vectors)], N, D, R, L, alpha, seed, nthreads, passes,
            deg.ctypes.data, nbrs.ctypes.data, ctypes.byref(med))
    if rc != 0:
        raise RuntimeError("bang_fixture_build_vamana failed")
    return deg, nbrs, int(med.value)


def make_fixture(prefix: str, n: int, d: int, dtype: str, nq: int, m: int | None, k_gt: int = 100, R: int = 64,
                 L_build: int = 100, alpha: float = 1.2, nthreads: int = 0, n_clusters: int | None = None,
                 seed: int = synth.BASE_SEED, passes: int = 2) -> dict:
    """Generate data + queries + graph + PQ + ground truth and write every file (both pivot layouts)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)) or ".", exist_ok=True)
    base, centers = synth.make_clustered(n, d, dtype, seed=seed, n_clusters=n_clusters)
    queries, _ = synth.make_clustered(nq, d, dtype, seed=synth.QUERY_SEED ^ seed, centers=centers)
    base_np = base.numpy()
    deg, nbrs, medoid = build_vamana_cpu(base_np, R=R, L=L_build, alpha=alpha, nthreads=nthreads, passes=passes)
    piv = cen = offs = codes = None
    if m is not None:
        piv, cen, offs = synth.train_pq(base, m)
        codes = synth.encode_pq(base, piv, cen, offs).numpy()
    paths = formats.write_index(prefix, base_np, deg, nbrs, medoid, piv, cen, offs, codes)
    formats.write_bin(paths.query, queries.numpy())
    gt_ids, gt_d = synth.brute_force_gt(base, queries, min(k_gt, n))
    formats.write_truthset(paths.truth, gt_ids, gt_d)
    return dict(paths=paths, base=base_np, queries=queries.numpy(), deg=deg, nbrs=nbrs, medoid=medoid, pivots=piv,
                centroid=cen, chunk_offsets=offs, codes=codes, gt_ids=gt_ids, gt_dists=gt_d, dtype=dtype, R=R)


def build_vamana_gpu(base: torch.Tensor, medoid: int, L: int = 64, alpha: float = 1.2, passes: int = 2, seed: int = 1,
                     max_batch: int = 0, device_out: bool = False):
    """Batch-parallel Vamana on the GPU (csrc/builder.cu).  base: CUDA tensor [N][D] uint8/int8/float32.
    device_out: return the adjacency as a CUDA int32 tensor [N][64] (unused = -1) instead of host arrays."""
    assert base.is_cuda and base.is_contiguous()
    lib = ctypes.CDLL(_build.build_cuda())
    fn = lib.bang_b200_build_vamana
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float,
                   ctypes.c_uint64, ctypes.c_float, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.bang_b200_builder_last_error.restype = ctypes.c_char_p
    N, D = base.shape
    dt = {torch.uint8: "uint8", torch.int8: "int8", torch.float32: "float"}[base.dtype]
    g = torch.Generator(device=base.device)
    g.manual_seed(seed)
    order = torch.cat([torch.randperm(N, generator=g, device=base.device) for _ in range(passes)]).to(torch.int32).contiguous()
    n_first = N if passes > 1 else 0
    stats = np.zeros(4, dtype=np.float32)
    torch.cuda.synchronize(base.device)
    if device_out:
        d_nbrs = torch.empty((N, 64), dtype=torch.int32, device=base.device)
        with torch.cuda.device(base.device):
            rc = fn(_DT[dt], base.data_ptr(), N, D, L, 1.0, n_first, alpha, order.data_ptr(), order.numel(), medoid,
                    (max_batch & 0x7FFFFFFF) | 0x80000000, None, d_nbrs.data_ptr(), stats.ctypes.data)
        if rc != 0:
            raise RuntimeError("bang_b200_build_vamana: " + lib.bang_b200_builder_last_error().decode())
        if os.environ.get("BANG_B200_BUILD_TIMERS"):
            print("[builder] batches %d search %.0f ms prune %.0f ms reverse %.0f ms" % tuple(stats.tolist()), flush=True)
        return d_nbrs
    deg = np.zeros(N, dtype=np.uint32)
    nbrs = np.zeros((N, 64), dtype=np.uint32)
    with torch.cuda.device(base.device):
        rc = fn(_DT[dt], base.data_ptr(), N, D, L, 1.0, n_first, alpha, order.data_ptr(), order.numel(), medoid, max_batch,
                deg.ctypes.data, nbrs.ctypes.data, stats.ctypes.data)
    if rc != 0:
        raise RuntimeError("bang_b200_build_vamana: " + lib.bang_b200_builder_last_error().decode())
    return deg, nbrs, stats


def find_medoid(base: torch.Tensor) -> int:
    n = base.shape[0]
    mean = torch.zeros(base.shape[1], dtype=torch.float64, device=base.device)
    for s in range(0, n, 1 << 20):
        mean += base[s:s + (1 << 20)].double().sum(0)
    mean = (mean / n).float()
    best, arg = float("inf"), 0
    for s in range(0, n, 1 << 20):
        d = ((base[s:s + (1 << 20)].float() - mean) ** 2).sum(1)
        v, i = d.min(0)
        if float(v) < best:
            best, arg = float(v), s + int(i)
    return arg


def make_fixture_auto(prefix: str, n: int, d: int, dtype: str, nq: int, m: int | None, k_gt: int = 100, device="cpu",
                      builder: str = "auto", L_build: int = 64, alpha: float = 1.2, seed: int = synth.BASE_SEED,
                      n_gt_queries: int | None = None) -> dict:
    """Like make_fixture, but everything (data, graph, PQ, ground truth) is produced on `device` when it is a GPU.
    n_gt_queries: ground truth for the first n_gt_queries queries only (bench.py: one batch of the query file)."""
    import time
    dev = torch.device(device)
    use_gpu = dev.type == "cuda" and builder in ("auto", "gpu")
    if builder == "gpu" and dev.type != "cuda":
        raise RuntimeError("GPU builder requested without a CUDA device")
    os.makedirs(os.path.dirname(os.path.abspath(prefix)) or ".", exist_ok=True)
    t = {}
    t0 = time.time()
    gen_dev = dev if dev.type == "cuda" else "cpu"
    base, centers = synth.make_clustered(n, d, dtype, seed=seed, device=gen_dev)
    queries, _ = synth.make_clustered(nq, d, dtype, seed=synth.QUERY_SEED ^ seed, centers=centers, device=gen_dev)
    t["data"] = time.time() - t0
    t0 = time.time()
    if use_gpu:
        medoid = find_medoid(base)
        deg, nbrs, _ = build_vamana_gpu(base, medoid, L=L_build, alpha=alpha)
    else:
        deg, nbrs, medoid = build_vamana_cpu(base.cpu().numpy(), L=max(L_build, 64), alpha=alpha)
    t["graph"] = time.time() - t0
    t0 = time.time()
    piv = cen = offs = codes = None
    if m is not None:
        piv, cen, offs = synth.train_pq(base, m)
        codes = synth.encode_pq(base, piv, cen, offs).cpu().numpy()
    t["pq"] = time.time() - t0
    t0 = time.time()
    gt_ids, gt_d = synth.brute_force_gt(base, queries if n_gt_queries is None else queries[:n_gt_queries], min(k_gt, n))
    t["gt"] = time.time() - t0
    t0 = time.time()
    paths = formats.write_index(prefix, base.cpu().numpy(), deg, nbrs, medoid, piv, cen, offs, codes)
    formats.write_bin(paths.query, queries.cpu().numpy())
    formats.write_truthset(paths.truth, gt_ids, gt_d)
    t["write"] = time.time() - t0
    return dict(builder="gpu" if use_gpu else "cpu", medoid=int(medoid), mean_degree=float(deg.mean()),
                seconds={k: round(v, 2) for k, v in t.items()})
