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
