"""Synthetic clustered datasets, PQ training/encoding and brute-force ground truth.

The reference ships no dataset (sift10kfiles.tar.gz is absent) and relies on DiskANN for
PQ pivots and ground truth (README.md:46-58).  This module produces inputs of each BASELINE.json shape
with fixed seeds.  It is fixture tooling, not the search path: torch is used as a plain
array library (CPU here, CUDA on the GPU box for the 10^6+ shapes).

PQ follows DiskANN's scheme as consumed by the reference (bang_search.cu:1083-1130): the data mean
("centroid") is subtracted, D dims are split into m contiguous chunks, 256 k-means centres per chunk,
pivots stored as float[256][D].
"""
from __future__ import annotations

import os

import numpy as np
import torch

BASE_SEED = 0xBA5E
QUERY_SEED = 0x9E41
PQ_SEED = 0x50


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


FLOAT_LATENT_DIM = 32      # intrinsic dimension of the float shapes (GIST/DEEP-like data are low-dimensional)
FLOAT_LATENT_SIGMA = 0.35  # within-cluster spread in the latent space (cluster centres ~ N(0, 1))
FLOAT_AMBIENT_SIGMA = 0.02 # isotropic noise added in the full space


def _float_embedding(d: int, device) -> torch.Tensor:
    """Fixed [latent][d] matrix with orthonormal rows: latent-space distances are preserved in the full space."""
    lat = min(d, FLOAT_LATENT_DIM)
    g = torch.Generator(device="cpu")
    g.manual_seed(BASE_SEED ^ 0xE3B)
    q, _ = torch.linalg.qr(torch.randn(d, lat, generator=g, dtype=torch.float64))
    return q.T.contiguous().float().to(device)


def make_clustered(n: int, d: int, dtype: str, seed: int = BASE_SEED, n_clusters: int | None = None,
                   device="cpu", centers: torch.Tensor | None = None):
    """Gaussian mixture; returns (points [n][d] torch tensor of dtype, centers).

    uint8 / int8 shapes (SIFT-like): centres ~ U[32, 224]^d, sigma 24, rounded and clipped.
    float shapes (GIST/DEEP-like): a mixture in a 32-dimensional latent space (centres ~ N(0,1), sigma 0.35)
    embedded isometrically into d dimensions plus 0.02 ambient noise.  SURVEY §8d proposed an isotropic
    d-dimensional mixture; at d = 960 all inter-cluster distances concentrate and NO graph search (the
    reference's included) can navigate it (recall@10 < 1 % at L = 96, measured), so the float shapes use the
    low-intrinsic-dimension variant real descriptor datasets resemble.
    """
    if n_clusters is None:
        n_clusters = max(16, n // 1000)
    g = _gen(seed, device)
    is_float = dtype == "float"
    if is_float:
        emb = _float_embedding(d, device)
        if centers is None:
            centers = torch.randn(n_clusters, emb.shape[0], generator=_gen(BASE_SEED ^ 0xC0, device), device=device)
        sigma = FLOAT_LATENT_SIGMA
    else:
        if centers is None:
            centers = torch.rand(n_clusters, d, generator=_gen(BASE_SEED ^ 0xC0, device), device=device) * 192.0 + 32.0
        sigma = 24.0
    out_dtype = {"float": torch.float32, "uint8": torch.uint8, "int8": torch.int8}[dtype]
    out = torch.empty(n, d, dtype=out_dtype, device=device)
    step = 1 << 18 if is_float else 1 << 20
    for s in range(0, n, step):
        e = min(n, s + step)
        assign = torch.randint(0, centers.shape[0], (e - s,), generator=g, device=device)
        pts = centers[assign] + sigma * torch.randn(e - s, centers.shape[1], generator=g, device=device)
        if is_float:
            pts = pts @ emb + FLOAT_AMBIENT_SIGMA * torch.randn(e - s, d, generator=g, device=device)
        elif dtype == "uint8":
            pts = pts.round().clamp_(0, 255)
        else:
            pts = (pts - 128.0).round().clamp_(-128, 127)
        out[s:e] = pts.to(out_dtype)
    return out, centers


def chunk_offsets_even(d: int, m: int) -> np.ndarray:
    """DiskANN splits D dims into m chunks as evenly as possible (low chunks get the extra dim)."""
    lo, extra = divmod(d, m)
    sizes = [lo + (1 if c < extra else 0) for c in range(m)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)


@torch.no_grad()
def train_pq(base: torch.Tensor, m: int, iters: int = 12, max_train: int = 65536, seed: int = PQ_SEED):
    """Returns pivots float32[256][D], centroid float32[D], chunk_offsets u32[m+1] (numpy).
    CUDA tensors go through the library's k-means kernels (csrc/prep_kernels.cu); the torch code below is the host
    path of the small committed fixtures (tests/golden/make_fixture.py) and the reference the kernels are tested against."""
    if base.is_cuda and not os.environ.get("BANG_B200_TORCH_PREP"):
        from . import api
        offs = chunk_offsets_even(base.shape[1], m)
        piv, cen = api.pq_train(base.contiguous(), offs, iters=iters, max_train=max_train, seed=seed)
        return piv, cen, offs
    dev = base.device
    n, d = base.shape
    g = _gen(seed, dev)
    if n > max_train:
        idx = torch.randperm(n, generator=g, device=dev)[:max_train]
        train = base[idx].float()
    else:
        train = base.float()
    # centroid over the whole base in fp64 for stability
    centroid = torch.zeros(d, dtype=torch.float64, device=dev)
    for s in range(0, n, 1 << 20):
        centroid += base[s:s + (1 << 20)].double().sum(0)
    centroid = (centroid / n).float()
    train = train - centroid
    offs = chunk_offsets_even(d, m)
    pivots = torch.zeros(256, d, dtype=torch.float32, device=dev)
    nt = train.shape[0]
    for c in range(m):
        a, b = int(offs[c]), int(offs[c + 1])
        x = train[:, a:b].contiguous()
        if nt >= 256:
            init = torch.randperm(nt, generator=g, device=dev)[:256]
            cen = x[init].clone()
        else:  # tiny fixtures: pad with jittered copies
            rep = x[torch.randint(0, nt, (256,), generator=g, device=dev)]
            cen = rep + 1e-3 * torch.randn(256, b - a, generator=g, device=dev)
        for _ in range(iters):
            d2 = torch.cdist(x, cen) if x.shape[1] > 0 else torch.zeros(nt, 256, device=dev)
            lab = d2.argmin(1)
            sums = torch.zeros_like(cen).index_add_(0, lab, x)
            cnt = torch.zeros(256, device=dev).index_add_(0, lab, torch.ones(nt, device=dev))
            nz = cnt > 0
            cen[nz] = sums[nz] / cnt[nz, None]
        pivots[:, a:b] = cen
    return pivots.cpu().numpy(), centroid.cpu().numpy(), offs


@torch.no_grad()
def encode_pq(base: torch.Tensor, pivots: np.ndarray, centroid: np.ndarray, offs: np.ndarray) -> torch.Tensor:
    if base.is_cuda and not os.environ.get("BANG_B200_TORCH_PREP"):
        from . import api
        return api.pq_encode(base.contiguous(), pivots, centroid, offs)
    dev = base.device
    piv = torch.from_numpy(pivots).to(dev)
    cen = torch.from_numpy(centroid).to(dev)
    n = base.shape[0]
    m = len(offs) - 1
    codes = torch.empty(n, m, dtype=torch.uint8, device=dev)
    step = 1 << 17
    for s in range(0, n, step):
        x = base[s:s + step].float() - cen
        for c in range(m):
            a, b = int(offs[c]), int(offs[c + 1])
            xc = x[:, a:b]
            pc = piv[:, a:b]
            # argmin ||x - p||^2 = argmin (||p||^2 - 2 x.p)
            sc = (pc * pc).sum(1)[None, :] - 2.0 * (xc @ pc.T)
            codes[s:s + step, c] = sc.argmin(1).to(torch.uint8)
    return codes


@torch.no_grad()
def brute_force_gt(base: torch.Tensor, queries: torch.Tensor, k: int, block: int = 1 << 18, q_block: int = 8192):
    """Exact kNN by squared L2.  Returns (ids u32 [nq][k], dists f32 [nq][k]) as numpy, ties ordered by id.

    Candidates are found with the matmul expansion in fp32 over blocks (over-fetching 2k per block),
    then the final k are re-scored with direct (a-b)^2 sums so distances are the exact ones the
    search kernels produce (integers for u8/i8 inputs).  Queries are processed q_block at a time so the
    distance tile stays at q_block x block floats (8 GB).
    CUDA tensors go through the library's brute-force kernels (csrc/prep_kernels.cu); this torch code is the host path.
    """
    if base.is_cuda and not os.environ.get("BANG_B200_TORCH_PREP"):
        from . import api
        ids, d = api.bruteforce_gt(base.contiguous(), queries.contiguous(), k)
        return ids.cpu().numpy().astype(np.uint32), d.cpu().numpy()
    if queries.shape[0] > q_block:
        parts = [brute_force_gt(base, queries[s:s + q_block], k, block, q_block) for s in range(0, queries.shape[0], q_block)]
        return np.concatenate([p[0] for p in parts], 0), np.concatenate([p[1] for p in parts], 0)
    dev = base.device
    nq = queries.shape[0]
    q = queries.float()
    qn = (q * q).sum(1)
    kk = min(base.shape[0], 2 * k + 8)
    best_d = torch.full((nq, kk), float("inf"), device=dev)
    best_i = torch.zeros((nq, kk), dtype=torch.int64, device=dev)
    for s in range(0, base.shape[0], block):
        b = base[s:s + block].float()
        d2 = qn[:, None] + (b * b).sum(1)[None, :] - 2.0 * (q @ b.T)
        kb = min(kk, b.shape[0])
        dd, ii = torch.topk(d2, kb, dim=1, largest=False)
        del d2
        cat_d = torch.cat([best_d, dd], 1)
        cat_i = torch.cat([best_i, ii + s], 1)
        sel = torch.topk(cat_d, kk, dim=1, largest=False)[1]
        best_d = torch.gather(cat_d, 1, sel)
        best_i = torch.gather(cat_i, 1, sel)
    # exact re-score (in slices: nq x kk x d floats can be large for d = 960)
    exact = torch.empty(nq, kk, device=dev)
    for s in range(0, nq, 512):
        cand = base[best_i[s:s + 512].reshape(-1)].float().reshape(-1, kk, base.shape[1])
        diff = cand - q[s:s + 512, None, :]
        exact[s:s + 512] = (diff * diff).sum(2)
    # order by (dist, id)
    order = torch.argsort(best_i, dim=1, stable=True)
    exact = torch.gather(exact, 1, order)
    ids = torch.gather(best_i, 1, order)
    order = torch.argsort(exact, dim=1, stable=True)
    exact = torch.gather(exact, 1, order)[:, :k]
    ids = torch.gather(ids, 1, order)[:, :k]
    return ids.cpu().numpy().astype(np.uint32), exact.cpu().numpy().astype(np.float32)
