"""Three degenerate graphs (shared by the oracle tests on CPU and the GPU parity tests): an isolated entry point,
fewer points than k, and a dead end (a node without out-neighbours) on the only path."""
import numpy as np
import torch

from bang_b200 import synth

NO_ID = 0xFFFFFFFF
QUERY = np.array([[9, 0, 3, 0]], np.uint8)


def _toy(n, D=4, R=64):
    base = np.zeros((n, D), dtype=np.uint8)
    base[:, 0] = np.arange(n) * 3
    base[:, 2] = (np.arange(n) * 7) % 11
    return base, np.zeros(n, np.uint32), np.zeros((n, R), np.uint32)


def _pq(base):
    piv, cen, offs = synth.train_pq(torch.from_numpy(base), 2, iters=2)
    return piv, cen, offs, synth.encode_pq(torch.from_numpy(base), piv, cen, offs).numpy()


def cases():
    """name -> (base, deg, nbrs, medoid, pivots, centroid, chunk_offsets, codes, k, L)"""
    out = {}
    base, deg, nbrs = _toy(20)              # entry point 0 has no edges; the rest is a ring it cannot reach
    for i in range(1, 20):
        deg[i] = 1
        nbrs[i, 0] = (i % 19) + 1
    out["isolated_entry"] = (base, deg, nbrs, 0, *_pq(base), 3, 5)
    base, deg, nbrs = _toy(5)               # complete graph on 5 points, k = 10
    for i in range(5):
        deg[i] = 4
        nbrs[i, :4] = [j for j in range(5) if j != i]
    out["fewer_points_than_k"] = (base, deg, nbrs, 2, *_pq(base), 10, 10)
    base, deg, nbrs = _toy(6)               # 0 -> 1 -> 2, node 2 is a dead end, 3..5 unreachable
    deg[0] = 1; nbrs[0, 0] = 1
    deg[1] = 1; nbrs[1, 0] = 2
    out["dead_end"] = (base, deg, nbrs, 0, *_pq(base), 4, 4)
    return out
