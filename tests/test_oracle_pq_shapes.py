"""The oracle's PQ arithmetic for chunk layouts beyond the committed fixtures (which all have 4-dim chunks and m <= 10):
m > 32 (the reference's SIFT1B build uses 74 chunks, parANN.h:101; SIFT10K 128, :87), uneven chunk sizes, 3-dim
chunks.  Checked against a float64 evaluation of the reference formulas (bang_search.cu:1118-1129, 1201-1241)."""
import numpy as np
import pytest
import torch

from bang_b200 import formats, synth

import oracle as O


@pytest.mark.parametrize("D,m,dtype", [(128, 74, "uint8"), (128, 128, "uint8"), (50, 20, "uint8"), (96, 32, "float"), (24, 8, "int8")])
def test_oracle_pq_distance_general_chunking(D, m, dtype):
    rng = np.random.default_rng(D * 1000 + m)
    n = 600
    if dtype == "float":
        base = rng.normal(size=(n, D)).astype(np.float32)
    elif dtype == "uint8":
        base = rng.integers(0, 256, size=(n, D), dtype=np.uint8)
    else:
        base = rng.integers(-128, 128, size=(n, D), dtype=np.int8)
    piv, cen, offs = synth.train_pq(torch.from_numpy(base), m, iters=3)
    assert len(offs) == m + 1 and offs[-1] == D
    codes = synth.encode_pq(torch.from_numpy(base), piv, cen, offs).numpy()
    nbrs = np.zeros((n, 64), np.uint32)
    deg = np.ones(n, np.uint32)
    ox = O.OracleIndex(formats.pack_disk_bin(base, deg, nbrs), dtype, D, 64, 0, codes, piv, cen, offs)
    q = base[7]
    tbl = ox.pq_table(q)
    assert tbl.shape == (m, 256)
    q64 = q.astype(np.float64) - cen.astype(np.float64)
    ref_tbl = np.stack([((piv[:, a:b].astype(np.float64) - q64[a:b]) ** 2).sum(1) for a, b in zip(offs[:-1], offs[1:])])
    assert np.allclose(tbl, ref_tbl, rtol=1e-5, atol=1e-4)
    for node in (0, 7, 311, n - 1):
        want = ref_tbl[np.arange(m), codes[node]].sum()
        assert ox.pq_dist(tbl, node) == pytest.approx(want, rel=1e-5, abs=1e-3)
