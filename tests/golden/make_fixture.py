"""Generates the small committed fixtures tests/golden/fx_*.npz (run once in the build container).

    python tests/golden/make_fixture.py

Each npz holds a complete index in array form (base vectors, R=64 Vamana graph, PQ pivots/centroid/
chunk offsets/codes, queries, brute-force ground truth).  tests/conftest.py materialises them into the
reference's file formats with bang_b200.formats.write_index.  The builder runs single-threaded so the
output is reproducible.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import bang_b200  # noqa: E402,F401
from bang_b200 import builder  # noqa: E402

SPECS = {
    # name: (n, d, dtype, nq, m, L_build)
    "fx_u8": (3000, 32, "uint8", 64, 8, 64),
    "fx_f32": (2000, 24, "float", 48, 6, 64),
    "fx_i8": (1500, 40, "int8", 32, 10, 48),
}


def main():
    import tempfile
    for name, (n, d, dt, nq, m, lb) in SPECS.items():
        with tempfile.TemporaryDirectory() as tmp:
            fx = builder.make_fixture(os.path.join(tmp, name), n, d, dt, nq, m, k_gt=32, L_build=lb, nthreads=1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), base=fx["base"], deg=fx["deg"], nbrs=fx["nbrs"],
                            medoid=np.uint64(fx["medoid"]), pivots=fx["pivots"], centroid=fx["centroid"],
                            chunk_offsets=fx["chunk_offsets"], codes=fx["codes"], queries=fx["queries"],
                            gt_ids=fx["gt_ids"], gt_dists=fx["gt_dists"])
        print(name, "deg mean", fx["deg"].mean(), "medoid", fx["medoid"], os.path.getsize(os.path.join(HERE, name + ".npz")))


def make_c1():
    """C1 (BASELINE config 1): the shape of the reference's stock SIFT10K index — N = 10^4, D = 128 uint8, 100 queries,
    entry length 388 B — with two PQ layouts: m = 32 (C2's) and m = 128 (the reference's SIFT1BSMALL block,
    BANG_Inmemory/parANN.h:81-91).  The stock files (sift10kfiles.tar.gz) are absent from the mount."""
    import tempfile
    from bang_b200 import synth
    n, d, nq = 10_000, 128, 100
    with tempfile.TemporaryDirectory() as tmp:
        fx = builder.make_fixture(os.path.join(tmp, "c1"), n, d, "uint8", nq, 32, k_gt=32, L_build=100, nthreads=1)
    import torch
    base = torch.from_numpy(fx["base"])
    piv128, cen128, offs128 = synth.train_pq(base, 128)
    codes128 = synth.encode_pq(base, piv128, cen128, offs128).numpy()
    out = os.path.join(HERE, "fx_c1.npz")
    np.savez_compressed(out, base=fx["base"], deg=fx["deg"], nbrs=fx["nbrs"], medoid=np.uint64(fx["medoid"]),
                        pivots=fx["pivots"], centroid=fx["centroid"], chunk_offsets=fx["chunk_offsets"], codes=fx["codes"],
                        pivots128=piv128, centroid128=cen128, chunk_offsets128=offs128, codes128=codes128,
                        queries=fx["queries"], gt_ids=fx["gt_ids"], gt_dists=fx["gt_dists"])
    print("fx_c1 deg mean", fx["deg"].mean(), "medoid", fx["medoid"], os.path.getsize(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "c1":
        make_c1()
    else:
        main()
