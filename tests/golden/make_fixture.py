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


if __name__ == "__main__":
    main()
