"""GPU parity for PQ chunk layouts the committed fixtures do not reach (they all have 4-dim chunks and m <= 10, i.e.
the CS = 4 kernel): m > 32 (two or more 32-chunk groups per code row — the reference's SIFT1B build has 74 chunks,
parANN.h:101, SIFT10K 128, :87), uneven chunk sizes (CS = 0 kernel) and 3-dim chunks (CS = 3 kernel).
Same bar as test_gpu_parity: ids, distance bits and counters identical to the oracle.

The oracle side of these shapes is covered on CPU by test_oracle_pq_shapes.py."""
import os

import numpy as np
import pytest

from bang_b200 import api, builder, formats

import oracle as O

pytestmark = [pytest.mark.gpu]

MODE_O = {"base": O.MODE_BASE, "inmemory": O.MODE_INMEMORY}


@pytest.mark.parametrize("D,m,dtype", [(128, 74, "uint8"), (128, 128, "uint8"), (50, 20, "uint8"), (96, 32, "float"), (24, 8, "int8"),
                                       (128, 32, "int8")])
def test_search_bit_exact_general_chunking(tmp_path, D, m, dtype):
    import torch
    prefix = str(tmp_path / "idx")
    nq = 64
    builder.make_fixture_auto(prefix, 20_000, D, dtype, nq, m, k_gt=10, device=torch.device("cuda", 0))
    npdt = {"uint8": np.uint8, "int8": np.int8, "float": np.float32}[dtype]
    queries = formats.read_bin(prefix + "_query.bin", npdt)
    ox = O.OracleIndex.from_files(prefix)
    for mode in ("inmemory", "base"):
        s = api.BANGSearch(dtype, mode)
        assert s.bang_load(prefix)
        s.set_dists_layout(api.DISTS_QUERY_MAJOR)
        for L in (16, 72):
            s.bang_set_searchparams(10, L)
            s.bang_alloc(nq)
            s.bang_init(nq)
            ids, dists = s.bang_query(queries)
            st = s.last_stats(nq)
            s.bang_free()
            oids, od, ost = ox.search(queries, 10, L, mode=MODE_O[mode], order=O.ORDER_GPU, stats=True)
            assert np.array_equal(ids, oids), f"{mode} L={L}: {(ids != oids).any(1).sum()} of {nq} queries differ"
            assert np.array_equal(dists.view(np.uint32), od.view(np.uint32))
            assert np.array_equal(st["n_cand"], ost["n_cand"]) and np.array_equal(st["hops"], ost["hops"])
        s.bang_unload()
