"""Rows with the slot block (precomputed visited-filter slots of every neighbour, search_kernel.cuh kSlotBytes) against rows
without it and against the oracle: the two layouts run different kernel instantiations (`PH`) and different load kernels,
the answers must be the same bits.  The slots replace hashFn1_d / hashFn2_d + the modulus (bang_search.cu:1168-1189) at
search time, so any slip in the load-time hash or in the word format shows up as a changed candidate count."""
import os

import numpy as np
import pytest

from bang_b200 import api, formats

import oracle as O

pytestmark = pytest.mark.gpu

MODE_O = {"base": O.MODE_BASE, "inmemory": O.MODE_INMEMORY}


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _run(s, queries, k, L):
    s.set_dists_layout(api.DISTS_QUERY_MAJOR)
    s.bang_set_searchparams(k, L)
    s.bang_alloc(len(queries))
    s.bang_init(len(queries))
    ids, d = s.bang_query(queries)
    st = s.last_stats(len(queries))
    info = s.info()
    s.bang_free()
    s.bang_unload()
    return ids.copy(), d.copy(), st, info


@pytest.mark.parametrize("mode", ["base", "inmemory"])
@pytest.mark.parametrize("L", [20, 152])
def test_slot_block_rows_equal_plain_rows_and_oracle(fx_c1, mode, L):
    """C1 shape (D = 128 u8, m = 32: the 32-uniform-chunk kernels).  The slot block is opt-in (BANG_B200_PREHASH=1)."""
    res = {}
    for setting in ("0", "1", None):
        with _env(BANG_B200_PREHASH=setting):
            s = api.BANGSearch(fx_c1.dtype, mode)
            assert s.bang_load(fx_c1.prefix)
            res[setting] = _run(s, fx_c1.queries, 10, L)
    assert res["0"][3].slot_block == 0 and res["1"][3].slot_block == 1 and res[None][3].slot_block == 0
    assert res["1"][3].row_stride == res["0"][3].row_stride + 512
    oids, od, ost = fx_c1.oracle().search(fx_c1.queries, 10, L, mode=MODE_O[mode], order=O.ORDER_GPU, stats=True)
    for setting, (ids, d, st, _) in res.items():
        assert np.array_equal(ids, oids), (setting, int((ids != oids).any(1).sum()))
        assert np.array_equal(d.view(np.uint32), od.view(np.uint32)), setting
        assert np.array_equal(st["hops"], ost["hops"]) and np.array_equal(st["n_cand"], ost["n_cand"]) and np.array_equal(st["sum_deg"], ost["sum_deg"]), setting


def test_slot_block_is_for_the_uniform_chunk_pq_kernels_only(fx_u8, fx_c1):
    """m = 8 fixtures (general-chunk kernels) and Exactdistance never carry the block, whatever the switch says."""
    with _env(BANG_B200_PREHASH="1"):
        s = api.BANGSearch(fx_u8.dtype, "inmemory")
        assert s.bang_load(fx_u8.prefix)
        assert s.info().slot_block == 0
        s.bang_unload()
        s = api.BANGSearch(fx_c1.dtype, "exact")
        assert s.bang_load(fx_c1.prefix)
        assert s.info().slot_block == 0
        s.bang_unload()


def test_slot_block_device_resident_load(fx_c1):
    """bang_b200_load_device_rows writes the block too (pack_rows_kernel)."""
    import torch
    fx = fx_c1
    dev = torch.device("cuda", 0)
    vec = torch.from_numpy(np.ascontiguousarray(fx.base)).to(dev)
    adj_np = fx.nbrs.astype(np.uint32).copy()
    adj_np[np.arange(fx.R)[None, :] >= fx.deg[:, None]] = 0xFFFFFFFF
    adj = torch.from_numpy(adj_np.view(np.int32)).to(dev)
    codes = torch.from_numpy(np.ascontiguousarray(fx.codes)).to(dev)
    out = {}
    for setting in ("0", "1"):
        with _env(BANG_B200_PREHASH=setting):
            s = api.BANGSearch(fx.dtype, "inmemory")
            s.load_device_begin(fx.N, fx.D, fx.medoid, fx.pivots, fx.centroid, fx.chunk_offsets)
            s.load_device_rows(0, fx.N, vec.data_ptr(), adj.data_ptr())
            s.load_device_codes(0, fx.N, codes.data_ptr())
            s.load_device_end()
            torch.cuda.synchronize()
            out[setting] = _run(s, fx.queries, 10, 64)
    assert out["0"][3].slot_block == 0 and out["1"][3].slot_block == 1
    oids, od = fx.oracle().search(fx.queries, 10, 64, mode=O.MODE_INMEMORY, order=O.ORDER_GPU)
    for setting in out:
        assert np.array_equal(out[setting][0], oids) and np.array_equal(out[setting][1].view(np.uint32), od.view(np.uint32))


@pytest.mark.parametrize("mode,L", [("inmemory", 400), ("base", 300)])
def test_slot_block_with_spilled_filter_blocks(tmp_path, mode, L):
    """The heavy-filter case of test_visited_filter_spill_blocks on the slot-block kernels: a random 64-regular graph at
    D = 128, m = 32, where most filter blocks overflow into their bitmaps."""
    import torch
    from bang_b200 import synth
    rng = np.random.default_rng(78)
    N, D, R, m = 40_000, 128, 64, 32
    base = rng.integers(0, 256, size=(N, D), dtype=np.uint8)
    r = rng.integers(0, N - 1, size=(N, R))
    while True:
        srt = np.sort(r, axis=1)
        dup = np.nonzero((srt[:, 1:] == srt[:, :-1]).any(1))[0]
        if len(dup) == 0:
            break
        r[dup] = rng.integers(0, N - 1, size=(len(dup), R))
    nbrs = ((np.arange(N)[:, None] + 1 + r) % N).astype(np.uint32)
    deg = np.full(N, R, dtype=np.uint32)
    piv, cen, offs = synth.train_pq(torch.from_numpy(base), m, iters=3)
    codes = synth.encode_pq(torch.from_numpy(base), piv, cen, offs).numpy()
    prefix = str(tmp_path / "rnd32")
    formats.write_index(prefix, base, deg, nbrs, 321, piv, cen, offs, codes)
    queries = rng.integers(0, 256, size=(16, D), dtype=np.uint8)
    ox = O.OracleIndex(formats.pack_disk_bin(base, deg, nbrs), "uint8", D, R, 321, codes, piv, cen, offs)
    oids, od, ost = ox.search(queries, 10, L, mode=MODE_O[mode], order=O.ORDER_GPU, stats=True)
    for setting in ("1", "0"):
        with _env(BANG_B200_PREHASH=setting):
            s = api.BANGSearch("uint8", mode)
            assert s.bang_load(prefix)
            ids, d, st, info = _run(s, queries, 10, L)
        assert info.slot_block == int(setting)
        assert st["n_cand"].mean() > 10_000
        assert np.array_equal(ids, oids), (setting, int((ids != oids).any(1).sum()))
        assert np.array_equal(d.view(np.uint32), od.view(np.uint32))
        assert np.array_equal(st["n_cand"], ost["n_cand"]) and np.array_equal(st["hops"], ost["hops"])
