"""GPU checks of the device-resident entry points against the file/host path on the same index:
  * bang_b200_load_device_begin/_rows/_codes/_end (indices handed over in device memory) == bang_load from files,
  * bang_b200_query_device (device queries/results, caller's stream, no sync) == bang_query,
  * the GPU Vamana builder (bang_b200_build_vamana) produces a graph the search reaches high recall on.
These are the paths bench.py times (query_device) and the sharded C5 driver loads through (load_device_*)."""
import os

import numpy as np
import pytest

from bang_b200 import api, recall

pytestmark = [pytest.mark.gpu]


def _host_search(fx, mode, k, L):
    s = api.BANGSearch(fx.dtype, mode)
    assert s.bang_load(fx.prefix)
    s.set_dists_layout(api.DISTS_QUERY_MAJOR)
    s.bang_set_searchparams(k, L)
    s.bang_alloc(len(fx.queries))
    s.bang_init(len(fx.queries))
    ids, d = s.bang_query(fx.queries)
    ids, d = ids.copy(), d.copy()
    return s, ids, d


@pytest.mark.parametrize("name", ["fx_u8", "fx_f32", "fx_i8"])
@pytest.mark.parametrize("mode", ["inmemory", "exact"])
def test_device_resident_load_equals_file_load(fixtures, name, mode):
    import torch
    fx = fixtures[name]
    k, L = 10, 40
    s1, ids1, d1 = _host_search(fx, mode, k, L)
    s1.bang_free(); s1.bang_unload()
    dev = torch.device("cuda", 0)
    vec = torch.from_numpy(np.ascontiguousarray(fx.base)).to(dev)
    adj_np = fx.nbrs.astype(np.uint32).copy()
    adj_np[np.arange(fx.R)[None, :] >= fx.deg[:, None]] = 0xFFFFFFFF           # unused slots
    adj = torch.from_numpy(adj_np.view(np.int32)).to(dev)
    codes = torch.from_numpy(np.ascontiguousarray(fx.codes)).to(dev)
    s2 = api.BANGSearch(fx.dtype, mode)
    if mode == "exact":
        s2.load_device_begin(fx.N, fx.D, fx.medoid)
    else:
        s2.load_device_begin(fx.N, fx.D, fx.medoid, fx.pivots, fx.centroid, fx.chunk_offsets)
    half = fx.N // 2                                                            # two slices, as a chunked producer would
    s2.load_device_rows(0, half, vec[:half].data_ptr(), adj[:half].data_ptr())
    s2.load_device_rows(half, fx.N - half, vec[half:].data_ptr(), adj[half:].data_ptr())
    if mode != "exact":
        s2.load_device_codes(0, half, codes[:half].data_ptr())
        s2.load_device_codes(half, fx.N - half, codes[half:].data_ptr())
    s2.load_device_end()
    torch.cuda.synchronize()
    s2.set_dists_layout(api.DISTS_QUERY_MAJOR)
    s2.bang_set_searchparams(k, L)
    s2.bang_alloc(len(fx.queries))
    s2.bang_init(len(fx.queries))
    ids2, d2 = s2.bang_query(fx.queries)
    assert np.array_equal(ids1, ids2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))
    s2.bang_free(); s2.bang_unload()


def test_query_device_equals_query(fx_u8):
    import torch
    fx = fx_u8
    k, L = 10, 40
    s, ids1, d1 = _host_search(fx, "inmemory", k, L)
    dev = torch.device("cuda", 0)
    Q = len(fx.queries)
    d_q = torch.from_numpy(np.ascontiguousarray(fx.queries)).to(dev)
    d_ids = torch.zeros((Q, k), dtype=torch.int64, device=dev)
    d_d = torch.zeros((Q, k), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    s.bang_init(Q)
    torch.cuda.synchronize(dev)
    s.query_device(d_q.data_ptr(), Q, d_ids.data_ptr(), d_d.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    assert np.array_equal(d_ids.cpu().numpy().astype(np.uint64), ids1)
    assert np.array_equal(d_d.cpu().numpy().view(np.uint32), d1.view(np.uint32))
    s.bang_free(); s.bang_unload()


def test_gpu_builder_graph_is_searchable(tmp_path):
    import torch
    from bang_b200 import builder, formats
    prefix = str(tmp_path / "g")
    info = builder.make_fixture_auto(prefix, 50_000, 64, "uint8", 200, 16, k_gt=10, device=torch.device("cuda", 0), builder="gpu")
    assert info["builder"] == "gpu" and 20 <= info["mean_degree"] <= 64
    meta = formats.read_disk_metadata(prefix + "_disk_metadata.bin")
    _, deg, nbrs = formats.read_disk_bin(prefix + "_disk.bin", meta)
    assert int(deg.min()) >= 1 and int(nbrs[np.arange(64)[None, :] < deg[:, None]].max()) < 50_000
    for i in range(0, 50_000, 997):                                            # rows hold distinct ids, no self loops
        row = nbrs[i, :deg[i]]
        assert len(set(row.tolist())) == len(row) and i not in row
    q = formats.read_bin(prefix + "_query.bin", np.uint8)
    gi, gd = formats.read_truthset(prefix + "_gt.bin")
    s = api.BANGSearch("uint8", "exact")
    assert s.bang_load(prefix)
    s.bang_set_searchparams(10, 64)
    s.bang_alloc(len(q)); s.bang_init(len(q))
    ids, _ = s.bang_query(q)
    assert recall.calculate_recall(gi, gd, ids, 10) >= 97.0
    s.bang_free(); s.bang_unload()


@pytest.mark.parametrize("mode", ["base", "inmemory", "exact"])
def test_degenerate_graphs_match_oracle(tmp_path, mode):
    """Isolated entry point, fewer points than k, a dead end: same ids, distances and filler as the oracle."""
    import oracle as O
    import pathological as P
    from bang_b200 import formats
    omode = {"base": O.MODE_BASE, "inmemory": O.MODE_INMEMORY, "exact": O.MODE_EXACT}[mode]
    for name, (base, deg, nbrs, medoid, piv, cen, offs, codes, k, L) in P.cases().items():
        prefix = str(tmp_path / name)
        formats.write_index(prefix, base, deg, nbrs, medoid, piv, cen, offs, codes)
        s = api.BANGSearch("uint8", mode)
        assert s.bang_load(prefix)
        s.set_dists_layout(api.DISTS_QUERY_MAJOR)
        s.bang_set_searchparams(k, L)
        s.bang_alloc(1); s.bang_init(1)
        ids, d = s.bang_query(P.QUERY)
        st = s.last_stats(1)
        s.bang_free(); s.bang_unload()
        ox = O.OracleIndex(formats.pack_disk_bin(base, deg, nbrs), "uint8", 4, 64, medoid, codes, piv, cen, offs)
        oids, od, ost = ox.search(P.QUERY, k, L, mode=omode, stats=True)
        assert np.array_equal(ids, oids) and np.array_equal(d.view(np.uint32), od.view(np.uint32)), name
        assert np.array_equal(st["hops"], ost["hops"]) and np.array_equal(st["n_cand"], ost["n_cand"]), name


# ---- load-time validation (reference: only the first and last adjacency entry are asserted, bang_search.cu:335-345) ----
def _write_variant(fx, tmp_path, name, deg=None, nbrs=None, dtype_code=None):
    import struct
    from bang_b200 import formats
    prefix = str(tmp_path / name)
    formats.write_index(prefix, fx.base, fx.deg if deg is None else deg, fx.nbrs if nbrs is None else nbrs, fx.medoid,
                        fx.pivots, fx.centroid, fx.chunk_offsets, fx.codes)
    if dtype_code is not None:
        p = formats.IndexPaths(prefix).disk_meta
        raw = bytearray(open(p, "rb").read())
        raw[16:20] = struct.pack("<i", dtype_code)
        open(p, "wb").write(bytes(raw))
    return prefix


@pytest.mark.parametrize("mode", ["base", "exact"])
def test_ill_formed_rows_are_refused_at_load(fx_u8, tmp_path, mode):
    """A neighbour id >= N, a repeated id within one row, or a degree above R fail bang_load with BANG_E_FORMAT
    (corrupt _disk.bin: no out-of-bounds device reads, no silently different traversal)."""
    fx = fx_u8
    nb = fx.nbrs.copy()
    nb[17, 0] = fx.N + 5                                   # id out of range
    s = api.BANGSearch(fx.dtype, mode)
    assert not s.bang_load(_write_variant(fx, tmp_path, "oob", nbrs=nb))
    assert "neighbour id >= N" in s.last_error and " 1 with a neighbour" in s.last_error
    nb = fx.nbrs.copy()
    row = int(np.argmax(fx.deg >= 2))
    nb[row, 1] = nb[row, 0]                                # the same id twice
    assert not s.bang_load(_write_variant(fx, tmp_path, "dup", nbrs=nb))
    assert " 1 with a repeated" in s.last_error
    dg = fx.deg.copy()
    dg[3] = 65                                             # degree above R = 64
    assert not s.bang_load(_write_variant(fx, tmp_path, "deg", deg=dg))
    assert " 1 with degree > R" in s.last_error
    assert s.bang_load(fx.prefix)                          # the handle is reusable after a refused load
    s.bang_unload()


def test_element_type_mismatch_is_refused(fx_u8, fx_i8, tmp_path):
    """int8 and uint8 entries have the same length; the metadata's datatype word (bang_preprocess.py:42-51) tells them apart."""
    s = api.BANGSearch("uint8", "base")
    assert not s.bang_load(fx_i8.prefix) and "element type" in s.last_error
    s = api.BANGSearch("int8", "inmemory")
    assert not s.bang_load(fx_u8.prefix) and "element type" in s.last_error
    # a datatype word outside the converter's numbering is not trusted either way (the reference only prints it)
    s = api.BANGSearch("uint8", "base")
    assert s.bang_load(_write_variant(fx_u8, tmp_path, "odd", dtype_code=7))
    s.bang_unload()


def test_pivot_table_larger_than_shared_memory(tmp_path):
    """D = 320 floats: the 256 x D pivot table (320 KB) cannot live in shared memory, so the PQ modes read it from
    global memory (GIST1M-class indices in Base / Inmemory mode, which the reference runs).  Same bar: bit-exact."""
    import torch
    from bang_b200 import builder, formats
    import oracle as O
    prefix = str(tmp_path / "wide")
    nq = 32
    builder.make_fixture_auto(prefix, 6000, 320, "float", nq, 40, k_gt=10, device=torch.device("cuda", 0))
    queries = formats.read_bin(prefix + "_query.bin", np.float32)
    ox = O.OracleIndex.from_files(prefix)
    for mode, om in (("inmemory", O.MODE_INMEMORY), ("base", O.MODE_BASE)):
        s = api.BANGSearch("float", mode)
        assert s.bang_load(prefix), s.last_error
        s.set_dists_layout(api.DISTS_QUERY_MAJOR)
        s.bang_set_searchparams(10, 48)
        s.bang_alloc(nq); s.bang_init(nq)
        ids, d = s.bang_query(queries)
        st = s.last_stats(nq)
        s.bang_free(); s.bang_unload()
        oids, od, ost = ox.search(queries, 10, 48, mode=om, order=O.ORDER_GPU, stats=True)
        assert np.array_equal(ids, oids) and np.array_equal(d.view(np.uint32), od.view(np.uint32))
        assert np.array_equal(st["n_cand"], ost["n_cand"]) and np.array_equal(st["hops"], ost["hops"])


def test_query_device_on_two_streams_is_serialised(fx_u8):
    """Two bang_b200_query_device calls on different caller streams share the handle's filters and work counter;
    the library orders the second launch after the first, so both return the single-stream answer."""
    import torch
    fx = fx_u8
    s, ids_ref, _ = _host_search(fx, "inmemory", 10, 32)
    dev = torch.device("cuda", 0)
    Q = len(fx.queries)
    d_q = torch.from_numpy(fx.queries).to(dev)
    outs = [(torch.zeros((Q, 10), dtype=torch.int64, device=dev), torch.zeros((Q, 10), dtype=torch.float32, device=dev)) for _ in range(4)]
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    torch.cuda.synchronize(dev)
    for i, (oi, od) in enumerate(outs):
        s.query_device(d_q.data_ptr(), Q, oi.data_ptr(), od.data_ptr(), streams[i % 2].cuda_stream)
    torch.cuda.synchronize(dev)
    for oi, _ in outs:
        assert np.array_equal(oi.cpu().numpy().astype(np.uint64), ids_ref)
    s.bang_free(); s.bang_unload()


def test_failed_alloc_releases_everything(fx_u8):
    """bang_alloc that runs out of memory leaves the handle as it was (advisor finding: partial allocations leaked)."""
    import torch
    fx = fx_u8
    s = api.BANGSearch(fx.dtype, "inmemory")
    assert s.bang_load(fx.prefix)
    s.bang_set_searchparams(10, 32)
    free0, _ = torch.cuda.mem_get_info(0)
    with pytest.raises(api.BangError):
        s.bang_alloc(2_000_000_000)  # 2e9 queries x 32 B: cannot fit
    free1, _ = torch.cuda.mem_get_info(0)
    assert abs(free0 - free1) < (64 << 20)
    s.bang_alloc(len(fx.queries)); s.bang_init(len(fx.queries))   # and a sane request still works
    ids, _ = s.bang_query(fx.queries)
    assert (ids[:, 0] != api.NO_ID).all()
    s.bang_free(); s.bang_unload()
