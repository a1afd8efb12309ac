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
