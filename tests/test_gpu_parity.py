"""Parity of the sm_100a path against the oracle, through the C ABI (run with -m gpu on a B200).

Bit-exact bar: ids identical and distances bit-identical to the oracle (ORDER_GPU) for every query,
all three modes, all three element types; PQ tables bit-identical (tolerance stated by north_star is
1e-5 relative — we hold 0).
"""
import os
import subprocess

import numpy as np
import pytest

from bang_b200 import api, build, formats, recall

import oracle as O

pytestmark = pytest.mark.gpu

MODE_O = {"base": O.MODE_BASE, "inmemory": O.MODE_INMEMORY, "exact": O.MODE_EXACT}


def _search(fx, mode, k, L, queries=None):
    s = api.BANGSearch(fx.dtype, mode)
    assert s.bang_load(fx.prefix), getattr(s, "last_error", "")
    s.set_dists_layout(api.DISTS_QUERY_MAJOR)
    s.bang_set_searchparams(k, L)
    q = fx.queries if queries is None else queries
    s.bang_alloc(len(q))
    s.bang_init(len(q))
    ids, dists = s.bang_query(q)
    stats = s.last_stats(len(q))
    timing = s.last_timing()
    s.bang_free()
    s.bang_unload()
    return ids, dists, stats, timing


@pytest.mark.parametrize("name", ["fx_u8", "fx_f32", "fx_i8"])
def test_pq_table_bit_exact(fixtures, name):
    fx = fixtures[name]
    s = api.BANGSearch(fx.dtype, "base")
    assert s.bang_load(fx.prefix)
    got = s.pq_table(fx.queries)
    ox = fx.oracle()
    for i in range(len(fx.queries)):
        want = ox.pq_table(fx.queries[i])
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32)), f"query {i}"
    # and within 1e-5 relative of a float64 evaluation of the reference formula (bang_search.cu:1118-1129)
    q = fx.queries[0].astype(np.float64) - fx.centroid.astype(np.float64)
    ref = np.stack([((fx.pivots[:, a:b].astype(np.float64) - q[a:b]) ** 2).sum(1)
                    for a, b in zip(fx.chunk_offsets[:-1], fx.chunk_offsets[1:])])
    assert np.allclose(got[0], ref, rtol=1e-5, atol=1e-6)
    s.bang_unload()


@pytest.mark.parametrize("name", ["fx_u8", "fx_f32", "fx_i8"])
@pytest.mark.parametrize("mode", ["base", "inmemory", "exact"])
@pytest.mark.parametrize("L", [10, 32, 100])
def test_search_bit_exact_vs_oracle(fixtures, name, mode, L):
    fx = fixtures[name]
    k = 10
    ids, dists, stats, timing = _search(fx, mode, k, L)
    ox = fx.oracle()
    oids, odists, ost = ox.search(fx.queries, k, L, mode=MODE_O[mode], order=O.ORDER_GPU, stats=True)
    assert timing.launches == 1 and timing.kernel_ms > 0
    assert np.array_equal(ids, oids), f"{(ids != oids).any(1).sum()} of {len(ids)} queries differ"
    assert np.array_equal(dists.view(np.uint32), odists.view(np.uint32))
    assert np.array_equal(stats["hops"], ost["hops"])
    assert np.array_equal(stats["sum_deg"], ost["sum_deg"])
    assert np.array_equal(stats["n_cand"], ost["n_cand"])


@pytest.mark.parametrize("fxname", ["fx_c1", "fx_c1m128"])
@pytest.mark.parametrize("mode", ["base", "inmemory", "exact"])
@pytest.mark.parametrize("L", [20, 64, 152])
def test_c1_shape_bit_exact_vs_oracle(request, fxname, mode, L):
    """C1 (BASELINE config 1: SIFT10K shape, N = 10^4, D = 128 u8, 100 queries; PQ m = 32 and the reference's m = 128,
    BANG_Inmemory/parANN.h:81-91): ids, distance bits and counters identical to the oracle, and recall vs brute force."""
    fx = request.getfixturevalue(fxname)
    ids, dists, stats, _ = _search(fx, mode, 10, L)
    oids, odists, ost = fx.oracle().search(fx.queries, 10, L, mode=MODE_O[mode], order=O.ORDER_GPU, stats=True)
    assert np.array_equal(ids, oids), f"{(ids != oids).any(1).sum()} of {len(ids)} queries differ"
    assert np.array_equal(dists.view(np.uint32), odists.view(np.uint32))
    assert np.array_equal(stats["hops"], ost["hops"]) and np.array_equal(stats["n_cand"], ost["n_cand"])
    if L >= 64:  # (m = 32 on this isotropic mixture: 88 % at L = 64, 98 % at L = 152; m = 128 and exact: 100 %)
        assert recall.calculate_recall(fx.gt_ids, fx.gt_dists, ids, 10) >= 85.0


@pytest.mark.parametrize("name,mode,L,floor", [("fx_u8", "inmemory", 64, 95.0), ("fx_f32", "base", 64, 95.0),
                                                 ("fx_u8", "exact", 32, 99.0)])
def test_recall_vs_bruteforce(fixtures, name, mode, L, floor):
    fx = fixtures[name]
    ids, _, _, _ = _search(fx, mode, 10, L)
    r = recall.calculate_recall(fx.gt_ids, fx.gt_dists, ids, 10)
    assert r >= floor, r


def test_dists_layout_rank_major_matches_reference_layout(fx_u8):
    fx = fx_u8
    s = api.BANGSearch(fx.dtype, "base")
    assert s.bang_load(fx.prefix)
    s.bang_set_searchparams(5, 20)
    Q = 16
    s.bang_alloc(Q)
    s.bang_init(Q)
    ids, d_rank = s.bang_query(fx.queries[:Q])  # default: rank-major, dists[j*Q+q] (bang_search.cu:999)
    s.set_dists_layout(api.DISTS_QUERY_MAJOR)
    s.bang_init(Q)
    ids2, d_q = s.bang_query(fx.queries[:Q])
    assert np.array_equal(ids, ids2)
    assert d_rank.shape == (5, Q) and d_q.shape == (Q, 5)
    assert np.array_equal(d_rank.T, d_q)
    assert (np.diff(d_q, axis=1) >= 0).all()  # ascending exact distance
    s.bang_free()
    s.bang_unload()


def test_edge_cases(fx_u8):
    fx = fx_u8
    ox = fx.oracle()
    # single query, k == L, k == 1
    for k, L in [(1, 1), (10, 10), (3, 512)]:
        ids, dists, _, _ = _search(fx, "base", k, L, fx.queries[:1])
        oids, od = ox.search(fx.queries[:1], k, L, mode=O.MODE_BASE)
        assert np.array_equal(ids, oids) and np.array_equal(dists, od)
    # repeated init/query on one allocation gives identical results (init re-arms the search, bang_search.cu:440-506)
    s = api.BANGSearch(fx.dtype, "inmemory")
    assert s.bang_load(fx.prefix)
    s.bang_set_searchparams(10, 40)
    s.bang_alloc(len(fx.queries))
    outs = []
    for _ in range(3):
        s.bang_init(len(fx.queries))
        outs.append(s.bang_query(fx.queries)[0].copy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    # a ragged tail: fewer queries than allocated
    s.bang_init(7)
    ids7, _ = s.bang_query(fx.queries[:7])
    assert np.array_equal(ids7, outs[0][:7])
    s.bang_free()
    s.bang_unload()


def test_error_behaviour(fx_u8, tmp_path):
    fx = fx_u8
    s = api.BANGSearch(fx.dtype, "base")
    assert not s.bang_load(str(tmp_path / "missing"))  # reference: bang_load returns false (bang_search.cu:152-177)
    with pytest.raises(api.BangError):
        s.bang_alloc(4)  # alloc before load
    assert s.bang_load(fx.prefix)
    with pytest.raises(api.BangError):
        s.bang_alloc(4)  # alloc before set_searchparams (bang_search.cu:370,379-384)
    with pytest.raises(api.BangError):
        s.bang_set_searchparams(10, 5)  # L < k (test_driver.cpp:394-398)
    with pytest.raises(api.BangError):
        s.bang_set_searchparams(10, 513)  # > MAX_L (bang_search.cu:439)
    s.bang_set_searchparams(10, 20)
    with pytest.raises(api.BangError):
        s.bang_query(fx.queries[:4])  # query before alloc
    s.bang_alloc(4)
    with pytest.raises(api.BangError):
        s.bang_init(5)  # more queries than allocated
    # wrong element type: entry length check
    s2 = api.BANGSearch("float", "base")
    assert not s2.bang_load(fx.prefix)
    # truncated graph file
    bad = str(tmp_path / "bad")
    for suf in ("_pq_pivots.bin", "_pq_compressed.bin", "_disk_metadata.bin"):
        with open(fx.prefix + suf, "rb") as f, open(bad + suf, "wb") as g:
            g.write(f.read())
    with open(fx.prefix + "_disk.bin", "rb") as f, open(bad + "_disk.bin", "wb") as g:
        g.write(f.read()[:-100])
    s3 = api.BANGSearch(fx.dtype, "base")
    assert not s3.bang_load(bad)


def test_load_files_old_layout_matches_prefix_load(fx_f32):
    fx = fx_f32
    ids_a, d_a, _, _ = _search(fx, "inmemory", 10, 48)
    s = api.BANGSearch(fx.dtype, "inmemory")
    p = fx.paths
    assert s.bang_load_files(p.old_pivots, p.pq_compressed, p.disk, p.old_chunk_offsets, p.old_centroid, fx.N, fx.D, fx.medoid)
    s.set_dists_layout(api.DISTS_QUERY_MAJOR)
    s.bang_set_searchparams(10, 48)
    s.bang_alloc(len(fx.queries))
    s.bang_init(len(fx.queries))
    ids_b, d_b = s.bang_query(fx.queries)
    assert np.array_equal(ids_a, ids_b) and np.array_equal(d_a, d_b)
    # exact mode needs no PQ files at all
    e = api.BANGSearch(fx.dtype, "exact")
    assert e.bang_load_files(None, None, p.disk, None, None, fx.N, fx.D, fx.medoid)


def test_mips_query_padding(fx_f32):
    """ENUM_DIST_MIPS: queries carry D-1 dims and are padded with one zero dim in-kernel (bang_search.cu:1099-1113)."""
    fx = fx_f32
    q_short = np.ascontiguousarray(fx.queries[:, :-1])
    q_pad = np.concatenate([q_short, np.zeros((len(q_short), 1), np.float32)], axis=1)
    s = api.BANGSearch(fx.dtype, "base")
    assert s.bang_load(fx.prefix)
    s.set_dists_layout(api.DISTS_QUERY_MAJOR)
    s.bang_set_searchparams(10, 32, api.ENUM_DIST_MIPS)
    s.bang_alloc(len(q_short))
    s.bang_init(len(q_short))
    ids, d = s.bang_query(q_short)
    oids, od = fx.oracle().search(q_pad, 10, 32, mode=O.MODE_BASE)
    assert np.array_equal(ids, oids) and np.array_equal(d, od)


def test_cli_driver_reports_reference_table(fx_u8):
    fx = fx_u8
    exe = build.CLI
    assert os.path.exists(exe)
    env = dict(os.environ, BANG_B200_MODE="base")
    out = subprocess.run([exe, fx.prefix, fx.paths.query, fx.paths.truth, str(len(fx.queries)), "10", "uint8", "l2", "40"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("40\t")]
    assert len(lines) == 5  # five timed runs per L (test_driver.cpp:424)
    ids, _, _, _ = _search(fx, "base", 10, 40)
    want = recall.calculate_recall(fx.gt_ids, fx.gt_dists, ids, 10)
    assert abs(float(lines[-1].split("\t")[3]) - want) < 0.01


@pytest.mark.parametrize("mode", ["base", "inmemory"])
def test_timers_breakdown_like_the_reference(fx_u8, mode):
    """BANG_B200_TIMERS=2: the driver prints the reference's `_TIMERS` table (bang_search.cu:1028-1051) from the phase
    clocks of libbang_b200_prof.so; the eight buckets add up to the fused kernel's device time and the answers do not change."""
    fx = fx_u8
    exe = build.build_cli()
    build.build_prof()
    out = {}
    for timers in ("2", None):
        env = dict(os.environ, BANG_B200_MODE=mode)
        env.pop("BANG_B200_TIMERS", None)
        if timers:
            env["BANG_B200_TIMERS"] = timers
        r = subprocess.run([exe, fx.prefix, fx.paths.query, fx.paths.truth, str(len(fx.queries)), "10", "uint8", "l2", "40"],
                           capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr
        out[timers] = r.stdout
    txt = out["2"]
    assert txt.count("STATS:") == 5 and "Total Search iterations" in txt
    import re
    blk = txt.split("STATS:")[-1]
    vals = {int(m.group(1)): float(m.group(2)) for m in re.finditer(r"^\((\d)\) [^=]*= ([0-9.]+) ms", blk, re.M)}
    assert sorted(vals) == [1, 2, 3, 4, 5, 6, 7, 8], blk
    total = float(re.search(r"Total time from timers[^=]*=[^=]*= ([0-9.]+) ms", blk).group(1))
    assert abs(sum(vals[i] for i in (1, 2, 3, 4, 5, 6, 8)) - total) < 0.02 * total + 0.01, blk
    assert vals[2] > 0 and vals[4] > 0 and vals[3] > 0, blk
    rows = lambda t: [l.split("\t")[3] for l in t.splitlines() if l.startswith("40\t")]
    assert rows(out["2"]) == rows(out[None])      # same recall column with and without the clocks


REF_DRIVER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_driver")


@pytest.mark.skipif(not os.path.exists(REF_DRIVER), reason="oracle/_ref not built (reference not mounted at build time)")
@pytest.mark.parametrize("name", ["fx_u8", "fx_f32"])
def test_against_reference_cuda_build(fixtures, name, tmp_path):
    """The UNMODIFIED reference (oracle/_ref/libbang.so) on this GPU vs the oracle vs the sm_100a path.
    north_star: identical top-k ids on >= 99% of queries, recall within 0.1 pt."""
    fx = fixtures[name]
    k, L = 10, 48
    out = str(tmp_path / "ref_ids.bin")
    dt = {"uint8": "uint8", "int8": "int8", "float": "float"}[fx.dtype]
    r = subprocess.run([REF_DRIVER, fx.prefix, fx.paths.query, str(len(fx.queries)), str(k), str(L), dt, out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ref_ids = np.fromfile(out, dtype=np.uint64).reshape(len(fx.queries), k)
    ids, _, _, _ = _search(fx, "base", k, L)
    same = (np.sort(ref_ids, 1) == np.sort(ids, 1)).all(1).mean()
    r_ref = recall.calculate_recall(fx.gt_ids, fx.gt_dists, ref_ids, k)
    r_new = recall.calculate_recall(fx.gt_ids, fx.gt_dists, ids, k)
    print(f"{name}: identical top-k sets {same:.4f}; recall ref {r_ref:.2f} new {r_new:.2f}")
    assert same >= 0.99
    assert abs(r_ref - r_new) <= 0.1 + 1e-9


@pytest.mark.parametrize("scheme", ["vmm", "ipc"])
def test_sharded_graph_p2p_two_gpus(scheme):
    """Graph rows sharded over 2 GPUs, peers read with P2P loads in the kernel: bit-exact vs the oracle, for both ways
    of sharing the rows between the processes (VMM + file descriptors, the default; cudaMalloc + CUDA IPC)."""
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs on one box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BANG_B200_SHARD_VMM="1" if scheme == "vmm" else "0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517" if scheme == "vmm" else "29518", os.path.join(root, "tests", "p2p_check.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ids == oracle: True") == 6


DROPIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "bang_search_dropin")


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref not built (reference not mounted at build time)")
def test_reference_driver_runs_on_this_library(fx_u8):
    """The reference's own test_driver.cpp, compiled unchanged against include/bang.h and linked with libbang_b200.so,
    sweeps L and prints its table; its recall column must equal ours."""
    fx = fx_u8
    out = subprocess.run([DROPIN, fx.prefix, fx.paths.query, fx.paths.truth, str(len(fx.queries)), "10", "uint8", "l2", "auto"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, BANG_B200_MODE="base"))
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [l.split("\t") for l in out.stdout.splitlines() if l[:1].isdigit() and l.count("\t") >= 3]
    got = {int(r[0]): float(r[3]) for r in rows}
    assert 10 in got and 22 in got and max(got) > 400        # L = 10, 22, 34, ... (step 12, test_driver.cpp:409-417)
    for L in (10, 34, 106):
        ids, _, _, _ = _search(fx, "base", 10, L)
        assert abs(got[L] - recall.calculate_recall(fx.gt_ids, fx.gt_dists, ids, 10)) < 0.01


@pytest.mark.parametrize("mode,L", [("inmemory", 512), ("base", 300), ("exact", 128)])
def test_visited_filter_spill_blocks(tmp_path, mode, L):
    """A random 64-regular graph hands every hop ~64 unseen ids: 2 x 64 x (L + 120) slots of the 399 887-slot filter
    get set (tens per 255-slot block), so most blocks of the sparse filter spill into their bitmaps.  The answers
    must still be those of the reference's plain array (hashFn1_d/hashFn2_d + neighbor_filtering_new,
    bang_search.cu:1140-1189), i.e. the oracle's."""
    import torch
    from bang_b200 import synth
    rng = np.random.default_rng(77)
    N, D, R, m = 60_000, 32, 64, 8
    base = rng.integers(0, 256, size=(N, D), dtype=np.uint8)
    r = rng.integers(0, N - 1, size=(N, R))
    while True:   # rows of distinct neighbours, no self loops (as any Vamana index has)
        srt = np.sort(r, axis=1)
        dup = np.nonzero((srt[:, 1:] == srt[:, :-1]).any(1))[0]
        if len(dup) == 0:
            break
        r[dup] = rng.integers(0, N - 1, size=(len(dup), R))
    nbrs = ((np.arange(N)[:, None] + 1 + r) % N).astype(np.uint32)
    deg = np.full(N, R, dtype=np.uint32)
    piv, cen, offs = synth.train_pq(torch.from_numpy(base), m, iters=4)
    codes = synth.encode_pq(torch.from_numpy(base), piv, cen, offs).numpy()
    prefix = str(tmp_path / "rnd")
    formats.write_index(prefix, base, deg, nbrs, 123, piv, cen, offs, codes)
    queries = rng.integers(0, 256, size=(24, D), dtype=np.uint8)
    s = api.BANGSearch("uint8", mode)
    assert s.bang_load(prefix)
    s.set_dists_layout(api.DISTS_QUERY_MAJOR)
    s.bang_set_searchparams(10, L)
    s.bang_alloc(len(queries))
    s.bang_init(len(queries))
    ids, dists = s.bang_query(queries)
    stats = s.last_stats(len(queries))
    s.bang_free()
    s.bang_unload()
    ox = O.OracleIndex(formats.pack_disk_bin(base, deg, nbrs), "uint8", D, R, 123, codes, piv, cen, offs)
    oids, od, ost = ox.search(queries, 10, L, mode=MODE_O[mode], order=O.ORDER_GPU, stats=True)
    assert stats["n_cand"].mean() > 15_000 or mode == "exact"   # the case really is a heavy one
    assert np.array_equal(ids, oids), f"{(ids != oids).any(1).sum()} of {len(ids)} queries differ"
    assert np.array_equal(dists.view(np.uint32), od.view(np.uint32))
    assert np.array_equal(stats["n_cand"], ost["n_cand"]) and np.array_equal(stats["hops"], ost["hops"])


def test_inmemory_cli_reports_recall(fx_u8):
    """`bang` with the Inmemory fork's 15-argument command line (parANN.cu:79-93) + medoid and L: same recall as the API."""
    fx = fx_u8
    exe = build.build_cli_inmem()
    p = fx.paths
    L, k = 48, 10
    ids, _, _, _ = _search(fx, "inmemory", k, L)
    want = recall.calculate_recall(fx.gt_ids, fx.gt_dists, ids, k)
    env = dict(os.environ, BANG_B200_DTYPE="uint8")
    env.pop("BANG_B200_MODE", None)
    out = subprocess.run([exe, p.old_pivots, p.pq_compressed, p.disk, p.query, p.old_chunk_offsets, p.old_centroid, p.truth,
                          str(len(fx.queries)), "1", "256", "512", "256", str(k), "64", "0", str(fx.medoid), str(L)],
                         capture_output=True, text=True, env=env, stdin=subprocess.DEVNULL, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.splitlines()
    assert lines[0] == f"{fx.medoid}\t{len(fx.queries)}"
    assert any(ln.startswith("Throughput = ") for ln in lines) and "Try Next run ? [y|n]" in lines
    row = lines[lines.index(f"Ls\tRecall@{k}") + 1].split("\t")
    assert int(row[0]) == L and abs(float(row[1]) - want) < 0.006
