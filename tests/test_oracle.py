"""The oracle (oracle/bang_oracle.c) against: known answers computed independently in Python, the committed
outputs of the UNMODIFIED reference CUDA build on a B200 (tests/golden/ref_golden.npz), and brute force."""
import os

import numpy as np
import pytest

from bang_b200 import formats, recall

import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BF = 399887


def _py_hash(x, seed, mult):
    h = seed
    for s in (0, 8, 16, 24):
        h = ((h ^ ((x >> s) & 0xFF)) * mult) & 0xFFFFFFFFFFFFFFFF
    return h % BF


@pytest.mark.parametrize("x", [0, 1, 255, 256, 65535, 12345, 178757270, 999999999, 0xFFFFFFFF, 0x01010101])
def test_hash_known_answers(x):
    # hashFn1_d / hashFn2_d, bang_search.cu:1168-1189: uint64 xor-multiply rounds, LSB first, mod 399887
    assert O.hash1(x) == _py_hash(x, 0xCBF29CE4, 0x01000193)
    assert O.hash2(x) == _py_hash(x, 0x84222325, 0x1B3)
    assert O.hash1(x) < BF and O.hash2(x) < BF


def test_hash_fixed_values():
    # frozen values: a change in the hash changes which candidates the bloom filter drops
    assert [O.hash1(v) for v in (0, 1, 2, 1000000)] == [_py_hash(v, 0xCBF29CE4, 0x01000193) for v in (0, 1, 2, 1000000)]
    assert O.hash1(12345) == 0xB757 and O.hash2(12345) == 275367


@pytest.mark.parametrize("name", ["fx_u8", "fx_f32", "fx_i8"])
def test_pq_table_matches_float64_formula(fixtures, name):
    fx = fixtures[name]
    ox = fx.oracle()
    for qi in (0, 3, len(fx.queries) - 1):
        got = ox.pq_table(fx.queries[qi])
        q = fx.queries[qi].astype(np.float64) - fx.centroid.astype(np.float64)
        ref = np.stack([((fx.pivots[:, a:b].astype(np.float64) - q[a:b]) ** 2).sum(1)
                        for a, b in zip(fx.chunk_offsets[:-1], fx.chunk_offsets[1:])])
        assert got.shape == (fx.m, 256)
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)


def test_pq_distance_is_sum_of_table_entries(fx_u8):
    fx = fx_u8
    ox = fx.oracle()
    tbl = ox.pq_table(fx.queries[0])
    for node in (0, 7, fx.medoid, fx.N - 1):
        want = sum(float(tbl[c, fx.codes[node, c]]) for c in range(fx.m))
        assert abs(ox.pq_dist(tbl, node) - want) <= 1e-5 * max(1.0, want)


@pytest.mark.parametrize("name", ["fx_u8", "fx_i8"])
def test_integer_l2_is_exact_in_every_order(fixtures, name):
    fx = fixtures[name]
    ox = fx.oracle()
    q = fx.queries[1]
    for node in (0, 5, fx.N - 1):
        want = float(((fx.base[node].astype(np.int64) - q.astype(np.int64)) ** 2).sum())
        for order in (O.ORDER_REF, O.ORDER_GPU):
            for kind in (0, 1):
                assert ox.l2(node, q, order, kind) == want


def test_float_l2_orders_agree_to_rounding(fx_f32):
    fx = fx_f32
    ox = fx.oracle()
    q = fx.queries[2]
    for node in range(0, 200, 13):
        want = float(((fx.base[node].astype(np.float64) - q.astype(np.float64)) ** 2).sum())
        for order in (O.ORDER_REF, O.ORDER_GPU):
            for kind in (0, 1):
                assert abs(ox.l2(node, q, order, kind) - want) <= 1e-5 * want + 1e-7


@pytest.mark.parametrize("name", ["fx_u8", "fx_f32", "fx_i8"])
@pytest.mark.parametrize("L", [10, 32, 100])
def test_oracle_reproduces_reference_cuda_build(name, L, fixtures):
    """tests/golden/ref_golden.npz = ids returned by the unmodified BANG_Base built from /root/reference and
    run on a B200 (tests/golden/make_ref_golden.py).  Bar (north_star): identical top-k ids on >= 99% of queries."""
    g = np.load(os.path.join(GOLDEN, "ref_golden.npz"))
    fx = fixtures[name]
    ox = fx.oracle()
    for order in (O.ORDER_REF, O.ORDER_GPU):
        ids, _ = ox.search(fx.queries, 10, L, mode=O.MODE_BASE, order=order)
        for rep in range(3):
            ref = g[f"ids_{name}_L{L}_rep{rep}"]
            same = (ref == ids).all(1).mean()
            assert same >= 0.99, (name, L, order, rep, same)


FORK_CASES = [("c1", "fx_c1", 20), ("c1", "fx_c1", 64), ("c1", "fx_c1", 152), ("c1m128", "fx_c1m128", 64),
              ("f32", "fx_f32", 32), ("i8", "fx_i8", 32)]


@pytest.mark.parametrize("fork,mode", [("inmem", O.MODE_INMEMORY), ("exact", O.MODE_EXACT)])
@pytest.mark.parametrize("case,fxname,L", FORK_CASES)
def test_oracle_reproduces_reference_forks(fork, mode, case, fxname, L, request):
    """tests/golden/ref_forks_golden.npz = the ids the reference's OWN Inmemory / Exactdistance programs returned on a
    B200 (BANG_Inmemory/parANN.cu, BANG_Exactdistance/parANN.cu built by oracle/build_ref_forks.sh for each fixture and L;
    tests/golden/make_ref_forks_golden.py), including the C1 shape (N = 10^4, D = 128 u8, 100 queries; m = 32 and the
    reference's m = 128).  Bar (north_star): identical top-k ids on >= 99 % of the queries, every repetition."""
    g = np.load(os.path.join(GOLDEN, "ref_forks_golden.npz"))
    fx = request.getfixturevalue(fxname)
    ox = fx.oracle()
    for order in (O.ORDER_REF, O.ORDER_GPU):
        ids, _ = ox.search(fx.queries, 10, L, mode=mode, order=order)
        for rep in range(3):
            ref = g[f"ids_{fork}_{case}_L{L}_rep{rep}"].astype(np.uint64)
            same = (ref == ids).all(1).mean()
            assert same >= 0.99, (fork, case, L, order, rep, same)


@pytest.mark.parametrize("mode", [O.MODE_BASE, O.MODE_INMEMORY, O.MODE_EXACT])
def test_recall_against_bruteforce(fx_u8, mode):
    fx = fx_u8
    ox = fx.oracle()
    prev = 0.0
    for L in (10, 32, 100):
        ids, d = ox.search(fx.queries, 10, L, mode=mode)
        r = recall.calculate_recall(fx.gt_ids, fx.gt_dists, ids, 10)
        assert r >= prev - 1.0  # recall grows with the worklist length
        prev = r
        assert (np.diff(d, axis=1) >= 0).all()
    assert prev >= 99.0


def test_bruteforce_matches_ground_truth(fx_f32):
    fx = fx_f32
    ids, d = fx.oracle().bruteforce(fx.queries[:16], 10)
    assert np.array_equal(ids, fx.gt_ids[:16, :10])
    assert np.allclose(d, fx.gt_dists[:16, :10], rtol=1e-5)


def test_base_and_inmemory_agree_except_ties(fx_u8):
    # the two forks differ only in where the parent is chosen (prefetch vs after the merge, SURVEY A.2 / A.2')
    fx = fx_u8
    ox = fx.oracle()
    a, _ = ox.search(fx.queries, 10, 64, mode=O.MODE_BASE)
    b, _ = ox.search(fx.queries, 10, 64, mode=O.MODE_INMEMORY)
    assert (a == b).all(1).mean() >= 0.95


def test_stats_and_trace(fx_u8):
    fx = fx_u8
    ox = fx.oracle()
    L = 24
    ids, d, st = ox.search(fx.queries[:8], 10, L, mode=O.MODE_BASE, stats=True, trace_len=L + 51)
    assert (st["hops"] >= 10).all() and (st["hops"] <= L + 50).all()  # at most L+50 candidates (bang_search.cu:603,950)
    tr = st["trace"]
    assert (tr[:, 0] == fx.medoid).all()  # the medoid is every query's first candidate (bang_search.cu:455-462)
    for q in range(8):
        t = tr[q, :st["hops"][q]]
        assert len(set(t.tolist())) == len(t)  # a node is expanded at most once
        assert set(ids[q].tolist()) <= set(t.tolist())  # results come from the candidate log (re-rank)
        assert st["sum_deg"][q] <= fx.deg[t].sum()
        assert st["n_cand"][q] <= st["sum_deg"][q] + 1


def test_edge_cases(fx_u8):
    fx = fx_u8
    ox = fx.oracle()
    # k == L == 1; single query; L at the reference's maximum
    ids, d = ox.search(fx.queries[:1], 1, 1, mode=O.MODE_BASE)
    assert ids.shape == (1, 1) and ids[0, 0] < fx.N
    ids, d = ox.search(fx.queries[:2], 10, 512, mode=O.MODE_INMEMORY)
    assert (ids < fx.N).all()
    # exact mode with k > reachable worklist: filler id / FLT_MAX
    ids, d = ox.search(fx.queries[:1], 4, 4, mode=O.MODE_EXACT)
    assert (ids < fx.N).all()


def test_sequential_filter_semantics():
    """A.3: ids are tested and inserted in list order; an id whose two slots are already set is dropped."""
    # build a 3-node toy graph where node 0 lists the same neighbour ids in ascending order
    D, R = 4, 64
    base = np.zeros((70, D), dtype=np.uint8)
    base[:, 0] = np.arange(70)
    deg = np.zeros(70, dtype=np.uint32)
    nbrs = np.zeros((70, R), dtype=np.uint32)
    deg[0] = 64
    nbrs[0] = np.arange(1, 65)
    for i in range(1, 70):
        deg[i] = 1
        nbrs[i, 0] = 0
    disk = formats.pack_disk_bin(base, deg, nbrs)
    ox = O.OracleIndex(disk, "uint8", D, R, 0)
    q = np.array([[3, 0, 0, 0]], dtype=np.uint8)
    ids, d, st = ox.search(q, 5, 8, mode=O.MODE_EXACT, stats=True)
    assert ids[0].tolist() == [3, 2, 4, 1, 5] and d[0].tolist() == [0.0, 1.0, 1.0, 4.0, 4.0]  # ties by id
    assert st["n_cand"][0] == 65  # medoid + 64 neighbours, every later list is fully filtered


@pytest.mark.parametrize("mode", [O.MODE_BASE, O.MODE_INMEMORY, O.MODE_EXACT])
def test_degenerate_graphs(mode):
    """Searches that run out of graph return what they reached, in (exact distance, id) order, and fill the remaining
    ranks with id 0xFFFFFFFF / FLT_MAX (the reference reads uninitialised slots there)."""
    import pathological as P
    want = {"isolated_entry": [0], "fewer_points_than_k": [2, 4, 3, 1, 0], "dead_end": [2, 1, 0]}
    for name, (base, deg, nbrs, medoid, piv, cen, offs, codes, k, L) in P.cases().items():
        ox = O.OracleIndex(formats.pack_disk_bin(base, deg, nbrs), "uint8", 4, 64, medoid, codes, piv, cen, offs)
        ids, d, st = ox.search(P.QUERY, k, L, mode=mode, stats=True)
        n = len(want[name])
        assert ids[0, :n].tolist() == want[name], name
        assert (ids[0, n:] == P.NO_ID).all() and (d[0, n:] == np.float32(3.4028234663852886e+38)).all(), name
        exact = ((base[want[name]].astype(np.float32) - P.QUERY[0].astype(np.float32)) ** 2).sum(1)
        assert d[0, :n].tolist() == exact.tolist() and np.all(np.diff(d[0, :n]) >= 0), name
        assert st["hops"][0] == n or mode == O.MODE_EXACT, name     # every reached node was expanded once
