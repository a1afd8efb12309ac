"""Host logic of the sharded (C5) index builder, on CPU tensors: the order-independent merge of the two copies of
a node's adjacency list (DiskANN build_merged_vamana_index: union of the shards' out-neighbours, truncated to R)
and the per-chunk data generator every rank replays."""
import numpy as np
import torch

from bang_b200 import build_sharded as BS


def _ref_merge(own, new, gid):
    """python restatement: union, drop self / empties / duplicates, keep the 64 smallest hash(gid, nbr)"""
    cand = sorted({int(x) for x in list(own) + list(new) if x >= 0 and x != gid})
    h = BS._hash32(torch.full((len(cand),), gid, dtype=torch.int64), torch.tensor(cand, dtype=torch.int64)).tolist()
    keep = [c for _, c in sorted(zip(h, cand))[:64]]
    return set(keep)


def test_merge_lists_union_truncate_and_order_independence():
    rng = np.random.default_rng(5)
    n_own, G, rank = 300, 4, 1
    a = torch.from_numpy(rng.integers(0, 5000, size=(n_own, 64)).astype(np.int32))
    b = torch.from_numpy(rng.integers(0, 5000, size=(n_own, 64)).astype(np.int32))
    a[::7, 40:] = -1          # short lists
    b[::5, 10:] = -1
    b[3, :] = a[3, :]         # identical copies
    b[4, 0] = 4 * G + rank    # a self loop must go
    rows = torch.arange(n_own, dtype=torch.int32)
    adj1 = torch.full((n_own, 64), -1, dtype=torch.int32)
    BS.merge_lists(adj1, rows, a, G, rank)          # first copies land in empty rows unchanged
    assert torch.equal(adj1, a)
    BS.merge_lists(adj1, rows, b, G, rank)
    adj2 = torch.full((n_own, 64), -1, dtype=torch.int32)
    perm = torch.randperm(n_own)
    BS.merge_lists(adj2, rows[perm], b[perm], G, rank)    # other arrival order, other row order
    BS.merge_lists(adj2, rows, a, G, rank)
    for r in range(n_own):
        want = _ref_merge(a[r].tolist(), b[r].tolist(), r * G + rank)
        got1 = {int(x) for x in adj1[r].tolist() if x >= 0}
        got2 = {int(x) for x in adj2[r].tolist() if x >= 0}
        assert got1 == want and got2 == want, r
        assert len(got1) == int((adj1[r] >= 0).sum())     # no duplicates survive
        assert len(got1) == min(64, len(want))


def test_merge_lists_partial_rows_only():
    adj = torch.full((10, 64), -1, dtype=torch.int32)
    new = torch.arange(64, dtype=torch.int32)[None, :].repeat(2, 1) + 100
    BS.merge_lists(adj, torch.tensor([2, 7], dtype=torch.int32), new, 2, 0)
    assert int((adj >= 0).any(1).sum()) == 2 and torch.equal(adj[2], new[0]) and torch.equal(adj[7], new[1])


def test_gen_chunk_is_replayable_and_chunk_dependent():
    centers = BS.mixture_centers(16, 32, "cpu")
    x0 = BS.gen_chunk(centers, 0, 1000, 7)
    assert x0.dtype == torch.uint8 and x0.shape == (1000, 32)
    assert torch.equal(x0, BS.gen_chunk(centers, 0, 1000, 7))
    assert not torch.equal(x0, BS.gen_chunk(centers, 1, 1000, 7))
    assert torch.equal(centers, BS.mixture_centers(16, 32, "cpu"))


def test_hash32_range():
    a = torch.arange(0, 10**9, 10**6, dtype=torch.int64)
    h = BS._hash32(a, a.flip(0))
    assert int(h.min()) >= 0 and int(h.max()) < 2**31 and h.unique().numel() > 990


def test_partition_ownership_ids():
    """assign_ids: the i-th point owned by GPU g gets id i*G + g — so `id mod G` is the owner and `id div G` the local
    row, the rule the search library already uses — consistently across chunks."""
    G = 4
    rng = np.random.default_rng(9)
    counts = [0] * G
    seen = {}
    for chunk in range(5):
        owner = torch.from_numpy(rng.integers(0, G, size=1000))
        ids = BS.assign_ids(owner, counts, G)
        assert torch.equal(ids % G, owner)
        for g in range(G):
            rows = (ids[owner == g] // G).tolist()
            start = seen.get(g, 0)
            assert rows == list(range(start, start + len(rows)))       # dense, in generation order
            seen[g] = start + len(rows)
    assert counts == [seen[g] for g in range(G)] and sum(counts) == 5000


def test_balance_partitions():
    sizes = [50, 10, 10, 10, 30, 30, 5, 5, 40, 20, 20, 10]
    g = BS.balance_partitions(sizes, 4)
    loads = [sum(s for s, o in zip(sizes, g.tolist()) if o == k) for k in range(4)]
    assert sorted(set(g.tolist())) == [0, 1, 2, 3] and max(loads) - min(loads) <= 10 and sum(loads) == sum(sizes)
    assert torch.equal(g, BS.balance_partitions(sizes, 4))             # deterministic: every rank computes the same map
