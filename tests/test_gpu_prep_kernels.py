"""The data-preparation kernels (csrc/prep_kernels.cu: exact kNN ground truth, PQ k-means, PQ encoding) against the CPU
oracle and plain numpy/torch references.  These replace DiskANN's `compute_groundtruth` and the PQ stage of
`build_disk_index` (README.md:46-58) for indices built on the box; only the file formats are contractual there."""
import numpy as np
import pytest

from bang_b200 import api, formats, recall, synth

import oracle as O

pytestmark = [pytest.mark.gpu]


def _dev():
    import torch
    return torch.device("cuda", 0)


@pytest.mark.parametrize("name", ["fx_u8", "fx_f32", "fx_i8", "fx_c1"])
def test_ground_truth_matches_oracle_bruteforce(request, name):
    """ids ordered by (distance, id) identical to the oracle's brute force; integer distances exact, float within 1e-5."""
    import torch
    fx = request.getfixturevalue(name)
    base = torch.from_numpy(fx.base).to(_dev())
    q = torch.from_numpy(fx.queries).to(_dev())
    k = 32
    ids, d = api.bruteforce_gt(base, q, k)
    ids, d = ids.cpu().numpy().astype(np.uint32), d.cpu().numpy()
    oids, od = fx.oracle().bruteforce(fx.queries, k)
    if fx.dtype == "float":
        assert np.allclose(d, od, rtol=1e-5)
        assert (ids == oids).mean() > 0.999          # near-ties may swap under a different summation order
    else:
        assert np.array_equal(d, od)
        assert np.array_equal(ids, oids)
    assert (np.diff(d, axis=1) >= 0).all()
    # and it is what the committed truthset holds
    assert np.array_equal(ids[:, :10], fx.gt_ids[:, :10]) or fx.dtype == "float"


def test_ground_truth_ties_offsets_and_growth():
    """Many exact ties (few distinct points), an id offset, more passes than one (n > 3840), k > n padding."""
    import torch
    rng = np.random.default_rng(5)
    protos = rng.integers(0, 255, size=(7, 20), dtype=np.uint8)
    base = protos[rng.integers(0, 7, size=60_000)]
    q = protos[:5].copy()
    ids, d = api.bruteforce_gt(torch.from_numpy(base).to(_dev()), torch.from_numpy(q).to(_dev()), 100, id_offset=1000)
    ids, d = ids.cpu().numpy(), d.cpu().numpy()
    for i in range(5):
        same = np.nonzero((base == q[i]).all(1))[0][:100] + 1000
        assert (d[i] == 0).all() and np.array_equal(ids[i], same)      # the 100 smallest ids among thousands of ties
    small = torch.from_numpy(base[:40]).to(_dev())
    ids, d = api.bruteforce_gt(small, torch.from_numpy(q).to(_dev()), 64)
    assert (ids.cpu().numpy()[:, 40:] == 0xFFFFFFFF).all() and (ids.cpu().numpy()[:, :40] < 40).all()


@pytest.mark.parametrize("name", ["fx_u8", "fx_f32", "fx_i8"])
def test_pq_encode_is_the_closest_pivot(fixtures, name):
    import torch
    fx = fixtures[name]
    codes = api.pq_encode(torch.from_numpy(fx.base).to(_dev()), fx.pivots, fx.centroid, fx.chunk_offsets).cpu().numpy()
    x = fx.base.astype(np.float32) - fx.centroid[None, :]
    for c in range(fx.m):
        a, b = int(fx.chunk_offsets[c]), int(fx.chunk_offsets[c + 1])
        d2 = ((x[:, None, a:b] - fx.pivots[None, :, a:b]) ** 2).sum(2)
        best = d2.min(1)
        got = d2[np.arange(len(x)), codes[:, c]]
        assert np.all(got <= best * (1 + 1e-5) + 1e-6)      # the chosen centre is a minimiser (ties / rounding aside)
    assert (codes == fx.codes).mean() > 0.999               # and agrees with the codes of the committed fixture


def test_pq_train_quantisation_error_and_determinism():
    """k-means on the GPU: same result on every run (integer accumulation), distortion no worse than the torch reference."""
    import torch
    base, _ = synth.make_clustered(30_000, 64, "uint8", seed=11)
    offs = synth.chunk_offsets_even(64, 16)
    dev_base = base.to(_dev())
    p1, c1 = api.pq_train(dev_base, offs, iters=12, max_train=20_000, seed=3)
    p2, c2 = api.pq_train(dev_base, offs, iters=12, max_train=20_000, seed=3)
    assert np.array_equal(p1, p2) and np.array_equal(c1, c2)
    assert np.allclose(c1, base.float().mean(0).numpy(), atol=1e-3)

    def distortion(piv, cen):
        codes = api.pq_encode(dev_base, piv, cen, offs).cpu().numpy().astype(np.int64)
        x = base.float().numpy() - cen[None, :]
        rec = np.concatenate([piv[codes[:, c], int(offs[c]):int(offs[c + 1])] for c in range(16)], 1)
        return float(((x - rec) ** 2).sum(1).mean())

    import os
    os.environ["BANG_B200_TORCH_PREP"] = "1"
    try:
        pt, ct, _ = synth.train_pq(dev_base, 16, iters=12, max_train=20_000)
    finally:
        del os.environ["BANG_B200_TORCH_PREP"]
    assert distortion(p1, c1) <= distortion(pt, ct) * 1.05
