import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import bang_b200  # noqa: E402,F401  (root shim: exposes bang-billion-scale-ann_b200/ as `bang_b200`)
from bang_b200 import formats  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


class Fixture:
    """A committed golden index materialised into the reference's file formats."""

    def __init__(self, name: str, tmpdir: str, pq: str = "", npz: str | None = None):
        """pq: suffix of the PQ arrays inside the npz (fx_c1 carries a second PQ layout, m = 128, as `*128`)."""
        z = np.load(os.path.join(GOLDEN, (npz or name) + ".npz"))
        self.name = name
        self.base = z["base"]
        self.deg = z["deg"]
        self.nbrs = z["nbrs"]
        self.medoid = int(z["medoid"])
        self.pivots = z["pivots" + pq]
        self.centroid = z["centroid" + pq]
        self.chunk_offsets = z["chunk_offsets" + pq]
        self.codes = z["codes" + pq]
        self.queries = z["queries"]
        self.gt_ids = z["gt_ids"]
        self.gt_dists = z["gt_dists"]
        self.dtype = formats.dtype_name(self.base)
        self.N, self.D = self.base.shape
        self.R = self.nbrs.shape[1]
        self.m = self.codes.shape[1]
        self.prefix = os.path.join(tmpdir, name)
        self.paths = formats.write_index(self.prefix, self.base, self.deg, self.nbrs, self.medoid, self.pivots,
                                         self.centroid, self.chunk_offsets, self.codes)
        formats.write_bin(self.paths.query, self.queries)
        formats.write_truthset(self.paths.truth, self.gt_ids, self.gt_dists)

    def oracle(self):
        import oracle as O
        disk = formats.pack_disk_bin(self.base, self.deg, self.nbrs)
        return O.OracleIndex(disk, self.dtype, self.D, self.R, self.medoid, self.codes, self.pivots, self.centroid,
                             self.chunk_offsets)


@pytest.fixture(scope="session")
def fx_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("bang_fx"))


@pytest.fixture(scope="session")
def fx_u8(fx_dir):
    return Fixture("fx_u8", fx_dir)


@pytest.fixture(scope="session")
def fx_f32(fx_dir):
    return Fixture("fx_f32", fx_dir)


@pytest.fixture(scope="session")
def fx_i8(fx_dir):
    return Fixture("fx_i8", fx_dir)


@pytest.fixture(scope="session")
def fx_c1(fx_dir):
    """C1 (BASELINE config 1): SIFT10K shape — N = 10^4, D = 128 uint8, 100 queries, PQ m = 32."""
    return Fixture("fx_c1", fx_dir)


@pytest.fixture(scope="session")
def fx_c1m128(fx_dir):
    """C1 with the reference's SIFT1BSMALL chunk count, m = 128 (BANG_Inmemory/parANN.h:87)."""
    return Fixture("fx_c1m128", fx_dir, pq="128", npz="fx_c1")


@pytest.fixture(scope="session")
def fixtures(fx_u8, fx_f32, fx_i8):
    return {"fx_u8": fx_u8, "fx_f32": fx_f32, "fx_i8": fx_i8}
