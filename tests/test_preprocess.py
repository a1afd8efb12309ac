"""DiskANN `_disk.index` -> BANG `_disk.bin` / `_disk_metadata.bin` (csrc/bang_preprocess.cpp) against golden outputs
of the reference's own converter (BANG_Base/bang_preprocess.py, run unmodified by tests/golden/make_preprocess_golden.py)."""
import os
import subprocess

import numpy as np
import pytest

from bang_b200 import build, formats

from conftest import GOLDEN

NAMES = {"u8": "uint8", "f32": "float", "i8": "int8"}


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "preprocess_golden.npz"))


@pytest.mark.parametrize("name", ["u8", "f32", "i8"])
def test_converter_reproduces_reference_script_bytes(golden, name, tmp_path):
    d, code, R, n = (int(x) for x in golden[f"{name}_args"])
    idx = str(tmp_path / "x_disk.index")
    dst = str(tmp_path / "x_disk.bin")
    golden[f"{name}_index"].tofile(idx)
    assert formats.convert_diskann_index(idx, dst, d, NAMES[name], R) == n
    assert np.array_equal(np.fromfile(dst, dtype=np.uint8), golden[f"{name}_bin"])
    assert np.array_equal(np.fromfile(str(tmp_path / "x_disk_metadata.bin"), dtype=np.uint8), golden[f"{name}_meta"])
    # and the result is a loadable BANG index: metadata parses, adjacency is ascending within the degree
    meta = formats.read_disk_metadata(str(tmp_path / "x_disk_metadata.bin"))
    assert (meta.N, meta.D, meta.R, meta.dtype) == (n, d, R, NAMES[name])
    vec, deg, nb = formats.read_disk_bin(dst, meta)
    assert vec.shape == (n, d) and int(deg.min()) >= 1 and int(deg.max()) <= R
    for i in range(n):
        row = nb[i, :deg[i]].astype(np.int64)
        assert np.all(np.diff(row) >= 0)


def test_cli_same_arguments_as_the_script(golden, tmp_path):
    build.build_preprocess()
    d, code, R, n = (int(x) for x in golden["u8_args"])
    idx = str(tmp_path / "y_disk.index")
    dst = str(tmp_path / "y_disk.bin")
    golden["u8_index"].tofile(idx)
    out = subprocess.run([build.CLI_PREPROCESS, idx, dst, str(d), str(code), str(R)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and f"Total # of Nodes Discovered = {n}" in out.stdout
    assert np.array_equal(np.fromfile(dst, dtype=np.uint8), golden["u8_bin"])
    assert subprocess.run([build.CLI_PREPROCESS, idx], capture_output=True, text=True).returncode == 1   # usage


def test_round_trip_and_errors(tmp_path):
    rng = np.random.default_rng(1)
    n, d, R = 500, 24, 12
    vec = rng.normal(size=(n, d)).astype(np.float32)
    deg = rng.integers(1, R + 1, size=n).astype(np.uint32)
    nbrs = rng.integers(0, n, size=(n, R)).astype(np.uint32)
    idx = str(tmp_path / "z_disk.index")
    dst = str(tmp_path / "z_disk.bin")
    formats.write_diskann_index(idx, vec, deg, nbrs, medoid=7)
    assert formats.convert_diskann_index(idx, dst, d, "float", R) == n
    meta = formats.read_disk_metadata(str(tmp_path / "z_disk_metadata.bin"))
    assert meta.medoid == 7 and meta.entry_len == d * 4 + 4 + 4 * R
    v2, d2, n2 = formats.read_disk_bin(dst, meta)
    assert np.array_equal(v2, vec) and np.array_equal(d2, deg)
    for i in range(n):
        assert np.array_equal(n2[i, :deg[i]], np.sort(nbrs[i, :deg[i]]))
        assert np.array_equal(n2[i, deg[i]:], nbrs[i, deg[i]:])          # unused slots are carried over as they are
    # an index built with a larger degree bound than the R asked for: entries are found at the index's own stride
    deg_small = np.minimum(deg, 8).astype(np.uint32)
    formats.write_diskann_index(idx, vec, deg_small, nbrs, medoid=7)
    assert formats.convert_diskann_index(idx, dst, d, "float", 8) == n
    meta8 = formats.read_disk_metadata(str(tmp_path / "z_disk_metadata.bin"))
    meta8.entry_len = d * 4 + 4 + 4 * 8      # the metadata keeps the index's max_node_len, as the script does
    v3, d3, n3 = formats.read_disk_bin(dst, meta8)
    assert np.array_equal(v3, vec) and np.array_equal(d3, deg_small)
    # degree 0 or > R aborts (bang_preprocess.py:86-89)
    bad = deg.copy(); bad[123] = 0
    formats.write_diskann_index(idx, vec, bad, nbrs, medoid=7)
    with pytest.raises(ValueError, match="node 123 has degree 0"):
        formats.convert_diskann_index(idx, dst, d, "float", R)
    formats.write_diskann_index(idx, vec, deg, nbrs, medoid=7)
    assert int(deg.max()) == R
    with pytest.raises(ValueError, match=r"has degree \d+ \(must be 1\.\.8\)"):
        formats.convert_diskann_index(idx, dst, d, "float", 8)
    with pytest.raises(ValueError, match="cannot open"):
        formats.convert_diskann_index(str(tmp_path / "missing.index"), dst, d, "float", R)
    # truncated file
    raw = np.fromfile(idx, dtype=np.uint8)
    raw[:4096 * 3 + 100].tofile(idx)
    with pytest.raises(ValueError, match="ends inside node"):
        formats.convert_diskann_index(idx, dst, d, "float", R)
