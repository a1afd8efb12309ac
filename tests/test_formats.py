"""File formats on the search path: numpy writers/readers round-trip and the size checks
the reference performs (load_bin_impl bang_search.cuh:299-311, load_truthset test_driver.cpp:254-266)."""
import os
import struct

import numpy as np
import pytest

from bang_b200 import formats


def test_known_entry_lengths():
    # cross-check against the reference's dataset macros (BANG_Inmemory/parANN.h:47-151)
    assert formats.entry_len(128, "uint8", 64) == 388
    assert formats.entry_len(128, "float", 64) == 772
    assert formats.entry_len(960, "float", 64) == 4100
    assert formats.entry_len(96, "float", 64) == 644
    assert formats.entry_len(200, "float", 64) == 1060
    assert formats.entry_len(256, "float", 64) == 1284
    assert formats.entry_len(784, "uint8", 64) == 1044


def test_bin_roundtrip_and_size_check(tmp_path):
    a = np.arange(60, dtype=np.float32).reshape(12, 5)
    p = str(tmp_path / "a.bin")
    formats.write_bin(p, a)
    assert os.path.getsize(p) == 8 + a.nbytes
    assert np.array_equal(formats.read_bin(p, np.float32), a)
    assert np.array_equal(formats.read_bin(p, np.float32, max_rows=3), a[:3])
    with open(p, "ab") as f:
        f.write(b"x")
    with pytest.raises(ValueError):
        formats.read_bin(p, np.float32)


def test_metadata_is_32_packed_bytes(tmp_path):
    p = str(tmp_path / "m.bin")
    m = formats.GraphMeta(medoid=178757270, entry_len=388, dtype="uint8", D=128, R=64, N=1000000000)
    formats.write_disk_metadata(p, m)
    raw = open(p, "rb").read()
    assert len(raw) == 32  # GraphMedataData, bang_search.cuh:42-50
    assert struct.unpack("<QQiIII", raw) == (178757270, 388, 1, 128, 64, 1000000000)
    assert formats.read_disk_metadata(p) == m


def test_disk_bin_layout(fx_u8):
    fx = fx_u8
    raw = np.fromfile(fx.paths.disk, dtype=np.uint8).reshape(fx.N, 388 - 128 + fx.D)
    i = 17
    assert np.array_equal(raw[i, :fx.D], fx.base[i])
    deg = struct.unpack("<I", raw[i, fx.D:fx.D + 4].tobytes())[0]
    assert deg == fx.deg[i] and 1 <= deg <= 64
    nb = np.frombuffer(raw[i, fx.D + 4:].tobytes(), dtype="<u4")
    assert np.array_equal(nb, fx.nbrs[i])
    assert (np.diff(nb[:deg].astype(np.int64)) > 0).all()  # ascending (bang_preprocess.py:102-104)
    meta = formats.read_disk_metadata(fx.paths.disk_meta)
    v, d, n = formats.read_disk_bin(fx.paths.disk, meta)
    assert np.array_equal(v, fx.base) and np.array_equal(d, fx.deg) and np.array_equal(n, fx.nbrs)


def test_pq_pivots_new_layout(fx_f32):
    fx = fx_f32
    with open(fx.paths.pq_pivots, "rb") as f:
        assert struct.unpack("<ii", f.read(8)) == (4, 1)  # uNumPQSectionOffsets == 4 (bang_search.cu:247)
        off = struct.unpack("<QQQQ", f.read(32))
        f.seek(off[0])
        assert struct.unpack("<ii", f.read(8)) == (256, fx.D)
        assert off[3] == os.path.getsize(fx.paths.pq_pivots)
    piv, cen, chk = formats.read_pq_pivots_new(fx.paths.pq_pivots, fx.D, fx.m)
    assert np.array_equal(piv, fx.pivots) and np.array_equal(cen, fx.centroid) and np.array_equal(chk, fx.chunk_offsets)
    # old three-file layout holds the same numbers (parANN.cu:146-147,216,221)
    assert np.array_equal(formats.read_bin(fx.paths.old_pivots, np.float32), fx.pivots)
    assert np.array_equal(formats.read_bin(fx.paths.old_centroid, np.float32).ravel(), fx.centroid)
    assert np.array_equal(formats.read_bin(fx.paths.old_chunk_offsets, np.uint32).ravel(), fx.chunk_offsets)


def test_truthset_roundtrip(tmp_path, fx_u8):
    ids, d = formats.read_truthset(fx_u8.paths.truth)
    assert np.array_equal(ids, fx_u8.gt_ids) and np.array_equal(d, fx_u8.gt_dists)
    p = str(tmp_path / "gt.bin")
    formats.write_truthset(p, ids, d)
    assert os.path.getsize(p) == 8 + 8 * ids.size
    with open(p, "ab") as f:
        f.write(b"12345678")
    with pytest.raises(ValueError):
        formats.read_truthset(p)


def test_cli_mips_query_normaliser(tmp_path):
    """`bang_search <query.bin> <n>` = the reference driver's 2-argument form (test_driver.cpp:280-336,566-571):
    every float query is scaled to unit norm, one zero dimension is appended, result in <file>_transformed."""
    import subprocess
    from bang_b200 import build
    exe = build.build_cli()
    rng = np.random.default_rng(3)
    q = rng.normal(size=(17, 24)).astype(np.float32)
    path = str(tmp_path / "q.bin")
    formats.write_bin(path, q)
    out = subprocess.run([exe, path, "11"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    got = formats.read_bin(path + "_transformed", np.float32)
    assert got.shape == (11, 25)
    want = np.zeros((11, 25), np.float32)
    for i in range(11):
        norm = np.float32(0)
        for j in range(24):
            norm = np.float32(norm + q[i, j] * q[i, j])     # float accumulation in element order
        want[i, :24] = q[i] / np.sqrt(norm)
    assert np.array_equal(got[:, 24], np.zeros(11, np.float32))
    assert np.allclose(got, want, rtol=2e-7, atol=0)          # (fma contraction may differ by an ulp)
    # too many queries requested: error, no crash
    assert subprocess.run([exe, path, "18"], capture_output=True, text=True, timeout=120).returncode != 0


def test_inmemory_cli_argument_set(tmp_path):
    """`bang` takes the 15 positional arguments of the Inmemory / Exactdistance forks (parANN.cu:79-93); without a GPU
    it gets as far as creating the handle and reports the CUDA error (no CPU fallback)."""
    import subprocess
    from bang_b200 import build
    from conftest import has_gpu
    exe = build.build_cli_inmem()
    assert subprocess.run([exe], capture_output=True, text=True).returncode == 1
    args = ["p.bin", "c.bin", "g.bin", "q.bin", "o.bin", "cen.bin", "gt.bin", "10", "1", "256", "512", "256", "10", "64", "0"]
    r = subprocess.run([exe] + args, capture_output=True, text=True, env={k: v for k, v in os.environ.items() if k != "BANG_B200_MEDOID"})
    assert r.returncode == 1 and "medoid" in r.stdout
    r = subprocess.run([exe] + args + ["5"], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not open the file4" in r.stdout
    q = np.zeros((12, 8), np.uint8)
    formats.write_bin(str(tmp_path / "q.bin"), q)
    formats.write_bin(str(tmp_path / "c.bin"), np.zeros((100, 4), np.uint8))
    a2 = [str(tmp_path / "p.bin"), str(tmp_path / "c.bin"), str(tmp_path / "g.bin"), str(tmp_path / "q.bin")] + args[4:]
    r = subprocess.run([exe] + a2[:7] + ["13"] + a2[8:] + ["5"], capture_output=True, text=True)
    assert r.returncode == 1 and "out of range" in r.stdout          # more queries than the file holds
    if not has_gpu():
        r = subprocess.run([exe] + a2 + ["5", "20"], capture_output=True, text=True)
        assert r.returncode == 2 and ("no CUDA device" in r.stdout or "cudaGetDeviceCount" in r.stdout)
