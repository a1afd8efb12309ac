"""Readers / writers for every on-disk format on BANG's search path.

All formats are little-endian and are the reference's own:

* "bin"             int32 npts, int32 dim, data[npts][dim]
                    (reference reader: BANG_Base/bang_search.cuh:287-339, writer test_driver.cpp:215-235)
* _pq_compressed    bin of uint8[N][m]                          (bang_search.cu:218-234)
* _pq_pivots (new)  4-section offset table, then three bins     (bang_search.cu:246-296)
* _pq_pivots (old)  three separate bins                         (BANG_Inmemory/parANN.cu:146-147,216,221)
* _disk.bin         N entries of  T vec[D]; u32 degree; u32 nbr[R]   (bang_preprocess.py:81-110)
* _disk_metadata    packed 32 B: u64 medoid, u64 entry_len, i32 dtype, u32 D, u32 R, u32 N
                    (bang_search.cuh:42-50, bang_preprocess.py:42-51,116)
* truthset          int32 nq, int32 K, u32 ids[nq][K], f32 dists[nq][K]   (test_driver.cpp:238-272)

These are numpy-only helpers used by the fixture builder, the tests and bench.py.  The product's
loader is the C++ one inside libbang_b200.so (csrc/loader.cpp); tests check both agree.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

# bang_preprocess.py:13  "0 -> int8, 1 -> uint8, 2 -> float"
DTYPE_CODE = {"int8": 0, "uint8": 1, "float": 2}
CODE_DTYPE = {v: k for k, v in DTYPE_CODE.items()}
NP_DTYPE = {"int8": np.int8, "uint8": np.uint8, "float": np.float32}

PQ_PIVOTS_SUFFIX = "_pq_pivots.bin"
PQ_COMPRESSED_SUFFIX = "_pq_compressed.bin"
DISK_SUFFIX = "_disk.bin"
DISK_META_SUFFIX = "_disk_metadata.bin"
# old (DiskANN 0.1/0.2) layout used by BANG_Inmemory / BANG_Exactdistance
OLD_CENTROID_SUFFIX = "_pq_pivots.bin_centroid.bin"
OLD_CHUNK_OFFSETS_SUFFIX = "_pq_pivots.bin_chunk_offsets.bin"


def dtype_name(arr_or_dtype) -> str:
    dt = np.dtype(arr_or_dtype.dtype if hasattr(arr_or_dtype, "dtype") else arr_or_dtype)
    for name, npdt in NP_DTYPE.items():
        if dt == np.dtype(npdt):
            return name
    raise ValueError(f"unsupported element type {dt}")


# ----------------------------------------------------------------------------------------------
# generic bin
# ----------------------------------------------------------------------------------------------
def write_bin(path: str, data: np.ndarray) -> None:
    data = np.ascontiguousarray(data)
    if data.ndim == 1:
        data = data[:, None]
    assert data.ndim == 2
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", data.shape[0], data.shape[1]))
        f.write(data.tobytes())


def read_bin(path: str, dtype, max_rows: int | None = None) -> np.ndarray:
    dtype = np.dtype(dtype)
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        npts, dim = struct.unpack("<ii", f.read(8))
        expect = 8 + npts * dim * dtype.itemsize
        if size != expect:
            # same check as load_bin_impl, bang_search.cuh:299-311
            raise ValueError(f"{path}: file size {size} != expected {expect} (npts={npts} dim={dim})")
        rows = npts if max_rows is None else min(npts, max_rows)
        out = np.fromfile(f, dtype=dtype, count=rows * dim)
    return out.reshape(rows, dim)


# ----------------------------------------------------------------------------------------------
# PQ files
# ----------------------------------------------------------------------------------------------
def write_pq_pivots_new(path: str, pivots: np.ndarray, centroid: np.ndarray, chunk_offsets: np.ndarray) -> None:
    """New DiskANN layout read by BANG_Base (bang_search.cu:207-213,246-296)."""
    pivots = np.ascontiguousarray(pivots, dtype=np.float32)
    centroid = np.ascontiguousarray(centroid, dtype=np.float32).reshape(-1)
    chunk_offsets = np.ascontiguousarray(chunk_offsets, dtype=np.uint32).reshape(-1)
    assert pivots.shape[0] == 256 and pivots.shape[1] == centroid.shape[0]
    D = pivots.shape[1]
    # DiskANN places the offset table in the first 4 KiB; BANG only needs the offsets to be right.
    off_piv = 4096
    off_cen = off_piv + 8 + 256 * D * 4
    off_chk = off_cen + 8 + D * 4
    end = off_chk + 8 + chunk_offsets.shape[0] * 4
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", 4, 1))  # "npts=4, dim=1": uNumPQSectionOffsets must be 4 (bang_search.cu:247)
        f.write(struct.pack("<QQQQ", off_piv, off_cen, off_chk, end))
        f.write(b"\0" * (off_piv - f.tell()))
        f.write(struct.pack("<ii", 256, D))
        f.write(pivots.tobytes())
        f.write(struct.pack("<ii", D, 1))
        f.write(centroid.tobytes())
        f.write(struct.pack("<ii", chunk_offsets.shape[0], 1))
        f.write(chunk_offsets.tobytes())
        assert f.tell() == end


def read_pq_pivots_new(path: str, D: int, n_chunks: int):
    with open(path, "rb") as f:
        nsec = struct.unpack("<I", f.read(4))[0]
        if nsec != 4:
            raise ValueError("PQ pivots file does not contain the required number of sub-sections")
        f.seek(8)
        off_piv, off_cen, off_chk, _end = struct.unpack("<QQQQ", f.read(32))
        f.seek(off_piv + 8)
        piv = np.fromfile(f, dtype=np.float32, count=256 * D).reshape(256, D)
        f.seek(off_cen + 8)
        cen = np.fromfile(f, dtype=np.float32, count=D)
        f.seek(off_chk + 8)
        chk = np.fromfile(f, dtype=np.uint32, count=n_chunks + 1)
    return piv, cen, chk


def write_pq_pivots_old(prefix_pivots_path: str, pivots, centroid, chunk_offsets) -> None:
    """Old layout: <p>_pq_pivots.bin + ..._centroid.bin + ..._chunk_offsets.bin (parANN.cu:146-147,216,221)."""
    write_bin(prefix_pivots_path, np.asarray(pivots, dtype=np.float32))
    write_bin(prefix_pivots_path + "_centroid.bin", np.asarray(centroid, dtype=np.float32).reshape(-1, 1))
    write_bin(prefix_pivots_path + "_chunk_offsets.bin", np.asarray(chunk_offsets, dtype=np.uint32).reshape(-1, 1))


# ----------------------------------------------------------------------------------------------
# graph ("_disk.bin") + metadata
# ----------------------------------------------------------------------------------------------
@dataclass
class GraphMeta:
    medoid: int
    entry_len: int
    dtype: str
    D: int
    R: int
    N: int


def entry_len(D: int, dtype: str, R: int) -> int:
    return D * np.dtype(NP_DTYPE[dtype]).itemsize + 4 + 4 * R


def pack_disk_bin(vectors: np.ndarray, degrees: np.ndarray, nbrs: np.ndarray) -> np.ndarray:
    """vectors [N][D] T, degrees [N] u32, nbrs [N][R] u32 (first degree[i] entries valid, ascending)."""
    N, D = vectors.shape
    R = nbrs.shape[1]
    dt = dtype_name(vectors)
    el = entry_len(D, dt, R)
    vb = D * vectors.dtype.itemsize
    out = np.zeros((N, el), dtype=np.uint8)
    out[:, :vb] = np.ascontiguousarray(vectors).view(np.uint8).reshape(N, vb)
    out[:, vb:vb + 4] = np.ascontiguousarray(degrees, dtype=np.uint32).view(np.uint8).reshape(N, 4)
    out[:, vb + 4:] = np.ascontiguousarray(nbrs, dtype=np.uint32).view(np.uint8).reshape(N, 4 * R)
    return out


def write_disk_bin(path: str, vectors, degrees, nbrs) -> None:
    """Packs and writes in 256 MB pieces; large files (10^8 points: 64 GB) use several threads, each writing its pieces
    at their offsets (numpy's copies release the GIL)."""
    N = vectors.shape[0]
    el = entry_len(vectors.shape[1], dtype_name(vectors), nbrs.shape[1])
    step = max(1, (256 << 20) // max(1, el))
    starts = list(range(0, N, step))
    with open(path, "wb") as f:
        f.truncate(N * el)
    fd = os.open(path, os.O_WRONLY)
    try:
        def piece(s):
            buf = memoryview(pack_disk_bin(vectors[s:s + step], degrees[s:s + step], nbrs[s:s + step]).reshape(-1))
            off = s * el
            while len(buf):
                w = os.pwrite(fd, buf[:1 << 30], off)
                buf, off = buf[w:], off + w
        if len(starts) <= 4:
            for s in starts:
                piece(s)
        else:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
                list(ex.map(piece, starts))
    finally:
        os.close(fd)


def write_disk_metadata(path: str, meta: GraphMeta) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<QQiIII", meta.medoid, meta.entry_len, DTYPE_CODE[meta.dtype], meta.D, meta.R, meta.N))


def read_disk_metadata(path: str) -> GraphMeta:
    with open(path, "rb") as f:
        medoid, el, dt, D, R, N = struct.unpack("<QQiIII", f.read(32))
    return GraphMeta(medoid, el, CODE_DTYPE.get(dt, "uint8"), D, R, N)


def read_disk_bin(path: str, meta: GraphMeta):
    raw = np.fromfile(path, dtype=np.uint8).reshape(meta.N, meta.entry_len)
    npdt = np.dtype(NP_DTYPE[meta.dtype])
    vb = meta.D * npdt.itemsize
    vec = np.ascontiguousarray(raw[:, :vb]).view(npdt).reshape(meta.N, meta.D)
    deg = np.ascontiguousarray(raw[:, vb:vb + 4]).view(np.uint32).reshape(meta.N)
    nbr = np.ascontiguousarray(raw[:, vb + 4:]).view(np.uint32).reshape(meta.N, meta.R)
    return vec, deg, nbr


# ----------------------------------------------------------------------------------------------
# truthset
# ----------------------------------------------------------------------------------------------
def write_truthset(path: str, ids: np.ndarray, dists: np.ndarray) -> None:
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    dists = np.ascontiguousarray(dists, dtype=np.float32)
    assert ids.shape == dists.shape and ids.ndim == 2
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", ids.shape[0], ids.shape[1]))
        f.write(ids.tobytes())
        f.write(dists.tobytes())


def read_truthset(path: str):
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        nq, K = struct.unpack("<ii", f.read(8))
        if size != 8 + 8 * nq * K:  # test_driver.cpp:254-266
            raise ValueError(f"{path}: truthset size mismatch")
        ids = np.fromfile(f, dtype=np.uint32, count=nq * K).reshape(nq, K)
        dists = np.fromfile(f, dtype=np.float32, count=nq * K).reshape(nq, K)
    return ids, dists


# ----------------------------------------------------------------------------------------------
# a whole index on disk, both layouts at once
# ----------------------------------------------------------------------------------------------
@dataclass
class IndexPaths:
    prefix: str

    @property
    def pq_pivots(self): return self.prefix + PQ_PIVOTS_SUFFIX
    @property
    def pq_compressed(self): return self.prefix + PQ_COMPRESSED_SUFFIX
    @property
    def disk(self): return self.prefix + DISK_SUFFIX
    @property
    def disk_meta(self): return self.prefix + DISK_META_SUFFIX
    @property
    def old_pivots(self): return self.prefix + "_old" + PQ_PIVOTS_SUFFIX
    @property
    def old_centroid(self): return self.old_pivots + "_centroid.bin"
    @property
    def old_chunk_offsets(self): return self.old_pivots + "_chunk_offsets.bin"
    @property
    def query(self): return self.prefix + "_query.bin"
    @property
    def truth(self): return self.prefix + "_gt.bin"


def write_index(prefix: str, vectors, degrees, nbrs, medoid: int, pivots=None, centroid=None,
                chunk_offsets=None, codes=None) -> IndexPaths:
    p = IndexPaths(prefix)
    N, D = vectors.shape
    R = nbrs.shape[1]
    dt = dtype_name(vectors)
    write_disk_bin(p.disk, vectors, degrees, nbrs)
    write_disk_metadata(p.disk_meta, GraphMeta(int(medoid), entry_len(D, dt, R), dt, D, R, N))
    if pivots is not None:
        write_pq_pivots_new(p.pq_pivots, pivots, centroid, chunk_offsets)
        write_pq_pivots_old(p.old_pivots, pivots, centroid, chunk_offsets)
        write_bin(p.pq_compressed, np.asarray(codes, dtype=np.uint8))
    return p


# ----------------------------------------------------------------------------------------------------
# DiskANN `_disk.index` (what `build_disk_index` writes) and its conversion to `_disk.bin` + `_disk_metadata.bin`
# (BANG_Base/bang_preprocess.py; C++ here: csrc/bang_preprocess.cpp)
# ----------------------------------------------------------------------------------------------------
def write_diskann_index(path: str, vectors: np.ndarray, degrees: np.ndarray, nbrs: np.ndarray, medoid: int,
                        sector_len: int = 4096, rng: np.random.Generator | None = None) -> None:
    """Writes the on-disk layout the reference's converter reads (bang_preprocess.py:26-63,70-102): a metadata sector
    (two int32, then uint64 npts, ndims, medoid, max_node_len, nnodes_per_sector, 3 unused, file size) followed by
    sectors of nnodes_per_sector entries `T[D] | u32 degree | u32 nbr[R]` packed from the start of each sector.
    Neighbour slots beyond the degree hold arbitrary bytes in a real index: `rng` fills them with noise."""
    n, d = vectors.shape
    R = nbrs.shape[1]
    node_len = d * vectors.dtype.itemsize + 4 + 4 * R
    per_sector = sector_len // node_len
    if per_sector < 1:
        raise ValueError("node entry larger than a sector")
    n_sectors = 1 + (n + per_sector - 1) // per_sector
    nb = np.array(nbrs, dtype="<u4", copy=True)
    if rng is not None:
        noise = rng.integers(0, 2**32, size=nb.shape, dtype=np.uint64).astype("<u4")
        unused = np.arange(R)[None, :] >= np.asarray(degrees)[:, None]
        nb[unused] = noise[unused]
    entries = np.zeros((n, node_len), dtype=np.uint8)
    vb = d * vectors.dtype.itemsize
    entries[:, :vb] = np.ascontiguousarray(vectors).view(np.uint8).reshape(n, vb)
    entries[:, vb:vb + 4] = np.asarray(degrees, dtype="<u4").reshape(n, 1).view(np.uint8)
    entries[:, vb + 4:] = nb.view(np.uint8).reshape(n, 4 * R)
    buf = np.zeros(n_sectors * sector_len, dtype=np.uint8)
    hdr = struct.pack("<ii9Q", 9, 1, n, d, medoid, node_len, per_sector, 0, 0, 0, n_sectors * sector_len)
    buf[:len(hdr)] = np.frombuffer(hdr, dtype=np.uint8)
    for s in range(1, n_sectors):
        lo = (s - 1) * per_sector
        blk = entries[lo:lo + per_sector].reshape(-1)
        buf[s * sector_len:s * sector_len + blk.size] = blk
    buf.tofile(path)


def convert_diskann_index(index_path: str, out_bin_path: str, dim: int, dtype: str, degree: int, sector_len: int = 4096) -> int:
    """`bang_preprocess.py <index> <out.bin> <dim> <datatype> <R>` through csrc/bang_preprocess.cpp; returns the
    number of nodes written.  Creates out_bin_path and <out>_metadata.bin."""
    import ctypes
    from . import build
    lib = ctypes.CDLL(build.build_preprocess())
    lib.bang_preprocess_index.restype = ctypes.c_int
    lib.bang_preprocess_index.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                          ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint64)]
    lib.bang_preprocess_last_error.restype = ctypes.c_char_p
    n = ctypes.c_uint64(0)
    rc = lib.bang_preprocess_index(index_path.encode(), out_bin_path.encode(), dim, DTYPE_CODE[dtype], degree, sector_len, ctypes.byref(n))
    if rc != 0:
        raise ValueError(lib.bang_preprocess_last_error().decode())
    return int(n.value)
