"""ctypes wrapper over oracle/libbang_oracle.so (TEST INFRASTRUCTURE — see bang_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbang_oracle.so")

MODE_BASE, MODE_INMEMORY, MODE_EXACT = 0, 1, 2
ORDER_REF, ORDER_GPU = 0, 1
NO_ID = 0xFFFFFFFF
_DT = {"int8": 0, "uint8": 1, "float": 2}
_NP = {"int8": np.int8, "uint8": np.uint8, "float": np.float32}


class _Index(ctypes.Structure):
    _fields_ = [("N", ctypes.c_uint32), ("D", ctypes.c_uint32), ("R", ctypes.c_uint32), ("n_chunks", ctypes.c_uint32),
                ("dtype", ctypes.c_int32), ("medoid", ctypes.c_uint64), ("entry_len", ctypes.c_uint64),
                ("disk", ctypes.c_void_p), ("codes", ctypes.c_void_p), ("pivots", ctypes.c_void_p),
                ("centroid", ctypes.c_void_p), ("chunk_offsets", ctypes.c_void_p)]


class _Stats(ctypes.Structure):
    _fields_ = [("hops", ctypes.c_void_p), ("sum_deg", ctypes.c_void_p), ("n_cand", ctypes.c_void_p),
                ("trace", ctypes.c_void_p), ("trace_len", ctypes.c_uint32)]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, "bang_oracle.c"), os.path.join(_HERE, "bang_oracle.h")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in src)):
        return LIB_PATH
    subprocess.run(["make", "-C", _HERE, "libbang_oracle.so"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.bang_oracle_hash1.restype = ctypes.c_uint32
        _lib.bang_oracle_hash1.argtypes = [ctypes.c_uint32]
        _lib.bang_oracle_hash2.restype = ctypes.c_uint32
        _lib.bang_oracle_hash2.argtypes = [ctypes.c_uint32]
        _lib.bang_oracle_pq_table.restype = None
        _lib.bang_oracle_pq_table.argtypes = [ctypes.POINTER(_Index), ctypes.c_void_p, ctypes.c_void_p]
        _lib.bang_oracle_pq_dist.restype = ctypes.c_float
        _lib.bang_oracle_pq_dist.argtypes = [ctypes.POINTER(_Index), ctypes.c_void_p, ctypes.c_uint32]
        _lib.bang_oracle_l2.restype = ctypes.c_float
        _lib.bang_oracle_l2.argtypes = [ctypes.POINTER(_Index), ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        _lib.bang_oracle_search.restype = ctypes.c_int
        _lib.bang_oracle_search.argtypes = [ctypes.POINTER(_Index), ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32,
                                            ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(_Stats)]
        _lib.bang_oracle_bruteforce.restype = ctypes.c_int
        _lib.bang_oracle_bruteforce.argtypes = [ctypes.POINTER(_Index), ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                                ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def hash1(x: int) -> int:
    return lib().bang_oracle_hash1(x)


def hash2(x: int) -> int:
    return lib().bang_oracle_hash2(x)


class OracleIndex:
    """Holds numpy views of an index in the reference's file layouts."""

    def __init__(self, disk: np.ndarray, dtype: str, D: int, R: int, medoid: int, codes=None, pivots=None,
                 centroid=None, chunk_offsets=None):
        self.dtype = dtype
        self.D, self.R, self.medoid = int(D), int(R), int(medoid)
        self.entry_len = D * np.dtype(_NP[dtype]).itemsize + 4 + 4 * R
        self.disk = np.ascontiguousarray(disk).view(np.uint8).reshape(-1)
        self.N = self.disk.size // self.entry_len
        self.codes = None if codes is None else np.ascontiguousarray(codes, dtype=np.uint8)
        self.pivots = None if pivots is None else np.ascontiguousarray(pivots, dtype=np.float32)
        self.centroid = None if centroid is None else np.ascontiguousarray(centroid, dtype=np.float32).reshape(-1)
        self.chunk_offsets = None if chunk_offsets is None else np.ascontiguousarray(chunk_offsets, dtype=np.uint32).reshape(-1)
        self.n_chunks = 0 if self.codes is None else self.codes.shape[1]
        p = lambda a: None if a is None else a.ctypes.data
        self._c = _Index(self.N, self.D, self.R, self.n_chunks, _DT[dtype], self.medoid, self.entry_len,
                         p(self.disk), p(self.codes), p(self.pivots), p(self.centroid), p(self.chunk_offsets))

    @classmethod
    def from_files(cls, prefix: str, with_pq: bool = True, mmap: bool = False):
        """mmap: map `_disk.bin` and the PQ codes instead of reading them (indices of tens of GB: no second copy in RAM)."""
        import sys
        sys.path.insert(0, os.path.dirname(_HERE))
        import bang_b200  # noqa: F401
        from bang_b200 import formats
        paths = formats.IndexPaths(prefix)
        meta = formats.read_disk_metadata(paths.disk_meta)
        disk = np.memmap(paths.disk, dtype=np.uint8, mode="r") if mmap else np.fromfile(paths.disk, dtype=np.uint8)
        codes = piv = cen = offs = None
        if with_pq and mmap:
            hdr = np.fromfile(paths.pq_compressed, dtype=np.int32, count=2)
            codes = np.memmap(paths.pq_compressed, dtype=np.uint8, mode="r", offset=8, shape=(int(hdr[0]), int(hdr[1])))
            piv, cen, offs = formats.read_pq_pivots_new(paths.pq_pivots, meta.D, codes.shape[1])
        elif with_pq:
            codes = formats.read_bin(paths.pq_compressed, np.uint8)
            piv, cen, offs = formats.read_pq_pivots_new(paths.pq_pivots, meta.D, codes.shape[1])
        return cls(disk, meta.dtype, meta.D, meta.R, meta.medoid, codes, piv, cen, offs)

    def _q(self, queries):
        q = np.ascontiguousarray(queries, dtype=_NP[self.dtype])
        if q.ndim == 1:
            q = q[None, :]
        assert q.shape[1] == self.D
        return q

    def pq_table(self, query) -> np.ndarray:
        q = self._q(query)
        out = np.empty((self.n_chunks, 256), dtype=np.float32)
        lib().bang_oracle_pq_table(ctypes.byref(self._c), q.ctypes.data, out.ctypes.data)
        return out

    def pq_dist(self, tbl: np.ndarray, node: int) -> float:
        tbl = np.ascontiguousarray(tbl, dtype=np.float32)
        return lib().bang_oracle_pq_dist(ctypes.byref(self._c), tbl.ctypes.data, node)

    def l2(self, node: int, query, order=ORDER_GPU, kind=0) -> float:
        q = self._q(query)
        return lib().bang_oracle_l2(ctypes.byref(self._c), node, q.ctypes.data, order, kind)

    def search(self, queries, k: int, L: int, mode: int = MODE_BASE, order: int = ORDER_GPU, nthreads: int = 0,
               stats: bool = False, trace_len: int = 0):
        q = self._q(queries)
        Q = q.shape[0]
        ids = np.empty((Q, k), dtype=np.uint64)
        dists = np.empty((Q, k), dtype=np.float32)
        st = None
        extra = {}
        if stats or trace_len:
            extra = dict(hops=np.zeros(Q, np.uint32), sum_deg=np.zeros(Q, np.uint32), n_cand=np.zeros(Q, np.uint32))
            tr = np.zeros((Q, max(trace_len, 1)), np.uint32) if trace_len else None
            st = _Stats(extra["hops"].ctypes.data, extra["sum_deg"].ctypes.data, extra["n_cand"].ctypes.data,
                        None if tr is None else tr.ctypes.data, trace_len)
            if tr is not None:
                extra["trace"] = tr
        rc = lib().bang_oracle_search(ctypes.byref(self._c), mode, q.ctypes.data, Q, k, L, order, nthreads,
                                      ids.ctypes.data, dists.ctypes.data, None if st is None else ctypes.byref(st))
        if rc != 0:
            raise RuntimeError(f"bang_oracle_search failed: {rc}")
        return (ids, dists, extra) if (stats or trace_len) else (ids, dists)

    def bruteforce(self, queries, k: int, nthreads: int = 0):
        q = self._q(queries)
        ids = np.empty((q.shape[0], k), dtype=np.uint32)
        dists = np.empty((q.shape[0], k), dtype=np.float32)
        lib().bang_oracle_bruteforce(ctypes.byref(self._c), q.ctypes.data, q.shape[0], k, nthreads, ids.ctypes.data,
                                     dists.ctypes.data)
        return ids, dists
