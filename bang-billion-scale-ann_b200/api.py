"""Python mirror of the reference's `BANGSearch<T>` interface over the C ABI (include/bang_b200.h).

Method names, argument meaning and call order are the reference's (BANG_Base/bang.h:36-87,
test_driver.cpp:342,421-435,535,553):

    s = BANGSearch("uint8", mode="inmemory")
    s.bang_load(prefix); s.bang_set_searchparams(10, 64); s.bang_alloc(Q)
    s.bang_init(Q); ids, dists = s.bang_query(queries)
    s.bang_free(); s.bang_unload()

Everything runs in libbang_b200.so (hand-written sm_100a kernels).  There is no Python/CPU fallback:
if the library is missing or no GPU is present the calls raise.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

DT = {"int8": 0, "uint8": 1, "float": 2}
NP = {"int8": np.int8, "uint8": np.uint8, "float": np.float32}
MODES = {"base": 0, "inmemory": 1, "exact": 2, "exactdistance": 2}
ENUM_DIST_L2, ENUM_DIST_MIPS = 0, 1
DISTS_RANK_MAJOR, DISTS_QUERY_MAJOR = 0, 1
NO_ID = 0xFFFFFFFF

# every symbol include/bang_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "bang_b200_create", "bang_b200_destroy", "bang_b200_load", "bang_b200_load_files", "bang_b200_set_sharding",
    "bang_b200_export_shard", "bang_b200_import_shard", "bang_b200_export_shard_fd", "bang_b200_import_shard_fd",
    "bang_b200_load_device_begin", "bang_b200_load_device_rows",
    "bang_b200_load_device_codes", "bang_b200_load_device_codes_at", "bang_b200_load_device_end", "bang_b200_set_searchparams", "bang_b200_alloc",
    "bang_b200_init", "bang_b200_query", "bang_b200_free", "bang_b200_unload", "bang_b200_set_dists_layout",
    "bang_b200_query_device", "bang_b200_pq_table", "bang_b200_info", "bang_b200_filter_slots", "bang_b200_last_stats",
    "bang_b200_last_timing", "bang_b200_last_error", "bang_b200_bruteforce_gt", "bang_b200_pq_train", "bang_b200_pq_encode",
    "bang_b200_prep_last_error", "bang_load_c", "bang_set_searchparams_c", "bang_query_c",
    "bang_unload_c",
]


class BangError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class Info(ctypes.Structure):
    _fields_ = [("N", ctypes.c_uint64), ("medoid", ctypes.c_uint64), ("entry_len", ctypes.c_uint64),
                ("D", ctypes.c_uint32), ("R", ctypes.c_uint32), ("n_chunks", ctypes.c_uint32),
                ("dtype", ctypes.c_int32), ("mode", ctypes.c_int32), ("device_bytes", ctypes.c_uint64),
                ("row_stride", ctypes.c_uint32), ("slot_block", ctypes.c_uint32)]


class Timing(ctypes.Structure):
    _fields_ = [("kernel_ms", ctypes.c_float), ("launches", ctypes.c_uint32), ("h2d_bytes", ctypes.c_uint64),
                ("d2h_bytes", ctypes.c_uint64), ("grid", ctypes.c_uint32), ("block", ctypes.c_uint32),
                ("smem_bytes", ctypes.c_uint32), ("ctas_per_sm", ctypes.c_uint32)]


_lib = None


def load_library(path: str | None = None) -> ctypes.CDLL:
    """Loads libbang_b200.so (in-tree).  Raises if it is absent — there is no fallback path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("BANG_B200_LIB") or _build.LIB_CUDA
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(path)
    vp, ci, u64, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32
    sig = {
        "bang_b200_create": (ci, [ctypes.POINTER(vp), ci, ci, ci]),
        "bang_b200_destroy": (None, [vp]),
        "bang_b200_load": (ci, [vp, ctypes.c_char_p]),
        "bang_b200_load_files": (ci, [vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p,
                                      ctypes.c_char_p, u64, u32, u64]),
        "bang_b200_set_sharding": (ci, [vp, ci, ci]),
        "bang_b200_export_shard": (ci, [vp, vp]),
        "bang_b200_import_shard": (ci, [vp, ci, vp]),
        "bang_b200_export_shard_fd": (ci, [vp, ctypes.POINTER(ci), ctypes.POINTER(u64)]),
        "bang_b200_import_shard_fd": (ci, [vp, ci, ci, u64]),
        "bang_b200_load_device_begin": (ci, [vp, u64, u32, u64, u32, vp, vp, vp]),
        "bang_b200_load_device_rows": (ci, [vp, u64, u64, vp, vp]),
        "bang_b200_load_device_codes": (ci, [vp, u64, u64, vp]),
        "bang_b200_load_device_codes_at": (ci, [vp, vp, u64, vp]),
        "bang_b200_load_device_end": (ci, [vp]),
        "bang_b200_set_searchparams": (ci, [vp, ci, ci, ci]),
        "bang_b200_alloc": (ci, [vp, ci]),
        "bang_b200_init": (ci, [vp, ci]),
        "bang_b200_query": (ci, [vp, vp, ci, vp, vp]),
        "bang_b200_free": (ci, [vp]),
        "bang_b200_unload": (ci, [vp]),
        "bang_b200_set_dists_layout": (ci, [vp, ci]),
        "bang_b200_query_device": (ci, [vp, vp, ci, vp, vp, vp]),
        "bang_b200_pq_table": (ci, [vp, vp, ci, vp]),
        "bang_b200_info": (ci, [vp, ctypes.POINTER(Info)]),
        "bang_b200_filter_slots": (None, [ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32 * 2), ctypes.POINTER(ctypes.c_uint32 * 2)]),
        "bang_b200_last_stats": (ci, [vp, vp, vp, vp]),
        "bang_b200_last_timing": (ci, [vp, ctypes.POINTER(Timing)]),
        "bang_b200_last_error": (ctypes.c_char_p, []),
        "bang_b200_bruteforce_gt": (ci, [ci, vp, u64, u32, vp, u32, u32, u64, vp, vp, vp]),
        "bang_b200_pq_train": (ci, [ci, vp, u64, u32, vp, u32, u32, u64, u64, vp, vp, vp]),
        "bang_b200_pq_encode": (ci, [ci, vp, u64, u32, vp, vp, vp, u32, vp, vp]),
        "bang_b200_prep_last_error": (ctypes.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path == _build.LIB_CUDA:
        _lib = lib
    return lib


def filter_slots(point_id: int):
    """(positions, slot words) of a point id in the visited filter — host arithmetic of the library (bang_b200_filter_slots)."""
    lib = load_library()
    pos, words = (ctypes.c_uint32 * 2)(), (ctypes.c_uint32 * 2)()
    lib.bang_b200_filter_slots(ctypes.c_uint32(point_id), ctypes.byref(pos), ctypes.byref(words))
    return (pos[0], pos[1]), (words[0], words[1])


class BANGSearch:
    """`BANGSearch<T>` (bang.h:36-87).  dtype in {'uint8','int8','float'}; mode in {'base','inmemory','exact'}."""

    def __init__(self, dtype: str = "uint8", mode: str = "base", device: int = -1, lib_path: str | None = None):
        self._lib = load_library(lib_path)
        self.dtype = dtype
        self.mode = mode
        self._h = ctypes.c_void_p()
        self._k = self._L = 0
        self._dist = ENUM_DIST_L2
        self._layout = DISTS_RANK_MAJOR
        self._check(self._lib.bang_b200_create(ctypes.byref(self._h), DT[dtype], MODES[mode], device))

    # -- plumbing --
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise BangError(rc, self._lib.bang_b200_last_error().decode())

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self._lib.bang_b200_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    @property
    def handle(self) -> int:
        return self._h.value

    # -- the reference's seven verbs --
    def bang_load(self, indexfile_path_prefix: str) -> bool:
        rc = self._lib.bang_b200_load(self._h, indexfile_path_prefix.encode())
        if rc != 0:
            self.last_error = self._lib.bang_b200_last_error().decode()
            return False  # the reference returns false on missing / ill-formed files (bang_search.cu:152-177)
        return True

    def bang_load_files(self, pq_pivots, pq_compressed, disk, chunk_offsets, centroid, N: int, D: int, medoid: int) -> bool:
        """Inmemory/Exactdistance argument set (BANG_Inmemory/parANN.cu:79-93; N/D/MEDOID are macros there)."""
        enc = lambda s: None if s is None else s.encode()
        rc = self._lib.bang_b200_load_files(self._h, enc(pq_pivots), enc(pq_compressed), enc(disk), enc(chunk_offsets),
                                            enc(centroid), N, D, medoid)
        if rc != 0:
            self.last_error = self._lib.bang_b200_last_error().decode()
            return False
        return True

    def bang_set_searchparams(self, recall: int, worklist_length: int, nDistFunc: int = ENUM_DIST_L2) -> None:
        self._check(self._lib.bang_b200_set_searchparams(self._h, recall, worklist_length, nDistFunc))
        self._k, self._L, self._dist = recall, worklist_length, nDistFunc

    def bang_alloc(self, numQueries: int) -> None:
        self._check(self._lib.bang_b200_alloc(self._h, numQueries))

    def bang_init(self, numQueries: int) -> None:
        self._check(self._lib.bang_b200_init(self._h, numQueries))

    def bang_query(self, query_array: np.ndarray, num_queries: int | None = None, nearestNeighbours: np.ndarray | None = None,
                   nearestNeighbours_dist: np.ndarray | None = None):
        """Returns (ids u64 [Q][k], dists f32) — dists is [k][Q] (rank-major, the reference's layout,
        bang_search.cu:999) unless set_dists_layout(DISTS_QUERY_MAJOR) was called."""
        q = np.ascontiguousarray(query_array, dtype=NP[self.dtype])
        if q.ndim == 1:
            q = q[None, :]
        Q = q.shape[0] if num_queries is None else num_queries
        k = self._k
        ids = nearestNeighbours if nearestNeighbours is not None else np.empty((Q, k), dtype=np.uint64)
        shape = (Q, k) if self._layout == DISTS_QUERY_MAJOR else (k, Q)
        dists = nearestNeighbours_dist if nearestNeighbours_dist is not None else np.empty(shape, dtype=np.float32)
        assert ids.dtype == np.uint64 and ids.flags.c_contiguous and ids.size >= Q * k
        assert dists.dtype == np.float32 and dists.flags.c_contiguous and dists.size >= Q * k
        self._check(self._lib.bang_b200_query(self._h, q.ctypes.data, Q, ids.ctypes.data, dists.ctypes.data))
        return ids, dists

    def bang_free(self) -> None:
        self._check(self._lib.bang_b200_free(self._h))

    def bang_unload(self) -> None:
        self._check(self._lib.bang_b200_unload(self._h))

    # -- extensions --
    def set_dists_layout(self, layout: int) -> None:
        self._check(self._lib.bang_b200_set_dists_layout(self._h, layout))
        self._layout = layout

    def set_sharding(self, shard: int, n_shards: int) -> None:
        self._check(self._lib.bang_b200_set_sharding(self._h, shard, n_shards))

    def export_shard(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self._check(self._lib.bang_b200_export_shard(self._h, buf))
        return buf.raw

    def import_shard(self, shard: int, handle: bytes) -> None:
        buf = ctypes.create_string_buffer(handle, 64)
        self._check(self._lib.bang_b200_import_shard(self._h, shard, buf))

    def export_shard_fd(self) -> tuple[int, int]:
        """(file descriptor, bytes) of this rank's rows; only for rows allocated with BANG_B200_SHARD_VMM=1.
        The caller passes the descriptor to its peers (sharding.exchange_fds) and closes it."""
        fd, nbytes = ctypes.c_int(-1), ctypes.c_uint64(0)
        self._check(self._lib.bang_b200_export_shard_fd(self._h, ctypes.byref(fd), ctypes.byref(nbytes)))
        return fd.value, int(nbytes.value)

    def import_shard_fd(self, shard: int, fd: int, nbytes: int) -> None:
        self._check(self._lib.bang_b200_import_shard_fd(self._h, shard, fd, nbytes))

    # device-resident load (indices built on the GPUs, too large for files)
    def load_device_begin(self, N: int, D: int, medoid: int, pivots=None, centroid=None, chunk_offsets=None) -> None:
        m = 0 if chunk_offsets is None else len(chunk_offsets) - 1
        keep = [None if a is None else np.ascontiguousarray(a, dtype=dt) for a, dt in
                ((pivots, np.float32), (centroid, np.float32), (chunk_offsets, np.uint32))]
        ptr = lambda a: None if a is None else a.ctypes.data
        self._check(self._lib.bang_b200_load_device_begin(self._h, N, D, medoid, m, ptr(keep[0]), ptr(keep[1]), ptr(keep[2])))

    def load_device_rows(self, first_local_row: int, n_rows: int, d_vectors_ptr: int, d_adj_ptr: int) -> None:
        self._check(self._lib.bang_b200_load_device_rows(self._h, first_local_row, n_rows, d_vectors_ptr, d_adj_ptr))

    def load_device_codes(self, first_id: int, n: int, d_codes_ptr: int) -> None:
        self._check(self._lib.bang_b200_load_device_codes(self._h, first_id, n, d_codes_ptr))

    def load_device_codes_at(self, d_ids_ptr: int, n: int, d_codes_ptr: int) -> None:
        """codes of n nodes whose ids (u32, device) are not consecutive"""
        self._check(self._lib.bang_b200_load_device_codes_at(self._h, d_ids_ptr, n, d_codes_ptr))

    def load_device_end(self) -> None:
        self._check(self._lib.bang_b200_load_device_end(self._h))

    def query_device(self, d_queries_ptr: int, Q: int, d_ids_ptr: int, d_dists_ptr: int, stream: int = 0) -> None:
        self._check(self._lib.bang_b200_query_device(self._h, d_queries_ptr, Q, d_ids_ptr, d_dists_ptr, stream))

    def pq_table(self, queries: np.ndarray) -> np.ndarray:
        q = np.ascontiguousarray(queries, dtype=NP[self.dtype])
        if q.ndim == 1:
            q = q[None, :]
        m = self.info().n_chunks
        out = np.empty((q.shape[0], m, 256), dtype=np.float32)
        self._check(self._lib.bang_b200_pq_table(self._h, q.ctypes.data, q.shape[0], out.ctypes.data))
        return out

    def info(self) -> Info:
        out = Info()
        self._check(self._lib.bang_b200_info(self._h, ctypes.byref(out)))
        return out

    def last_stats(self, Q: int):
        hops, sd, nc = (np.zeros(Q, np.uint32) for _ in range(3))
        self._check(self._lib.bang_b200_last_stats(self._h, hops.ctypes.data, sd.ctypes.data, nc.ctypes.data))
        return dict(hops=hops, sum_deg=sd, n_cand=nc)

    def last_timing(self) -> Timing:
        out = Timing()
        self._check(self._lib.bang_b200_last_timing(self._h, ctypes.byref(out)))
        return out


# ---- data preparation kernels (csrc/prep_kernels.cu): CUDA tensors in, see include/bang_b200.h ----
def _prep_check(lib, rc: int) -> None:
    if rc != 0:
        raise BangError(rc, lib.bang_b200_prep_last_error().decode())


def _torch_dtype_name(t) -> str:
    import torch
    return {torch.uint8: "uint8", torch.int8: "int8", torch.float32: "float"}[t.dtype]


def bruteforce_gt(base, queries, k: int, id_offset: int = 0):
    """Exact kNN ground truth on the GPU.  base [n][D], queries [nq][D]: contiguous CUDA tensors of one element type.
    Returns CUDA tensors (ids int64 [nq][k] holding u32 values, dists float32 [nq][k]) ordered by (distance, id)."""
    import torch
    assert base.is_cuda and queries.is_cuda and base.is_contiguous() and queries.is_contiguous() and base.dtype == queries.dtype
    lib = load_library()
    nq = queries.shape[0]
    ids = torch.empty((nq, k), dtype=torch.int32, device=base.device)
    dists = torch.empty((nq, k), dtype=torch.float32, device=base.device)
    with torch.cuda.device(base.device):
        st = torch.cuda.current_stream(base.device).cuda_stream
        _prep_check(lib, lib.bang_b200_bruteforce_gt(DT[_torch_dtype_name(base)], base.data_ptr(), base.shape[0], base.shape[1],
                                                    queries.data_ptr(), nq, k, id_offset, ids.data_ptr(), dists.data_ptr(), st))
    return ids.to(torch.int64) & 0xFFFFFFFF, dists


def pq_train(base, chunk_offsets: np.ndarray, iters: int = 12, max_train: int = 65536, seed: int = 0x50):
    """k-means PQ pivots on the GPU.  Returns (pivots f32 [256][D], centroid f32 [D]) as numpy arrays."""
    import torch
    assert base.is_cuda and base.is_contiguous()
    lib = load_library()
    D = base.shape[1]
    offs = np.ascontiguousarray(chunk_offsets, dtype=np.uint32)
    piv = np.zeros((256, D), dtype=np.float32)
    cen = np.zeros(D, dtype=np.float32)
    with torch.cuda.device(base.device):
        st = torch.cuda.current_stream(base.device).cuda_stream
        _prep_check(lib, lib.bang_b200_pq_train(DT[_torch_dtype_name(base)], base.data_ptr(), base.shape[0], D, offs.ctypes.data,
                                               len(offs) - 1, iters, max_train, seed, piv.ctypes.data, cen.ctypes.data, st))
    return piv, cen


def pq_encode(base, pivots: np.ndarray, centroid: np.ndarray, chunk_offsets: np.ndarray):
    """PQ codes on the GPU: CUDA uint8 tensor [n][m]."""
    import torch
    assert base.is_cuda and base.is_contiguous()
    lib = load_library()
    D = base.shape[1]
    offs = np.ascontiguousarray(chunk_offsets, dtype=np.uint32)
    piv = np.ascontiguousarray(pivots, dtype=np.float32)
    cen = np.ascontiguousarray(centroid, dtype=np.float32)
    m = len(offs) - 1
    codes = torch.empty((base.shape[0], m), dtype=torch.uint8, device=base.device)
    with torch.cuda.device(base.device):
        st = torch.cuda.current_stream(base.device).cuda_stream
        _prep_check(lib, lib.bang_b200_pq_encode(DT[_torch_dtype_name(base)], base.data_ptr(), base.shape[0], D, piv.ctypes.data,
                                                cen.ctypes.data, offs.ctypes.data, m, codes.data_ptr(), st))
    return codes


def algorithmic_bytes(stats: dict, mode: str, D: int, elem_size: int, n_chunks: int, k: int) -> np.ndarray:
    """Per-query algorithmic bytes B_q (SURVEY.md §8d / BASELINE.md §4)."""
    hops = stats["hops"].astype(np.int64)
    sum_deg = stats["sum_deg"].astype(np.int64)
    n_cand = stats["n_cand"].astype(np.int64)
    vec = D * elem_size
    adj = 4 * hops + 4 * sum_deg
    if mode in ("exact", "exactdistance"):
        return adj + n_cand * vec + vec + 8 * k
    return adj + n_cand * n_chunks + hops * vec + vec + 8 * k
