/* bang_b200.h — C ABI of the B200-native (sm_100a) BANG search path.
 *
 * This is the drop-in boundary.  Every entry point replaces one method of the reference's public
 * class `BANGSearch<T>` (reference: BANG_Base/bang.h:36-87, implemented in bang_search.cu:70-135) or of
 * the reference's own — `#if 0`'d — C API (bang.h:89-101, bang_search.cu:1787-1806).  Plain pointers and
 * sizes only; no C++/torch types.  All functions return 0 on success and a negative code on failure
 * (bang_b200_last_error() gives the message); nothing here calls exit().
 *
 * Required call order (as the reference's driver, test_driver.cpp:342,421-435,535,553):
 *   create -> load -> { set_searchparams -> alloc(Q) -> n x ( init(Q) -> query ) -> free }* -> unload -> destroy
 *
 * There is NO CPU fallback: every query runs the fused sm_100a kernel; without a CUDA device the
 * calls fail with BANG_E_CUDA.
 */
#ifndef BANG_B200_H_
#define BANG_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BANG_B200_MAX_L 512        /* bang.h:20 MAX_L */
#define BANG_B200_MAX_R 64         /* bang_search.cu:35 MAX_R (asserted at load, :190) */
#define BANG_B200_BF_ENTRIES 399887u /* bang_search.cu:48 */
#define BANG_B200_NO_ID 0xFFFFFFFFull /* filler id when a query yields fewer than k results */

typedef enum { BANG_DT_INT8 = 0, BANG_DT_UINT8 = 1, BANG_DT_FLOAT = 2 } bang_dtype_t; /* bang_preprocess.py:13 */
typedef enum { BANG_DIST_L2 = 0, BANG_DIST_MIPS = 1 } bang_distfn_t;                  /* bang.h:26-30 DistFunc */
/* Storage / search modes = the reference's three forks (SURVEY.md §2.1 #6,#12,#15). */
typedef enum {
  BANG_MODE_BASE = 0,          /* BANG_Base: PQ distances, prefetched parent selection, exact re-rank   */
  BANG_MODE_INMEMORY = 1,      /* BANG_Inmemory: PQ distances, parent picked after the merge, re-rank   */
  BANG_MODE_EXACTDISTANCE = 2  /* BANG_Exactdistance: exact distances, single-hash filter, no re-rank   */
} bang_mode_t;
/* layout of the `dists` output of bang_b200_query */
typedef enum {
  BANG_DISTS_RANK_MAJOR = 0,   /* dists[j*Q + q]  — what BANG_Base returns (bang_search.cu:999,1297; SURVEY C-6) */
  BANG_DISTS_QUERY_MAJOR = 1   /* dists[q*k + j]  — same layout as ids                                          */
} bang_dists_layout_t;

enum {
  BANG_OK = 0, BANG_E_ARG = -1, BANG_E_IO = -2, BANG_E_FORMAT = -3, BANG_E_CUDA = -4, BANG_E_STATE = -5,
  BANG_E_NOMEM = -6, BANG_E_UNSUPPORTED = -7
};

typedef struct bang_b200_ctx* bang_handle_t;

/* BANGSearch<T>::BANGSearch() (bang_search.cu:70-75).  device < 0 = current device. */
int bang_b200_create(bang_handle_t* out, bang_dtype_t dtype, bang_mode_t mode, int device);
/* BANGSearch<T>::~BANGSearch() */
void bang_b200_destroy(bang_handle_t h);

/* BANGSearch<T>::bang_load (bang.h:53, bang_search.cu:139-362): reads <prefix>_pq_pivots.bin (4-section
 * layout), <prefix>_pq_compressed.bin, <prefix>_disk.bin, <prefix>_disk_metadata.bin.  In
 * BANG_MODE_EXACTDISTANCE the two PQ files are not required.  Returns BANG_E_IO / BANG_E_FORMAT where
 * the reference returns false. */
int bang_b200_load(bang_handle_t h, const char* indexfile_path_prefix);

/* The Inmemory / Exactdistance forks take explicit file names and compile-time N/D/MEDOID
 * (BANG_Inmemory/parANN.cu:79-93,146-221; parANN.h:38-158): old three-file PQ layout.  pq_* may be NULL
 * in BANG_MODE_EXACTDISTANCE.  R is fixed at 64 as in the reference. */
int bang_b200_load_files(bang_handle_t h, const char* pq_pivots_bin, const char* pq_compressed_bin,
                         const char* disk_bin, const char* chunk_offsets_bin, const char* centroid_bin,
                         uint64_t N, uint32_t D, uint64_t medoid);

/* Multi-GPU graph sharding (replaces BANG_Base's host-RAM graph, bang_search.cu:709-845):
 * call before load; this process keeps only rows with id % n_shards == shard.  After load every rank
 * exports its shard (IPC handle), imports its peers' and the traversal kernel reads remote rows with
 * P2P loads.  n_shards == 1 (default) = fully replicated index. */
int bang_b200_set_sharding(bang_handle_t h, int shard, int n_shards);
int bang_b200_export_shard(bang_handle_t h, void* ipc_handle_64B);
int bang_b200_import_shard(bang_handle_t h, int shard, const void* ipc_handle_64B);
/* The same exchange for rows allocated with the CUDA virtual-memory API (BANG_B200_SHARD_VMM=1 in the environment at
 * load time: cuMemCreate at the device's recommended granularity, mapped with matching alignment on both sides).  The
 * shard travels as a POSIX file descriptor, which the caller passes between the processes (SCM_RIGHTS over a
 * Unix-domain socket: bang_b200/sharding.py) and closes after the import. */
int bang_b200_export_shard_fd(bang_handle_t h, int* fd_out, uint64_t* bytes_out);
int bang_b200_import_shard_fd(bang_handle_t h, int shard, int fd, uint64_t bytes);

/* Device-resident load (no files): for indices that are built on the GPUs themselves and are too large for the
 * box's disk (SIFT1B-shape: a 388 GB `_disk.bin`).  Same HBM layout and search as bang_b200_load; honours
 * bang_b200_set_sharding (local row r of shard s holds node r * n_shards + s).  begin -> rows* -> codes* -> end.
 * pivots/centroid/chunk_offsets are HOST arrays in the file's order (float[256][D], float[D], u32[n_chunks+1]);
 * d_vectors (T[n][D]), d_adj (u32[n][64], unused slots 0xFFFFFFFF) and d_codes (u8[n][n_chunks]) are DEVICE arrays. */
int bang_b200_load_device_begin(bang_handle_t h, uint64_t N, uint32_t D, uint64_t medoid, uint32_t n_chunks,
                                const float* pivots, const float* centroid, const uint32_t* chunk_offsets);
int bang_b200_load_device_rows(bang_handle_t h, uint64_t first_local_row, uint64_t n_rows, const void* d_vectors,
                               const uint32_t* d_adj);
int bang_b200_load_device_codes(bang_handle_t h, uint64_t first_id, uint64_t n, const uint8_t* d_codes);
/* the same for n nodes with non-consecutive ids: d_ids is a DEVICE array u32[n], d_codes u8[n][n_chunks] in that order */
int bang_b200_load_device_codes_at(bang_handle_t h, const uint32_t* d_ids, uint64_t n, const uint8_t* d_codes);
int bang_b200_load_device_end(bang_handle_t h);

/* BANGSearch<T>::bang_set_searchparams (bang.h:60-62, bang_search.cu:562-567) */
int bang_b200_set_searchparams(bang_handle_t h, int recall, int worklist_length, bang_distfn_t dist);
/* BANGSearch<T>::bang_alloc (bang.h:55, bang_search.cu:367-423) */
int bang_b200_alloc(bang_handle_t h, int num_queries);
/* BANGSearch<T>::bang_init (bang.h:58, bang_search.cu:428-507) */
int bang_b200_init(bang_handle_t h, int num_queries);
/* BANGSearch<T>::bang_query (bang.h:75-78, bang_search.cu:570-1068).  queries: host T[Q][D] (D-1 per row
 * for MIPS); ids: host u64[Q*k] query-major, ascending exact distance; dists: host float[Q*k], squared
 * L2, layout per bang_b200_set_dists_layout (default rank-major like the reference). */
int bang_b200_query(bang_handle_t h, const void* queries, int num_queries, uint64_t* ids, float* dists);
/* BANGSearch<T>::bang_free (bang.h:80, bang_search.cu:510-548) */
int bang_b200_free(bang_handle_t h);
/* BANGSearch<T>::bang_unload (bang.h:82, bang_search.cu:551-559) */
int bang_b200_unload(bang_handle_t h);

int bang_b200_set_dists_layout(bang_handle_t h, bang_dists_layout_t layout);

/* Device-resident variant of bang_query: queries, ids (u64[Q*k]) and dists (float[Q*k], query-major)
 * are DEVICE pointers; the search is enqueued on `cuda_stream` (a cudaStream_t, may be NULL) and not
 * synchronised.  Used to time the kernel with inputs already in HBM. */
int bang_b200_query_device(bang_handle_t h, const void* d_queries, int num_queries, uint64_t* d_ids,
                           float* d_dists, void* cuda_stream);

/* Stage (1) on its own — populate_pqDist_par (bang_search.cu:1083-1130): host queries in, host
 * float[Q][n_chunks][256] out. */
int bang_b200_pq_table(bang_handle_t h, const void* queries, int num_queries, float* tables);

/* Index facts after load. */
typedef struct {
  uint64_t N, medoid, entry_len;
  uint32_t D, R, n_chunks;
  int32_t dtype, mode;
  uint64_t device_bytes;  /* HBM held by the index on this device */
  uint32_t row_stride;    /* bytes per graph row in HBM: 64 neighbour ids [+ 512 B of precomputed visited-filter slots] + vector */
  uint32_t slot_block;    /* 1: the rows carry the slot block (opt-in for PQ modes with 32 uniform chunks: BANG_B200_PREHASH=1) */
} bang_b200_info_t;
int bang_b200_info(bang_handle_t h, bang_b200_info_t* out);

/* The two visited-filter slots of a point id, computed on the host exactly as the load kernels write them into the rows'
 * slot block and as the kernels evaluate them per hop: pos[i] = hashFn1_d / hashFn2_d (bang_search.cu:1168-1189) of the
 * id, in [0, 399887); words[i] = (pos / 255 * 8) | (pos % 255) << 24, i.e. the byte offset of the slot's 8-byte block in
 * a query's sparse filter and the slot's offset inside the block.  No device is touched. */
void bang_b200_filter_slots(uint32_t id, uint32_t pos[2], uint32_t words[2]);

/* Per-query counters of the last query call (host arrays of length Q, any may be NULL):
 * hops = expanded nodes incl. medoid, sum_deg = sum of their degrees, n_cand = candidates that passed
 * the visited filter.  These feed the roofline's algorithmic-bytes figure. */
int bang_b200_last_stats(bang_handle_t h, uint32_t* hops, uint32_t* sum_deg, uint32_t* n_cand);
/* Timing of the last query call: device ms of the fused kernel (CUDA events on its stream), number of
 * kernel launches, H2D and D2H bytes. */
typedef struct {
  float kernel_ms;
  uint32_t launches;
  uint64_t h2d_bytes, d2h_bytes;
  uint32_t grid, block, smem_bytes, ctas_per_sm;
} bang_b200_timing_t;
int bang_b200_last_timing(bang_handle_t h, bang_b200_timing_t* out);

const char* bang_b200_last_error(void);

/* ---- data preparation on the GPU (csrc/prep_kernels.cu) -------------------------------------------------------
 * The reference takes its ground truth and its PQ files from DiskANN's tools (`compute_groundtruth`,
 * `build_disk_index`; README.md:46-58).  These entry points produce the same arrays with hand-written kernels, for
 * indices built on the box.  All pointers prefixed d_ are DEVICE pointers; `cuda_stream` is a cudaStream_t or NULL;
 * the calls return after the work is done.  Errors: bang_b200_prep_last_error(). */
/* Exact k nearest neighbours (squared L2) of nq queries among n base rows, k <= 128: d_ids u32[nq][k] (row index +
 * id_offset), d_dists float[nq][k], both ordered by (distance, id) — the two halves of the truthset file
 * (test_driver.cpp:238-272).  u8/i8 distances are exact integers. */
int bang_b200_bruteforce_gt(bang_dtype_t dtype, const void* d_base, uint64_t n, uint32_t D, const void* d_queries, uint32_t num_queries,
                            uint32_t k, uint64_t id_offset, uint32_t* d_ids, float* d_dists, void* cuda_stream);
/* PQ training: centroid = column means of the base (host float[D] out); pivots = 256 k-means centres per chunk of the
 * centred data (host float[256][D] out, the layout of `_pq_pivots.bin`), Lloyd iterations on at most max_train rows
 * spread evenly over the base (0 = all rows).  chunk_offsets: host u32[n_chunks+1], chunks of 1..32 dimensions. */
int bang_b200_pq_train(bang_dtype_t dtype, const void* d_base, uint64_t n, uint32_t D, const uint32_t* chunk_offsets, uint32_t n_chunks,
                       uint32_t iters, uint64_t max_train, uint64_t seed, float* pivots, float* centroid, void* cuda_stream);
/* PQ encoding: d_codes u8[n][n_chunks] = index of the closest pivot of every chunk (first minimum) — the payload of
 * `_pq_compressed.bin`.  pivots / centroid / chunk_offsets are host arrays as above. */
int bang_b200_pq_encode(bang_dtype_t dtype, const void* d_base, uint64_t n, uint32_t D, const float* pivots, const float* centroid,
                        const uint32_t* chunk_offsets, uint32_t n_chunks, uint8_t* d_codes, void* cuda_stream);
const char* bang_b200_prep_last_error(void);

/* The reference's own C API (bang.h:89-101, `#if 0` there): one process-wide uint8 instance. */
int bang_load_c(char* indexfile_path_prefix);
void bang_set_searchparams_c(int recall, int worklist_length, int nDistFunc);
void bang_query_c(uint8_t* query_array, int num_queries, unsigned long* nearestNeighbours, float* nearestNeighbours_dist);
void bang_unload_c(void);

#ifdef __cplusplus
}
#endif
#endif /* BANG_B200_H_ */
