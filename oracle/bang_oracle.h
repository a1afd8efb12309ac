/* bang_oracle.h — CPU restatement of BANG's batched greedy Vamana search (TEST INFRASTRUCTURE).
 *
 * This is the parity checker for the CUDA path, not a product path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may call it.  See oracle/bang_oracle.c for the
 * reference file:line each function follows.
 */
#ifndef BANG_ORACLE_H_
#define BANG_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { BANG_ORACLE_MODE_BASE = 0, BANG_ORACLE_MODE_INMEMORY = 1, BANG_ORACLE_MODE_EXACT = 2 };
enum { BANG_ORACLE_DT_INT8 = 0, BANG_ORACLE_DT_UINT8 = 1, BANG_ORACLE_DT_FLOAT = 2 };
/* floating-point summation order of exact L2 distances: the reference kernels' own order, or the
 * order the sm_100a kernels use (identical results for u8/i8 while D*255^2 < 2^24). */
enum { BANG_ORACLE_ORDER_REF = 0, BANG_ORACLE_ORDER_GPU = 1 };

#define BANG_ORACLE_BF_ENTRIES 399887u
#define BANG_ORACLE_NO_ID 0xFFFFFFFFull

typedef struct {
  uint32_t N, D, R, n_chunks;
  int32_t dtype;
  uint64_t medoid;
  uint64_t entry_len;          /* D*sizeof(T) + 4 + 4R */
  const uint8_t* disk;         /* raw `_disk.bin` */
  const uint8_t* codes;        /* uint8[N][n_chunks]         (NULL in exact mode) */
  const float* pivots;         /* float[256][D], file layout (NULL in exact mode) */
  const float* centroid;       /* float[D] */
  const uint32_t* chunk_offsets; /* u32[n_chunks+1] */
} bang_oracle_index;

typedef struct {
  uint32_t* hops;      /* [Q] expanded nodes incl. medoid (|Cand|)                    or NULL */
  uint32_t* sum_deg;   /* [Q] sum of degrees of the expanded nodes                    or NULL */
  uint32_t* n_cand;    /* [Q] candidates that passed the visited filter               or NULL */
  uint32_t* trace;     /* [Q][trace_len] expanded node ids in order, padded with ~0u  or NULL */
  uint32_t trace_len;
} bang_oracle_stats;

uint32_t bang_oracle_hash1(uint32_t x);
uint32_t bang_oracle_hash2(uint32_t x);

/* tbl[n_chunks][256] for one query (query points at D elements of the index dtype) */
void bang_oracle_pq_table(const bang_oracle_index* ix, const void* query, float* tbl);

/* asymmetric PQ distance of node `id` from a table */
float bang_oracle_pq_dist(const bang_oracle_index* ix, const float* tbl, uint32_t id);

/* exact squared L2 between a node and a query; order = BANG_ORACLE_ORDER_*; kind 0 = re-rank kernel,
 * 1 = Exactdistance neighbour kernel (only matters for ORDER_REF on float data) */
float bang_oracle_l2(const bang_oracle_index* ix, uint32_t id, const void* query, int order, int kind);

/* Full search.  ids/dists are query-major [Q][k]; slots beyond the available results hold
 * BANG_ORACLE_NO_ID / FLT_MAX.  Returns 0 on success. */
int bang_oracle_search(const bang_oracle_index* ix, int mode, const void* queries, uint32_t Q, uint32_t k,
                       uint32_t L, int order, int nthreads, uint64_t* ids, float* dists,
                       const bang_oracle_stats* stats);

/* blocked multi-threaded brute force exact kNN (host-core baseline), ties by id */
int bang_oracle_bruteforce(const bang_oracle_index* ix, const void* queries, uint32_t Q, uint32_t k, int nthreads,
                           uint32_t* ids, float* dists);

#ifdef __cplusplus
}
#endif
#endif
