// search_kernel.cuh — the fused, persistent, per-query-block traversal kernel for sm_100a.
//
// One CTA (128 threads) owns one query at a time and walks the whole greedy search for it without
// leaving the SM: PQ table build (stage 1) -> { adjacency fetch, visited filter (stage 4a), PQ/exact
// distances (stage 3), (dist,id) sort + worklist merge (stage 4b), parent selection (stage 2) }* ->
// exact re-rank + top-k (stage 5).  CTAs are persistent: the grid is sized to the number of resident
// CTAs and each CTA pulls query indices from a global counter.  What the reference does with ~5 kernel
// launches, 2 memsets, up to 5 PCIe copies and 3 stream syncs per hop (bang_search.cu:701-958) is one
// launch here; the LUT, worklist, candidate log and neighbour lists never leave shared memory.
//
// Semantics are the reference's with the deterministic choices documented in
// oracle/bang_oracle.c; the oracle is bit-exact with this kernel (ORDER_GPU).
//   stage 1  populate_pqDist_par            BANG_Base/bang_search.cu:1083-1130
//   stage 2  compute_parent1 / 2            bang_search.cu:1464-1521 / 1384-1459 (Inmemory: parANN.cu:1399-1418)
//   stage 3  compute_neighborDist_par       bang_search.cu:1201-1241 (exact: BANG_Exactdistance/parANN.cu:1139-1179)
//   stage 4  neighbor_filtering_new + sort + merge   bang_search.cu:1140-1189, 1533-1585, 1605-1715
//   stage 5  compute_L2Dist + compute_NearestNeighbours  bang_search.cu:1254-1368
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bang {

constexpr int kThreads = 128;           // threads per query CTA
constexpr int kMaxR = 64;               // MAX_R, bang_search.cu:35
constexpr int kListCap = kMaxR + 8;     // medoid + R neighbours, padded
constexpr uint32_t kBfEntries = 399887u;  // BF_ENTRIES, bang_search.cu:48
constexpr uint32_t kBloomWords = 12512u;  // ceil(399887/32)=12497, padded to a multiple of 32 words
constexpr uint32_t kNoNbr = 0xFFFFFFFFu;  // padding id in the HBM adjacency rows
constexpr int kAdjBytes = kMaxR * 4;    // 256 B adjacency block at the head of each HBM row
constexpr int kMaxShards = 8;

enum Mode : int { kBase = 0, kInmemory = 1, kExact = 2 };

struct SearchArgs {
  // ---- index (HBM layout) ----
  const uint8_t* rows[kMaxShards];  // shard s holds ids with id % n_shards == s at local row id / n_shards
  uint32_t n_shards;
  uint32_t row_stride;      // bytes, multiple of 32: [ R x u32 adjacency | vector padded to 16 B ]
  const uint8_t* codes;     // [N][code_stride], bytes permuted per 32-chunk group (see repack_codes)
  uint32_t code_stride;
  uint32_t n_chunks;
  const float* pivT;        // [D][256]   pivots transposed as the reference does at load (:281-285)
  const float* centroid;    // [D]
  const uint32_t* chunk_off;  // [n_chunks+1]
  uint32_t D;               // dims of the index
  uint32_t vec_units;       // 16-byte units per vector (D*sizeof(T) rounded up / 16)
  uint32_t medoid;
  // ---- search ----
  uint32_t L, k, Q;
  uint32_t q_dim;           // elements per query row (D, or D-1 for MIPS)
  uint32_t max_iter;        // L+50 (Base) / L+120 (Inmemory) / 4L+20 (Exact)
  uint32_t cand_cap;        // max_iter + 1
  const void* queries;      // device T[Q][q_dim]
  uint64_t* out_ids;        // device [Q][k]
  float* out_dists;         // device [Q][k] (query-major)
  uint32_t* bloom;          // device [gridDim.x][kBloomWords]
  uint32_t* counter;        // device work counter (zeroed before launch)
  uint32_t* st_hops;        // device [Q] or null
  uint32_t* st_sumdeg;
  uint32_t* st_npass;
  // optional dump of the expanded-node log (used by the index builder): ids [Q][dump_stride], count [Q]
  uint32_t* dump_ids;
  uint32_t* dump_n;
  uint32_t dump_stride;
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash1(uint32_t x) {  // hashFn1_d, bang_search.cu:1168-1178
  uint64_t h = 0xcbf29ce4ull;
  h = (h ^ (x & 0xff)) * 0x01000193ull;
  h = (h ^ ((x >> 8) & 0xff)) * 0x01000193ull;
  h = (h ^ ((x >> 16) & 0xff)) * 0x01000193ull;
  h = (h ^ ((x >> 24) & 0xff)) * 0x01000193ull;
  return (uint32_t)(h % kBfEntries);
}
__device__ __forceinline__ uint32_t hash2(uint32_t x) {  // hashFn2_d, bang_search.cu:1179-1189
  uint64_t h = 0x84222325ull;
  h = (h ^ (x & 0xff)) * 0x1B3ull;
  h = (h ^ ((x >> 8) & 0xff)) * 0x1B3ull;
  h = (h ^ ((x >> 16) & 0xff)) * 0x1B3ull;
  h = (h ^ ((x >> 24) & 0xff)) * 0x1B3ull;
  return (uint32_t)(h % kBfEntries);
}

__device__ __forceinline__ bool key_less(float da, uint32_t ia, float db, uint32_t ib) {
  return da < db || (da == db && ia < ib);
}

__device__ __forceinline__ float tree8(float v) {
  // 8-lane tree ((a0+a1)+(a2+a3))+((a4+a5)+(a6+a7)); every lane ends with the full sum
  v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 4));
  return v;
}

template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int kPerUnit = 4;
  __device__ static __forceinline__ void unpack(const uint4& u, float* f) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
};
template <> struct Elem<uint8_t> {
  static constexpr int kPerUnit = 16;
  __device__ static __forceinline__ void unpack(const uint4& u, float* f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[4 * i + 0] = (float)(w[i] & 0xff); f[4 * i + 1] = (float)((w[i] >> 8) & 0xff);
      f[4 * i + 2] = (float)((w[i] >> 16) & 0xff); f[4 * i + 3] = (float)(w[i] >> 24);
    }
  }
};
template <> struct Elem<int8_t> {
  static constexpr int kPerUnit = 16;
  __device__ static __forceinline__ void unpack(const uint4& u, float* f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[4 * i + 0] = (float)(int8_t)(w[i] & 0xff); f[4 * i + 1] = (float)(int8_t)((w[i] >> 8) & 0xff);
      f[4 * i + 2] = (float)(int8_t)((w[i] >> 16) & 0xff); f[4 * i + 3] = (float)(int8_t)(w[i] >> 24);
    }
  }
};

__device__ __forceinline__ const uint8_t* row_ptr(const SearchArgs& a, uint32_t id) {
  if (a.n_shards == 1) return a.rows[0] + (size_t)id * a.row_stride;
  const uint32_t s = id % a.n_shards;
  return a.rows[s] + (size_t)(id / a.n_shards) * a.row_stride;  // local HBM or a peer mapping over NVLink
}

__device__ __forceinline__ uint4 ld_nc_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_nc_u32(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// Exact squared L2 between one HBM row vector and the query (fp32 copy in shared memory).
// 8 lanes per row: lane t takes the 16-byte units t, t+8, ...; fmaf in element order; 8-lane tree.
// Rows and the query are zero-padded to whole units, which leaves the sum unchanged.
template <typename T>
__device__ __forceinline__ float l2_row_8lane(const uint8_t* vec, const float* q_f, uint32_t units, uint32_t t) {
  constexpr int E = Elem<T>::kPerUnit;
  float acc = 0.0f;
  uint32_t u = t;
  // two independent 16-byte loads in flight per lane
  for (; u + 8 < units; u += 16) {
    const uint4 r0 = ld_nc_u4(vec + (size_t)u * 16);
    const uint4 r1 = ld_nc_u4(vec + (size_t)(u + 8) * 16);
    float f[E];
    Elem<T>::unpack(r0, f);
#pragma unroll
    for (int e = 0; e < E; ++e) { const float d = __fsub_rn(f[e], q_f[u * E + e]); acc = __fmaf_rn(d, d, acc); }
    Elem<T>::unpack(r1, f);
#pragma unroll
    for (int e = 0; e < E; ++e) { const float d = __fsub_rn(f[e], q_f[(u + 8) * E + e]); acc = __fmaf_rn(d, d, acc); }
  }
  if (u < units) {
    const uint4 r0 = ld_nc_u4(vec + (size_t)u * 16);
    float f[E];
    Elem<T>::unpack(r0, f);
#pragma unroll
    for (int e = 0; e < E; ++e) { const float d = __fsub_rn(f[e], q_f[u * E + e]); acc = __fmaf_rn(d, d, acc); }
  }
  return tree8(acc);
}

// ------------------------------------------------------------------------------------------------
// shared-memory state of one query
// ------------------------------------------------------------------------------------------------
struct QState {
  float* q_f;        // [vec_units * E] query as fp32, zero padded
  float* lut;        // [n_chunks][256]            (PQ modes)
  float* w_d0;       // worklist, double buffered  [2][w_stride]
  uint32_t* w_id0;
  uint8_t* w_v0;
  uint32_t w_stride;  // elements between the two buffers
  __device__ __forceinline__ float* wd(uint32_t buf) const { return w_d0 + buf * w_stride; }
  __device__ __forceinline__ uint32_t* wid(uint32_t buf) const { return w_id0 + buf * w_stride; }
  __device__ __forceinline__ uint8_t* wv(uint32_t buf) const { return w_v0 + buf * w_stride; }
  uint32_t* lst;     // [kListCap] raw candidate list (medoid +) adjacency
  uint32_t* n_id;    // [kListCap] filtered, list order
  float* n_d;
  uint32_t* s_id;    // [kListCap] sorted by (dist, id)
  float* s_d;
  uint32_t* cand_id; // [cand_cap] expanded-node log
  float* cand_d;     // exact distances of the log entries
  uint32_t* scal;    // scalars, see below
};
// scalar slots in QState::scal
enum { S_NLIST = 0, S_NSIZE, S_WSIZE, S_CUR, S_PARENT, S_HAVE, S_NCAND, S_MARK, S_WMASK0, S_WMASK1, S_WMASK2, S_WMASK3,
       S_QUERY, S_SUMDEG, S_NPASS, S_FOUND, S_COUNT };

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename T>
__host__ __device__ inline size_t smem_bytes(int mode, uint32_t n_chunks, uint32_t vec_units, uint32_t L, uint32_t cand_cap) {
  size_t b = 0;
  b += align_up((size_t)vec_units * Elem<T>::kPerUnit * 4, 16);
  if (mode != kExact) b += (size_t)n_chunks * 256 * 4;
  b += 2 * (align_up(L, 16) * 4 * 2 + align_up(L, 16));
  b += (size_t)kListCap * 4 * 5;
  b += align_up((size_t)cand_cap * 4, 16) * 2;
  b += S_COUNT * 4 + 16;
  return align_up(b, 16);
}

template <typename T>
__device__ __forceinline__ void carve(QState& s, uint8_t* base, int mode, const SearchArgs& a) {
  size_t o = 0;
  s.q_f = (float*)(base + o); o += align_up((size_t)a.vec_units * Elem<T>::kPerUnit * 4, 16);
  s.lut = (float*)(base + o); if (mode != kExact) o += (size_t)a.n_chunks * 256 * 4;
  s.w_stride = (uint32_t)align_up(a.L, 16);
  s.w_d0 = (float*)(base + o); o += (size_t)s.w_stride * 4 * 2;
  s.w_id0 = (uint32_t*)(base + o); o += (size_t)s.w_stride * 4 * 2;
  s.w_v0 = (uint8_t*)(base + o); o += (size_t)s.w_stride * 2;
  s.lst = (uint32_t*)(base + o); o += kListCap * 4;
  s.n_id = (uint32_t*)(base + o); o += kListCap * 4;
  s.n_d = (float*)(base + o); o += kListCap * 4;
  s.s_id = (uint32_t*)(base + o); o += kListCap * 4;
  s.s_d = (float*)(base + o); o += kListCap * 4;
  s.cand_id = (uint32_t*)(base + o); o += align_up((size_t)a.cand_cap * 4, 16);
  s.cand_d = (float*)(base + o); o += align_up((size_t)a.cand_cap * 4, 16);
  s.scal = (uint32_t*)(base + o);
}

// ------------------------------------------------------------------------------------------------
// stage 1: per-query PQ distance table into shared memory (or global for the standalone kernel)
// tbl[c][k] = sum_{j in chunk c} (pivT[j][k] - (q[j] - centroid[j]))^2, j ascending, fmaf
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void build_pq_table(const SearchArgs& a, const float* q_f, float* tbl /*[m][256]*/) {
  const uint32_t tid = threadIdx.x;
  // 128 threads x 2 consecutive centres each (float2, coalesced over pivT rows)
  for (uint32_t c = 0; c < a.n_chunks; ++c) {
    const uint32_t j0 = a.chunk_off[c], j1 = a.chunk_off[c + 1];
    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 8
    for (uint32_t j = j0; j < j1; ++j) {
      const float qc = __fsub_rn(q_f[j], __ldg(a.centroid + j));
      const float2 p = __ldg(reinterpret_cast<const float2*>(a.pivT + (size_t)j * 256) + tid);
      const float d0 = __fsub_rn(p.x, qc), d1 = __fsub_rn(p.y, qc);
      acc0 = __fmaf_rn(d0, d0, acc0);
      acc1 = __fmaf_rn(d1, d1, acc1);
    }
    reinterpret_cast<float2*>(tbl + (size_t)c * 256)[tid] = make_float2(acc0, acc1);
  }
}

template <typename T>
__device__ __forceinline__ void load_query(const SearchArgs& a, uint32_t q, float* q_f) {
  const T* src = reinterpret_cast<const T*>(a.queries) + (size_t)q * a.q_dim;
  const uint32_t n = a.vec_units * Elem<T>::kPerUnit;
  for (uint32_t i = threadIdx.x; i < n; i += kThreads) q_f[i] = i < a.q_dim ? (float)src[i] : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// stage 4a: visited filter (bloom filter in L2-resident global memory, one region per resident CTA).
// Sequential-in-list-order semantics: each accepted id's bits are visible to the ids after it.
// Executed by warp 0; rounds of 32 ids.  Fast path: test, atomicOr, and detect through the atomics'
// return values whether two lanes of the round touched the same bit; only then (rare) the accept
// decisions are re-derived sequentially (the final bit state is the union either way).
// ------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void filter_list(const QState& s, uint32_t* bloom, uint32_t n_list) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t n_out = 0;
  for (uint32_t base = 0; base < n_list; base += 32) {
    const uint32_t i = base + lane;
    const bool valid = i < n_list;
    const uint32_t id = valid ? s.lst[i] : 0u;
    const uint32_t p1 = hash1(id);
    const uint32_t p2 = (MODE == kExact) ? p1 : hash2(id);
    const uint32_t w1 = p1 >> 5, b1 = 1u << (p1 & 31), w2 = p2 >> 5, b2 = 1u << (p2 & 31);
    bool s1 = false, s2 = false;
    if (valid) {
      s1 = (__ldcg(bloom + w1) & b1) != 0;
      s2 = (MODE == kExact) ? s1 : ((__ldcg(bloom + w2) & b2) != 0);
    }
    bool accept = valid && !(s1 && s2);
    bool conflict = false;
    __syncwarp();  // every lane's test precedes every lane's set
    if (accept) {
      const uint32_t o1 = atomicOr(bloom + w1, b1);
      conflict = ((o1 & b1) != 0) && !s1;
      if (MODE != kExact && p2 != p1) {
        const uint32_t o2 = atomicOr(bloom + w2, b2);
        conflict = conflict || (((o2 & b2) != 0) && !s2);
      }
    }
    const uint32_t amask = __ballot_sync(0xffffffffu, accept);
    if (__any_sync(0xffffffffu, conflict)) {
      // re-derive accept decisions in list order
      uint32_t todo = amask;
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const bool acc_src = __shfl_sync(0xffffffffu, (int)(!(s1 && s2)), src) != 0;
        const uint32_t q1 = __shfl_sync(0xffffffffu, p1, src);
        const uint32_t q2 = __shfl_sync(0xffffffffu, p2, src);
        if (acc_src && (int)lane > src) {
          s1 = s1 || p1 == q1 || p1 == q2;
          s2 = s2 || p2 == q1 || p2 == q2;
        }
      }
      accept = valid && ((amask >> lane) & 1u) && !(s1 && s2);
    }
    const uint32_t fmask = __ballot_sync(0xffffffffu, accept);
    if (accept) s.n_id[n_out + __popc(fmask & ((1u << lane) - 1u))] = id;
    n_out += __popc(fmask);
  }
  if (lane == 0) s.scal[S_NSIZE] = n_out;
}

// ------------------------------------------------------------------------------------------------
// stage 3 (PQ): dist[i] = sum_c lut[c][code[id_i][c]] — 8 lanes per candidate; lane t owns chunks
// t, t+8, ... ascending (the reference's split), partials combined by the 8-lane tree.  The HBM code
// rows are permuted at load so that lane t's chunks 32g+t, 32g+8+t, 32g+16+t, 32g+24+t are one aligned
// 32-bit word: one 32-byte sector per candidate per 32 chunks, fully used.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pq_distances(const SearchArgs& a, const QState& s, uint32_t n) {
  const uint32_t tid = threadIdx.x, t = tid & 7, slot = tid >> 3;  // 16 candidates per pass
  const uint32_t groups = (a.n_chunks + 31) >> 5;
  constexpr int kPass = (kListCap + 15) / 16;  // 5
  if (groups == 1) {
    uint32_t w[kPass];
#pragma unroll
    for (int p = 0; p < kPass; ++p) {
      const uint32_t i = p * 16 + slot;
      w[p] = 0;
      if (i < n) w[p] = ld_nc_u32(a.codes + (size_t)s.n_id[i] * a.code_stride + 4 * t);
    }
#pragma unroll
    for (int p = 0; p < kPass; ++p) {
      const uint32_t i = p * 16 + slot;
      if (p * 16 < (int)n) {  // uniform per pass
        float sum = 0.0f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t c = t + 8 * b;
          if (c < a.n_chunks) sum = __fadd_rn(sum, s.lut[c * 256 + ((w[p] >> (8 * b)) & 0xff)]);
        }
        sum = tree8(sum);
        if (t == 0 && i < n) s.n_d[i] = sum;
      }
    }
  } else {
    for (uint32_t p0 = 0; p0 < n; p0 += 16) {
      const uint32_t i = p0 + slot;
      float sum = 0.0f;
      const uint8_t* row = a.codes + (size_t)(i < n ? s.n_id[i] : 0u) * a.code_stride + 4 * t;
      for (uint32_t g = 0; g < groups; g += 2) {
        const uint32_t wa = (i < n) ? ld_nc_u32(row + g * 32) : 0u;
        const uint32_t wb = (i < n && g + 1 < groups) ? ld_nc_u32(row + (g + 1) * 32) : 0u;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t c = g * 32 + t + 8 * b;
          if (c < a.n_chunks) sum = __fadd_rn(sum, s.lut[c * 256 + ((wa >> (8 * b)) & 0xff)]);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t c = (g + 1) * 32 + t + 8 * b;
          if (c < a.n_chunks) sum = __fadd_rn(sum, s.lut[c * 256 + ((wb >> (8 * b)) & 0xff)]);
        }
      }
      sum = tree8(sum);
      if (t == 0 && i < n) s.n_d[i] = sum;
    }
  }
}

// stage 3 (exact): full-precision rows, 8 lanes per candidate, whole 128-byte lines per request
template <typename T>
__device__ __forceinline__ void exact_distances(const SearchArgs& a, const QState& s, uint32_t n) {
  const uint32_t tid = threadIdx.x, t = tid & 7, slot = tid >> 3;
  for (uint32_t p0 = 0; p0 < n; p0 += 16) {
    const uint32_t i = p0 + slot;
    const uint32_t id = i < n ? s.n_id[i] : a.medoid;
    const float d = l2_row_8lane<T>(row_ptr(a, id) + kAdjBytes, s.q_f, a.vec_units, t);
    if (t == 0 && i < n) s.n_d[i] = d;
  }
}

// rank sort of the (<= 65) filtered neighbours by (dist, id) -> s_id / s_d
__device__ __forceinline__ void sort_neighbours(const QState& s, uint32_t n) {
  const uint32_t tid = threadIdx.x;
  if (tid < n) {
    const float d = s.n_d[tid];
    const uint32_t id = s.n_id[tid];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; ++j) rank += key_less(s.n_d[j], s.n_id[j], d, id) ? 1u : 0u;
    s.s_d[rank] = d;
    s.s_id[rank] = id;
  }
}

// stage 4b: merge the sorted neighbours into the worklist (compute_BestLSets_par_merge,
// bang_search.cu:1636-1709).  Block-uniform control flow; returns with W in buffer s.scal[S_CUR].
__device__ __forceinline__ void merge_worklist(const SearchArgs& a, const QState& s, uint32_t n, bool first) {
  const uint32_t tid = threadIdx.x;
  const uint32_t L = a.L;
  if (n == 0) return;  // uniform
  uint32_t cur = s.scal[S_CUR];
  if (first) {
    const uint32_t nb = min(n, L);
    for (uint32_t i = tid; i < nb; i += kThreads) {
      s.wid(cur)[i] = s.s_id[i];
      s.wd(cur)[i] = s.s_d[i];
      s.wv(cur)[i] = (s.s_id[i] == a.medoid) ? 1 : 0;
    }
    __syncthreads();
    if (tid == 0) s.scal[S_WSIZE] = nb;
    __syncthreads();
    return;
  }
  const uint32_t ws = s.scal[S_WSIZE];
  const float maxd = s.wd(cur)[ws - 1];
  const uint32_t lim = min(L, n);
  // sorted ascending => the leading run with d < maxd is exactly the set with d < maxd
  uint32_t nb = __syncthreads_count(tid < lim && s.s_d[tid] < maxd);
  nb = max(nb, min(L - ws, n));
  if (nb == 0) return;  // uniform
  const uint32_t newsize = min(ws + nb, L);
  const uint32_t nxt = cur ^ 1u;
  if (tid < nb) {  // new entry: position = lower_bound(W, d) + index  (new before old on ties)
    const float d = s.s_d[tid];
    uint32_t lo = 0, hi = ws;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (d <= s.wd(cur)[mid]) hi = mid; else lo = mid + 1;
    }
    const uint32_t pos = lo + tid;
    if (pos < newsize) {
      s.wd(nxt)[pos] = d;
      s.wid(nxt)[pos] = s.s_id[tid];
      s.wv(nxt)[pos] = 0;
    }
  }
  for (uint32_t j = tid; j < ws; j += kThreads) {  // old entry: position = upper_bound(new, d) + index
    const float d = s.wd(cur)[j];
    uint32_t lo = 0, hi = nb;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (d >= s.s_d[mid]) lo = mid + 1; else hi = mid;
    }
    const uint32_t pos = lo + j;
    if (pos < newsize) {
      s.wd(nxt)[pos] = d;
      s.wid(nxt)[pos] = s.wid(cur)[j];
      s.wv(nxt)[pos] = s.wv(cur)[j];
    }
  }
  __syncthreads();
  if (tid == 0) { s.scal[S_WSIZE] = newsize; s.scal[S_CUR] = nxt; }
  __syncthreads();
}

// index of the first unvisited worklist entry, or 0xFFFFFFFF (all threads get the value)
__device__ __forceinline__ uint32_t first_unvisited(const QState& s) {
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t ws = s.scal[S_WSIZE], cur = s.scal[S_CUR];
  for (uint32_t base = 0; base < ws; base += kThreads) {
    const uint32_t j = base + tid;
    const bool un = j < ws && s.wv(cur)[j] == 0;
    const uint32_t m = __ballot_sync(0xffffffffu, un);
    if (lane == 0) s.scal[S_WMASK0 + warp] = m;
    __syncthreads();
    uint32_t found = 0xFFFFFFFFu;
#pragma unroll
    for (int w = kThreads / 32 - 1; w >= 0; --w) {
      const uint32_t mw = s.scal[S_WMASK0 + w];
      if (mw) found = base + w * 32 + (__ffs(mw) - 1);
    }
    __syncthreads();
    if (found != 0xFFFFFFFFu) return found;
  }
  return 0xFFFFFFFFu;
}

// ------------------------------------------------------------------------------------------------
// expansion of one node: adjacency fetch (+ the node's own exact distance for the re-rank, PQ modes),
// filter, distances, sort.  On return s_id/s_d hold the sorted new neighbours, scal[S_NSIZE] their count.
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE>
__device__ __forceinline__ void expand(const SearchArgs& a, const QState& s, uint32_t* bloom, uint32_t parent,
                                       bool with_medoid, uint32_t cand_slot) {
  const uint32_t tid = threadIdx.x;
  const uint8_t* row = row_ptr(a, parent);
  const uint32_t off = with_medoid ? 1u : 0u;
  uint32_t nb = kNoNbr;
  if (tid < kMaxR) nb = ld_nc_u32(row + 4 * tid);
  if (MODE != kExact && tid >= 96) {
    // warp 3 (its four 8-lane groups read the same addresses, which coalesce into one request): exact
    // distance of the expanded node itself (feeds stage 5; the vector sits right behind the adjacency
    // block in the same HBM row, so it rides the same fetch)
    const float d = l2_row_8lane<T>(row + kAdjBytes, s.q_f, a.vec_units, tid & 7);
    if (tid == 96) s.cand_d[cand_slot] = d;
  }
  if (tid < kMaxR) s.lst[off + tid] = nb;
  if (with_medoid && tid == 0) s.lst[0] = a.medoid;
  const uint32_t deg = __syncthreads_count(tid < kMaxR && nb != kNoNbr);
  const uint32_t n_list = off + deg;
  if (tid < 32) filter_list<MODE>(s, bloom, n_list);
  __syncthreads();
  const uint32_t n = s.scal[S_NSIZE];
  if (tid == 0) { s.scal[S_SUMDEG] += deg; s.scal[S_NPASS] += n; }
  if (MODE == kExact) exact_distances<T>(a, s, n);
  else pq_distances(a, s, n);
  __syncthreads();
  sort_neighbours(s, n);
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// stage 5: top-k of the candidate log by (exact distance, id)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void write_topk(const SearchArgs& a, uint32_t q, const uint32_t* ids, const float* d,
                                           uint32_t n) {
  const uint32_t tid = threadIdx.x;
  for (uint32_t i = tid; i < n; i += kThreads) {
    const float di = d[i];
    const uint32_t idi = ids[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; ++j) rank += key_less(d[j], ids[j], di, idi) ? 1u : 0u;
    if (rank < a.k) {
      a.out_ids[(size_t)q * a.k + rank] = idi;
      a.out_dists[(size_t)q * a.k + rank] = di;
    }
  }
  for (uint32_t r = n + tid; r < a.k; r += kThreads) {
    a.out_ids[(size_t)q * a.k + r] = 0xFFFFFFFFull;
    a.out_dists[(size_t)q * a.k + r] = 3.402823466e+38f;
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads) bang_search_kernel(const SearchArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  QState s;
  carve<T>(s, smem_raw, MODE, a);
  const uint32_t tid = threadIdx.x;
  uint32_t* bloom = a.bloom + (size_t)blockIdx.x * kBloomWords;

  for (;;) {
    if (tid == 0) s.scal[S_QUERY] = atomicAdd(a.counter, 1u);
    __syncthreads();
    const uint32_t q = s.scal[S_QUERY];
    if (q >= a.Q) break;

    // ---- per-query setup: query -> smem, bloom filter cleared, PQ table built in place ----
    load_query<T>(a, q, s.q_f);
    {
      uint4* b4 = reinterpret_cast<uint4*>(bloom);
      for (uint32_t i = tid; i < kBloomWords / 4; i += kThreads) b4[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
      s.scal[S_WSIZE] = 0; s.scal[S_CUR] = 0; s.scal[S_NSIZE] = 0; s.scal[S_NCAND] = 1;
      s.scal[S_SUMDEG] = 0; s.scal[S_NPASS] = 0; s.scal[S_MARK] = 0x01010101u;
      s.cand_id[0] = a.medoid;  // bang_init: the medoid is every query's first candidate (:455-462)
    }
    __syncthreads();
    if (MODE != kExact) {
      build_pq_table(a, s.q_f, s.lut);
      __syncthreads();
    }

    if (MODE == kBase) {
      // ---- BANG_Base order: A.1 seed, then { merge(prev) ; mark ; expand(parent) ; compute_parent2 } ----
      expand<T, MODE>(a, s, bloom, a.medoid, true, 0);
      uint32_t n = s.scal[S_NSIZE];
      // compute_parent1: closest seeded neighbour that is not the medoid = first such entry of the sorted list
      bool have = false;
      uint32_t parent = 0, mark = 0x01010101u;
      {
        uint32_t pick = 0xFFFFFFFFu;
        if (n > 0 && s.s_id[0] != a.medoid) pick = 0;
        else if (n > 1) pick = 1;
        if (pick != 0xFFFFFFFFu) { have = true; parent = s.s_id[pick]; mark = parent; }
      }
      uint32_t ncand = 1;
      if (have) { if (tid == 0) s.cand_id[ncand] = parent; ++ncand; }
      uint32_t iter = 1;
      __syncthreads();
      while (have || n > 0) {
        merge_worklist(a, s, n, iter == 1);
        {  // mark the node chosen by the previous parent selection (:1711-1714)
          const uint32_t ws = s.scal[S_WSIZE], cur = s.scal[S_CUR];
          for (uint32_t j = tid; j < ws; j += kThreads)
            if (s.wid(cur)[j] == mark) s.wv(cur)[j] = 1;
        }
        __syncthreads();
        if (have) {
          expand<T, MODE>(a, s, bloom, parent, false, ncand - 1);
          n = s.scal[S_NSIZE];
        } else {
          n = 0;
        }
        ++iter;
        // compute_parent2 (:1403-1458)
        float xd = 3.402823E+38f;
        uint32_t xid = 0;
        bool hasx = false;
        if (n > 0 && s.s_id[0] != a.medoid) { hasx = true; xd = s.s_d[0]; xid = s.s_id[0]; }
        else if (n > 1) { hasx = true; xd = s.s_d[1]; xid = s.s_id[1]; }
        (void)hasx;
        const uint32_t u = first_unvisited(s);
        const uint32_t ws = s.scal[S_WSIZE], cur = s.scal[S_CUR];
        have = false;
        if (u != 0xFFFFFFFFu) {
          have = true;
          if (xd < s.wd(cur)[u]) { parent = xid; mark = xid; }
          else { parent = s.wid(cur)[u]; if (tid == 0) s.wv(cur)[u] = 1; }
        } else if (ws > 0 && xd < s.wd(cur)[ws - 1]) {
          have = true; parent = xid; mark = xid;
        }
        if (have) { if (tid == 0 && ncand < a.cand_cap) s.cand_id[ncand] = parent; if (ncand < a.cand_cap) ++ncand; }
        __syncthreads();
        if (iter == a.max_iter - 1) break;
      }
      if (tid == 0) s.scal[S_NCAND] = ncand;
      // candidates selected but never expanded (cap reached) still need their exact distance
      __syncthreads();
      {
        const uint32_t done = (iter == a.max_iter - 1 && have) ? ncand - 1 : ncand;
        const uint32_t t = tid & 7, slot = tid >> 3;
        for (uint32_t b0 = done; b0 < ncand; b0 += 16) {  // uniform trip count: shuffles need whole warps
          const uint32_t i = b0 + slot;
          const uint32_t id = i < ncand ? s.cand_id[i] : a.medoid;
          const float d = l2_row_8lane<T>(row_ptr(a, id) + kAdjBytes, s.q_f, a.vec_units, t);
          if (t == 0 && i < ncand) s.cand_d[i] = d;
        }
      }
      __syncthreads();
      write_topk(a, q, s.cand_id, s.cand_d, ncand);
    } else {
      // ---- BANG_Inmemory / BANG_Exactdistance order: { expand(parent) ; merge ; pick first unvisited } ----
      uint32_t parent = a.medoid, iter = 1, ncand = 1;
      bool capped = false;
      for (;;) {
        expand<T, MODE>(a, s, bloom, parent, iter == 1, ncand - 1);
        merge_worklist(a, s, s.scal[S_NSIZE], iter == 1);
        const uint32_t u = first_unvisited(s);
        if (u == 0xFFFFFFFFu) break;
        const uint32_t cur = s.scal[S_CUR];
        parent = s.wid(cur)[u];
        if (tid == 0) {
          s.wv(cur)[u] = 1;
          if (MODE != kExact && ncand < a.cand_cap) s.cand_id[ncand] = parent;
          if (a.dump_ids && ncand < a.dump_stride) a.dump_ids[(size_t)q * a.dump_stride + ncand] = parent;
        }
        if (ncand < a.cand_cap) ++ncand;
        __syncthreads();
        if (iter == a.max_iter - 1) { capped = true; break; }
        ++iter;
      }
      if (MODE == kExact) {
        // top-k = head of the worklist (Exact parANN.cu:1273-1276)
        const uint32_t ws = s.scal[S_WSIZE], cur = s.scal[S_CUR];
        for (uint32_t r = tid; r < a.k; r += kThreads) {
          a.out_ids[(size_t)q * a.k + r] = r < ws ? (uint64_t)s.wid(cur)[r] : 0xFFFFFFFFull;
          a.out_dists[(size_t)q * a.k + r] = r < ws ? s.wd(cur)[r] : 3.402823466e+38f;
        }
      } else {
        __syncthreads();
        if (capped) {  // the last logged parent was never expanded: score it now
          if (tid < 32) {
            const float d = l2_row_8lane<T>(row_ptr(a, s.cand_id[ncand - 1]) + kAdjBytes, s.q_f, a.vec_units, tid & 7);
            if (tid == 0) s.cand_d[ncand - 1] = d;
          }
          __syncthreads();
        }
        write_topk(a, q, s.cand_id, s.cand_d, ncand);
      }
      if (tid == 0) s.scal[S_NCAND] = ncand;
    }
    __syncthreads();
    if (tid == 0) {
      if (a.dump_ids) {
        a.dump_ids[(size_t)q * a.dump_stride] = a.medoid;
        a.dump_n[q] = min(s.scal[S_NCAND], a.dump_stride);
      }
      if (a.st_hops) a.st_hops[q] = s.scal[S_NCAND];
      if (a.st_sumdeg) a.st_sumdeg[q] = s.scal[S_SUMDEG];
      if (a.st_npass) a.st_npass[q] = s.scal[S_NPASS];
    }
    __syncthreads();
  }
}

// Standalone stage-1 kernel (parity test of populate_pqDist_par): one CTA per query, table to global.
template <typename T>
__global__ void __launch_bounds__(kThreads) pq_table_kernel(const SearchArgs a, float* tables) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* q_f = reinterpret_cast<float*>(smem_raw);
  const uint32_t q = blockIdx.x;
  load_query<T>(a, q, q_f);
  __syncthreads();
  build_pq_table(a, q_f, tables + (size_t)q * a.n_chunks * 256);
}

}  // namespace bang
