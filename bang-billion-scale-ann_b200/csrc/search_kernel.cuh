// search_kernel.cuh — the fused, persistent, per-query-block traversal kernel for sm_100a.
//
// One CTA (128 threads) owns one query at a time and walks the whole greedy search for it without
// leaving the SM: PQ table build (stage 1) -> { adjacency fetch, visited filter (stage 4a), PQ/exact
// distances (stage 3), parent selection (stage 2), (dist,id) sort + worklist merge (stage 4b) }* ->
// exact re-rank + top-k (stage 5).  CTAs are persistent: the grid is sized to the number of resident
// CTAs and each CTA pulls query indices from a global counter.  What the reference does with ~5 kernel
// launches, 2 memsets, up to 5 PCIe copies and 3 stream syncs per hop (bang_search.cu:701-958) is one
// launch here; the LUT, worklist, candidate log and neighbour lists never leave shared memory.
//
// The search is a dependent pointer chase, so the kernel is organised around the per-hop critical path
// (ncu: profiles/r1_*): each warp runs its own slice of the expansion (16 adjacency ids: hash, one L2
// load per bloom word, fire-and-forget `red.or` to set, 8 lanes per accepted candidate for the
// distance) with two CTA barriers per hop; the next node to expand is decided from the unsorted
// distances BEFORE the sort/merge (it is the smaller of the best new candidate and the first unvisited
// worklist entry — exactly what the merge would produce) and its adjacency row is requested at once, so
// the merge overlaps the DRAM latency of the next hop.
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
#define BANG_B200_KERNEL_MAX_L 512  // MAX_L, bang.h:20

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
constexpr int kWarps = kThreads / 32;         // 4
constexpr int kIdsPerWarp = kMaxR / kWarps;   // 16 adjacency ids per warp (+ the medoid in warp 0 on the first hop)
constexpr int kStage = 20;                    // per-warp staging slots for accepted ids (<= 17 used)
constexpr int kWEnt = (BANG_B200_KERNEL_MAX_L + kThreads - 1) / kThreads;  // worklist entries per thread in a merge
constexpr uint32_t kNone = 0xFFFFFFFFu;

struct QState {
  float* q_f;        // [vec_units * E] query as fp32, zero padded
  float* lut;        // [n_chunks][256] (PQ modes); re-used for the candidates' exact distances in stage 5
  float* w_d;        // worklist [w_cap], sorted by distance
  uint32_t* w_id;
  uint8_t* w_v;      // visited flags
  uint32_t* n_id0;   // filtered neighbours, two buffers (hop parity) of kListCap
  float* n_d0;
  uint32_t* wstage;  // [kWarps][kStage]
  uint32_t* cand_id; // [cand_cap] expanded-node log (PQ modes)
  uint32_t* scal;
  __device__ __forceinline__ uint32_t* n_id(uint32_t par) const { return n_id0 + par * kListCap; }
  __device__ __forceinline__ float* n_d(uint32_t par) const { return n_d0 + par * kListCap; }
};
enum { S_CNT0 = 0, S_CNT1, S_QUERY, S_SUMDEG, S_NPASS, S_POS0, S_COUNT = 8 };

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline size_t lut_region_bytes(int mode, uint32_t n_chunks, uint32_t cand_cap) {
  if (mode == kExact) return 0;
  const size_t lut = (size_t)n_chunks * 256 * 4, cd = align_up((size_t)cand_cap * 4, 16);
  return lut > cd ? lut : cd;
}

template <typename T>
__host__ __device__ inline size_t smem_bytes(int mode, uint32_t n_chunks, uint32_t vec_units, uint32_t L, uint32_t cand_cap) {
  size_t b = 0;
  b += align_up((size_t)vec_units * Elem<T>::kPerUnit * 4, 16);
  b += lut_region_bytes(mode, n_chunks, cand_cap);
  b += align_up(L, 16) * 9;                         // worklist: dist + id + visited
  b += (size_t)kListCap * 4 * 2 * 2;                // neighbour lists, two parities
  b += (size_t)kWarps * kStage * 4;
  if (mode != kExact) b += align_up((size_t)cand_cap * 4, 16);
  b += S_COUNT * 4;
  return align_up(b, 16);
}

template <typename T>
__device__ __forceinline__ void carve(QState& s, uint8_t* base, int mode, const SearchArgs& a) {
  size_t o = 0;
  s.q_f = (float*)(base + o); o += align_up((size_t)a.vec_units * Elem<T>::kPerUnit * 4, 16);
  s.lut = (float*)(base + o); o += lut_region_bytes(mode, a.n_chunks, a.cand_cap);
  const size_t wcap = align_up(a.L, 16);
  s.w_d = (float*)(base + o); o += wcap * 4;
  s.w_id = (uint32_t*)(base + o); o += wcap * 4;
  s.w_v = (uint8_t*)(base + o); o += wcap;
  s.n_id0 = (uint32_t*)(base + o); o += (size_t)kListCap * 4 * 2;
  s.n_d0 = (float*)(base + o); o += (size_t)kListCap * 4 * 2;
  s.wstage = (uint32_t*)(base + o); o += (size_t)kWarps * kStage * 4;
  s.cand_id = (uint32_t*)(base + o); if (mode != kExact) o += align_up((size_t)a.cand_cap * 4, 16);
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

// adjacency prefetch: lanes 0..15 of warp w request ids [16w, 16w+16) of `node`'s HBM row
__device__ __forceinline__ uint32_t fetch_adj(const SearchArgs& a, uint32_t node) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t v = kNoNbr;
  if (lane < (uint32_t)kIdsPerWarp) v = ld_nc_u32(row_ptr(a, node) + 4 * (kIdsPerWarp * warp + lane));
  return v;
}

// ------------------------------------------------------------------------------------------------
// expansion of one node = stages 4a + 3.  Every warp owns 16 adjacency ids (already requested by
// fetch_adj): visited filter with snapshot semantics (all tests of a list precede all insertions — one
// barrier), then 8 lanes per accepted candidate for the PQ / exact distance.  Accepted (id, dist) pairs
// land unordered in the neighbour buffer of this hop's parity; returns their count.
//   filter   neighbor_filtering_new + hashFn1_d/2_d   bang_search.cu:1140-1189 (Exact: hash 1 only)
//   PQ dist  compute_neighborDist_par                 bang_search.cu:1201-1241: lane t owns chunks t, t+8, ...
//            ascending, partials combined by the 8-lane tree.  The HBM code rows are permuted at load so
//            lane t's chunks 32g+t, 32g+8+t, 32g+16+t, 32g+24+t are one aligned 32-bit word: one 32-byte
//            sector per candidate per 32 chunks, fully used.
//   exact    compute_neighborDist_par                 BANG_Exactdistance/parANN.cu:1139-1179
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE>
__device__ __forceinline__ uint32_t expand(const SearchArgs& a, const QState& s, uint32_t* bloom, uint32_t my_nb, bool first,
                                           uint32_t par) {
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t id = kNoNbr;
  if (lane < (uint32_t)kIdsPerWarp) id = my_nb;
  else if (first && warp == 0 && lane == (uint32_t)kIdsPerWarp) id = a.medoid;  // [medoid] ++ adj(medoid) on the first hop
  const bool valid = id != kNoNbr;
  uint32_t w1 = 0, b1 = 0, w2 = 0, b2 = 0;
  bool acc = false;
  if (valid) {
    const uint32_t p1 = hash1(id);
    w1 = p1 >> 5; b1 = 1u << (p1 & 31);
    if (MODE == kExact) {
      acc = (__ldcg(bloom + w1) & b1) == 0;
    } else {
      const uint32_t p2 = hash2(id);
      w2 = p2 >> 5; b2 = 1u << (p2 & 31);
      const uint32_t x1 = __ldcg(bloom + w1), x2 = __ldcg(bloom + w2);
      acc = !((x1 & b1) && (x2 & b2));
    }
  }
  // all tests of this list are done before any insertion (also counts the degree for the statistics)
  const uint32_t deg = __syncthreads_count(valid && lane < (uint32_t)kIdsPerWarp);
  if (acc) {  // results unused -> RED.OR, nothing waits on them
    atomicOr(bloom + w1, b1);
    if (MODE != kExact) atomicOr(bloom + w2, b2);
  }
  const uint32_t amask = __ballot_sync(0xffffffffu, acc);
  const uint32_t cnt = __popc(amask);
  uint32_t base = 0;
  if (lane == 0 && cnt) base = atomicAdd(&s.scal[S_CNT0 + par], cnt);
  base = __shfl_sync(0xffffffffu, base, 0);
  uint32_t* stage = s.wstage + warp * kStage;
  if (acc) stage[__popc(amask & ((1u << lane) - 1u))] = id;
  __syncwarp();
  uint32_t* n_id = s.n_id(par);
  float* n_d = s.n_d(par);
  const uint32_t t = lane & 7, g = lane >> 3;  // 4 candidates per warp pass
  constexpr int kPass = (kIdsPerWarp + 1 + 3) / 4;  // 5
  if (MODE == kExact) {
    for (uint32_t k0 = 0; k0 < cnt; k0 += 4) {
      const uint32_t k = k0 + g;
      const uint32_t cid = k < cnt ? stage[k] : a.medoid;
      const float d = l2_row_8lane<T>(row_ptr(a, cid) + kAdjBytes, s.q_f, a.vec_units, t);
      if (t == 0 && k < cnt) { n_id[base + k] = cid; n_d[base + k] = d; }
    }
  } else if (a.n_chunks <= 32) {
    uint32_t w[kPass], cid[kPass];
#pragma unroll
    for (int p = 0; p < kPass; ++p) {
      const uint32_t k = p * 4 + g;
      w[p] = 0; cid[p] = 0;
      if (k < cnt) { cid[p] = stage[k]; w[p] = ld_nc_u32(a.codes + (size_t)cid[p] * a.code_stride + 4 * t); }
    }
#pragma unroll
    for (int p = 0; p < kPass; ++p) {
      if (p * 4 < (int)cnt) {  // warp-uniform
        const uint32_t k = p * 4 + g;
        float sum = 0.0f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t c = t + 8 * b;
          if (c < a.n_chunks) sum = __fadd_rn(sum, s.lut[c * 256 + ((w[p] >> (8 * b)) & 0xff)]);
        }
        sum = tree8(sum);
        if (t == 0 && k < cnt) { n_id[base + k] = cid[p]; n_d[base + k] = sum; }
      }
    }
  } else {
    const uint32_t groups = (a.n_chunks + 31) >> 5;
    for (uint32_t k0 = 0; k0 < cnt; k0 += 4) {
      const uint32_t k = k0 + g;
      const uint32_t cid = k < cnt ? stage[k] : 0u;
      const uint8_t* row = a.codes + (size_t)cid * a.code_stride + 4 * t;
      float sum = 0.0f;
      for (uint32_t gg = 0; gg < groups; gg += 2) {
        const uint32_t wa = ld_nc_u32(row + gg * 32);
        const uint32_t wb = (gg + 1 < groups) ? ld_nc_u32(row + (gg + 1) * 32) : 0u;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t c = gg * 32 + t + 8 * b;
          if (c < a.n_chunks) sum = __fadd_rn(sum, s.lut[c * 256 + ((wa >> (8 * b)) & 0xff)]);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t c = (gg + 1) * 32 + t + 8 * b;
          if (c < a.n_chunks) sum = __fadd_rn(sum, s.lut[c * 256 + ((wb >> (8 * b)) & 0xff)]);
        }
      }
      sum = tree8(sum);
      if (t == 0 && k < cnt) { n_id[base + k] = cid; n_d[base + k] = sum; }
    }
  }
  __syncthreads();
  const uint32_t n = s.scal[S_CNT0 + par];
  if (tid == 0) {
    s.scal[S_CNT0 + (par ^ 1u)] = 0;  // next hop's counter; its last readers passed the barrier above
    s.scal[S_SUMDEG] += deg;
    s.scal[S_NPASS] += n;
  }
  return n;
}

// (dist, id)-minimum of the unsorted neighbour list, the number of entries closer than `maxd`, and the
// medoid's distance if it is in the list.  Every warp computes the same values redundantly from shared
// memory, so no barrier is needed to publish the decision that follows.
struct Best { float d; uint32_t id; uint32_t below; float med_d; bool med_in; };
__device__ __forceinline__ Best scan_neighbours(const QState& s, uint32_t par, uint32_t n, uint32_t medoid, bool skip_medoid,
                                                float maxd) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t* n_id = s.n_id(par);
  const float* n_d = s.n_d(par);
  Best b{3.402823466e+38f, kNone, 0u, 0.0f, false};
  for (uint32_t i = lane; i < n; i += 32) {
    const float d = n_d[i];
    const uint32_t id = n_id[i];
    b.below += d < maxd ? 1u : 0u;
    if (id == medoid) { b.med_in = true; b.med_d = d; if (skip_medoid) continue; }
    if (key_less(d, id, b.d, b.id)) { b.d = d; b.id = id; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, b.d, off);
    const uint32_t oid = __shfl_xor_sync(0xffffffffu, b.id, off);
    if (key_less(od, oid, b.d, b.id)) { b.d = od; b.id = oid; }
    b.below += __shfl_xor_sync(0xffffffffu, b.below, off);
    const float omd = __shfl_xor_sync(0xffffffffu, b.med_d, off);
    const bool omi = __shfl_xor_sync(0xffffffffu, (int)b.med_in, off) != 0;
    if (omi) { b.med_in = true; b.med_d = omd; }
  }
  return b;
}

// first unvisited worklist entry at or after `start` (warp-redundant, shared memory only)
__device__ __forceinline__ uint32_t scan_unvisited(const QState& s, uint32_t start, uint32_t ws) {
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t b0 = start; b0 < ws; b0 += 32) {
    const uint32_t j = b0 + lane;
    const uint32_t m = __ballot_sync(0xffffffffu, j < ws && s.w_v[j] == 0);
    if (m) return b0 + (uint32_t)__ffs(m) - 1u;
  }
  return kNone;
}

// ------------------------------------------------------------------------------------------------
// stage 4b: (dist, id) sort of the new neighbours + merge into the worklist, in place.
// compute_BestLSets_par_sort_msort + compute_BestLSets_par_merge (bang_search.cu:1533-1585, 1605-1715):
// only the `nb` closest new entries take part (nb as the reference computes nbrsBound); a new entry
// goes before old entries of equal distance; the list is truncated to L.  New entries are unvisited,
// except `flag_id` (the node just chosen for expansion / the reference's d_mark) and, in the first
// merge, the medoid.  Returns the new size; scal[S_POS0] = position of the closest new entry.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t merge_worklist(const SearchArgs& a, const QState& s, uint32_t par, uint32_t n, uint32_t nb,
                                                   uint32_t ws, bool first, uint32_t flag_id) {
  const uint32_t tid = threadIdx.x;
  uint32_t* s_id = s.n_id(par ^ 1u);  // the other parity's buffer is idle until the next hop's expansion
  float* s_d = s.n_d(par ^ 1u);
  const uint32_t* n_id = s.n_id(par);
  const float* n_d = s.n_d(par);
  if (tid < n) {
    const float d = n_d[tid];
    const uint32_t id = n_id[tid];
    uint32_t r = 0;
    for (uint32_t j = 0; j < n; ++j) r += key_less(n_d[j], n_id[j], d, id) ? 1u : 0u;
    if (r < nb) { s_d[r] = d; s_id[r] = id; }
  }
  __syncthreads();
  if (first) {  // iter == 1 branch (:1636-1646): the worklist is the head of the sorted list
    if (tid < nb) {
      const uint32_t id = s_id[tid];
      s.w_d[tid] = s_d[tid];
      s.w_id[tid] = id;
      s.w_v[tid] = (id == a.medoid || id == flag_id) ? 1 : 0;
    }
    if (tid == 0) s.scal[S_POS0] = 0;
    __syncthreads();
    return nb;
  }
  const uint32_t newsize = min(ws + nb, a.L);
  float od[kWEnt];
  uint32_t oid[kWEnt], opos[kWEnt];
  uint8_t ov[kWEnt];
#pragma unroll
  for (int e = 0; e < kWEnt; ++e) {  // old entry: position = index + upper_bound(new, d)
    const uint32_t j = tid + e * kThreads;
    opos[e] = kNone;
    if (j < ws) {
      od[e] = s.w_d[j]; oid[e] = s.w_id[j]; ov[e] = s.w_v[j];
      uint32_t lo = 0, hi = nb;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (od[e] >= s_d[mid]) lo = mid + 1; else hi = mid;
      }
      opos[e] = j + lo;
    }
  }
  uint32_t npos = kNone, nid = 0;
  float nd = 0.0f;
  if (tid < nb) {  // new entry: position = index + lower_bound(W, d)  (new before old on ties)
    nd = s_d[tid]; nid = s_id[tid];
    uint32_t lo = 0, hi = ws;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (nd <= s.w_d[mid]) hi = mid; else lo = mid + 1;
    }
    npos = lo + tid;
    if (tid == 0) s.scal[S_POS0] = npos;
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < kWEnt; ++e)
    if (opos[e] < newsize) { s.w_d[opos[e]] = od[e]; s.w_id[opos[e]] = oid[e]; s.w_v[opos[e]] = ov[e]; }
  if (npos < newsize) { s.w_d[npos] = nd; s.w_id[npos] = nid; s.w_v[npos] = (nid == flag_id) ? 1 : 0; }
  __syncthreads();
  return newsize;
}

// number of new entries the reference's merge admits (nbrsBound, bang_search.cu:1651-1656)
__device__ __forceinline__ uint32_t admit_count(uint32_t below, uint32_t n, uint32_t ws, uint32_t L) {
  return max(min(below, min(L, n)), min(L - ws, n));
}

// ------------------------------------------------------------------------------------------------
// stage 5: exact distances of the expanded nodes (compute_L2Dist, bang_search.cu:1254-1299) with
// coalesced 16-byte loads, 8 lanes per row, two rows in flight per lane group; then the top-k by
// (exact distance, id) (compute_NearestNeighbours, :1312-1368).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void rerank_and_write(const SearchArgs& a, const QState& s, uint32_t q, uint32_t n) {
  const uint32_t tid = threadIdx.x, t = tid & 7, slot = tid >> 3;
  float* cd = s.lut;  // the PQ table is dead by now
  __syncthreads();
  for (uint32_t b0 = 0; b0 < n; b0 += 32) {
    const uint32_t i0 = b0 + slot, i1 = b0 + 16 + slot;
    const uint32_t id0 = i0 < n ? s.cand_id[i0] : a.medoid, id1 = i1 < n ? s.cand_id[i1] : a.medoid;
    const float d0 = l2_row_8lane<T>(row_ptr(a, id0) + kAdjBytes, s.q_f, a.vec_units, t);
    const float d1 = l2_row_8lane<T>(row_ptr(a, id1) + kAdjBytes, s.q_f, a.vec_units, t);
    if (t == 0 && i0 < n) cd[i0] = d0;
    if (t == 0 && i1 < n) cd[i1] = d1;
  }
  __syncthreads();
  for (uint32_t i = tid; i < n; i += kThreads) {
    const float di = cd[i];
    const uint32_t idi = s.cand_id[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; ++j) rank += key_less(cd[j], s.cand_id[j], di, idi) ? 1u : 0u;
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
    __syncthreads();
    if (tid == 0) s.scal[S_QUERY] = atomicAdd(a.counter, 1u);
    __syncthreads();
    const uint32_t q = s.scal[S_QUERY];
    if (q >= a.Q) break;

    // ---- per-query setup: query -> smem, bloom filter cleared, PQ table built in place ----
    uint32_t my_nb = fetch_adj(a, a.medoid);  // the first hop's adjacency row travels during the setup
    load_query<T>(a, q, s.q_f);
    {
      uint4* b4 = reinterpret_cast<uint4*>(bloom);
      for (uint32_t i = tid; i < kBloomWords / 4; i += kThreads) b4[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
      s.scal[S_CNT0] = 0; s.scal[S_CNT1] = 0; s.scal[S_SUMDEG] = 0; s.scal[S_NPASS] = 0; s.scal[S_POS0] = 0;
      if (MODE != kExact) s.cand_id[0] = a.medoid;  // bang_init: the medoid is every query's first candidate (:455-462)
    }
    __syncthreads();
    if (MODE != kExact) {
      build_pq_table(a, s.q_f, s.lut);
      __syncthreads();
    }

    uint32_t ws = 0, fu = kNone, ncand = 1, iter = 1;
    auto log_parent = [&](uint32_t node) {
      if (tid == 0) {
        if (MODE != kExact && ncand < a.cand_cap) s.cand_id[ncand] = node;
        if (a.dump_ids && ncand < a.dump_stride) a.dump_ids[(size_t)q * a.dump_stride + ncand] = node;
      }
      if (ncand < a.cand_cap) ++ncand;
    };

    if (MODE == kBase) {
      // ---- BANG_Base (A.1, A.2): seed, then { merge(previous) ; expand(parent) ; compute_parent2 } ----
      uint32_t par = 1;
      uint32_t n = expand<T, MODE>(a, s, bloom, my_nb, true, par);
      Best b = scan_neighbours(s, par, n, a.medoid, true, 0.0f);
      bool have = b.id != kNone;               // compute_parent1 (:1464-1521): closest seeded neighbour, medoid excluded
      uint32_t parent = b.id, mark = have ? b.id : 0x01010101u;
      if (have) log_parent(parent);
      uint32_t pend_n = n, pend_nb = min(n, a.L), scan_from = 0;
      while (have || pend_n > 0) {
        if (have) my_nb = fetch_adj(a, parent);  // in flight during the merge
        if (pend_n > 0 && pend_nb > 0) {          // sort + merge of the previous neighbours (:726,:738), mark (:1711-1714)
          ws = merge_worklist(a, s, par, pend_n, pend_nb, ws, iter == 1, mark);
          scan_from = min(scan_from, s.scal[S_POS0]);
        }
        fu = scan_unvisited(s, scan_from, ws);
        scan_from = fu == kNone ? ws : fu;
        if (have) { par ^= 1u; n = expand<T, MODE>(a, s, bloom, my_nb, false, par); }  // `par` = buffer of the pending list
        else n = 0u;
        ++iter;
        // compute_parent2 (:1403-1458)
        const float maxd = ws > 0 ? s.w_d[ws - 1] : 0.0f;
        b = scan_neighbours(s, par, n, a.medoid, true, maxd);
        const bool hasx = b.id != kNone;
        have = false;
        if (fu != kNone) {
          have = true;
          if (hasx && b.d < s.w_d[fu]) { parent = b.id; mark = b.id; }
          else { parent = s.w_id[fu]; if (tid == 0) s.w_v[fu] = 1; scan_from = fu + 1; }
        } else if (ws > 0 && hasx && b.d < maxd) {
          have = true; parent = b.id; mark = b.id;
        }
        if (have) log_parent(parent);
        pend_n = n;
        pend_nb = n ? admit_count(b.below, n, ws, a.L) : 0u;
        if (iter == a.max_iter - 1) break;
      }
      rerank_and_write<T>(a, s, q, ncand);
    } else {
      // ---- BANG_Inmemory / BANG_Exactdistance (A.2', A.2''): { expand(parent) ; merge ; first unvisited } ----
      // The first unvisited entry after the merge is decided before it: the closest new entry if it is
      // admitted and not farther than the first unvisited old entry (new goes before old on ties).
      uint32_t parent = a.medoid;
      for (;;) {
        const bool first = iter == 1;
        const uint32_t par = iter & 1u;
        const uint32_t n = expand<T, MODE>(a, s, bloom, my_nb, first, par);
        const float maxd = ws > 0 ? s.w_d[ws - 1] : 0.0f;
        const Best b = scan_neighbours(s, par, n, a.medoid, first, maxd);
        uint32_t nb;
        bool have = false, from_new = false;
        if (first) {
          nb = min(n, a.L);
          const uint32_t rank_x = (b.med_in && key_less(b.med_d, a.medoid, b.d, b.id)) ? 1u : 0u;
          if (b.id != kNone && rank_x < nb) { have = true; from_new = true; parent = b.id; }
        } else {
          nb = n ? admit_count(b.below, n, ws, a.L) : 0u;
          if (nb > 0 && (fu == kNone || b.d <= s.w_d[fu])) { have = true; from_new = true; parent = b.id; }
          else if (fu != kNone) { have = true; parent = s.w_id[fu]; }
        }
        if (!have) break;  // nothing unvisited and nothing admitted: the merge would be a no-op
        uint32_t scan_from = fu == kNone ? ws : fu;
        if (!from_new) { if (tid == 0) s.w_v[fu] = 1; scan_from = fu + 1; }
        log_parent(parent);  // thread 0, Inmemory parANN.cu:1399-1418
        const bool capped = iter == a.max_iter - 1;
        if (!capped) my_nb = fetch_adj(a, parent);  // in flight during the merge
        if (nb > 0) {
          ws = merge_worklist(a, s, par, n, nb, ws, first, from_new ? parent : kNone);
          scan_from = min(scan_from, s.scal[S_POS0]);
        }
        fu = scan_unvisited(s, scan_from, ws);
        if (capped) break;
        ++iter;
      }
      if (MODE == kExact) {
        // top-k = head of the worklist (Exact parANN.cu:1273-1276)
        __syncthreads();
        for (uint32_t r = tid; r < a.k; r += kThreads) {
          a.out_ids[(size_t)q * a.k + r] = r < ws ? (uint64_t)s.w_id[r] : 0xFFFFFFFFull;
          a.out_dists[(size_t)q * a.k + r] = r < ws ? s.w_d[r] : 3.402823466e+38f;
        }
      } else {
        rerank_and_write<T>(a, s, q, ncand);
      }
    }
    if (tid == 0) {
      if (a.dump_ids) {
        a.dump_ids[(size_t)q * a.dump_stride] = a.medoid;
        a.dump_n[q] = min(ncand, a.dump_stride);
      }
      if (a.st_hops) a.st_hops[q] = ncand;
      if (a.st_sumdeg) a.st_sumdeg[q] = s.scal[S_SUMDEG];
      if (a.st_npass) a.st_npass[q] = s.scal[S_NPASS];
    }
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
