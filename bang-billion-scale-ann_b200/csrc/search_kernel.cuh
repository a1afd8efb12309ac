// search_kernel.cuh — the fused, persistent traversal kernel for sm_100a.
//
// One warp owns one query at a time and walks the whole greedy search for it without leaving the SM:
// { adjacency fetch, visited filter (stage 4a), PQ/exact distances (stages 1+3, the table entries are
// evaluated on demand), parent selection (stage 2), (dist,id) sort + worklist merge (stage 4b) }* -> exact
// re-rank + top-k (stage 5).  CTAs (up to 32 query warps + one pivot table, 24 by default) are persistent: the grid is sized
// to the number of resident CTAs and every warp pulls query indices from a global counter.  What the reference
// does with ~5 kernel launches, 2 memsets, up to 5 PCIe copies and 3 stream syncs per hop
// (bang_search.cu:701-958) is one launch here; the worklist, candidate log and neighbour lists never leave
// shared memory.
//
// Why one warp per query and no per-query LUT: the search is a dependent pointer chase whose per-hop
// work is tiny (64 hashes, ~10 PQ distances, a 150-entry merge), so throughput = resident queries /
// per-hop latency.  Measured on B200 (profiles/r1_*): time per 10k-query batch falls almost linearly with
// the number of resident queries per SM, and the reference's per-query fp32 table (32 KB at m = 32) caps
// residency at 6.  A table entry is a pure function of (query, chunk, code): tbl[c][k] = the fmaf chain
// over the chunk's dimensions.  Evaluating that same chain on demand from ONE pivot table shared by all
// the warps of the CTA (128 KB of shared memory at D = 128) gives bit-identical distances and leaves
// ~2.5 KB of private state per query (D = 128, L = 176), i.e. 24 resident queries per SM below the 196 KB
// shared-memory step.  CTA barriers are used once (to publish the pivot table); afterwards every warp runs on
// its own with __syncwarp and redux.sync.
// Per hop:
//   * the adjacency row (256 B) was requested at the end of the previous hop, one 8-byte load per lane;
//   * visited filter with snapshot semantics: 4 filter blocks (8 B each) per lane in one L2 round trip, addressed as
//     32-bit offsets from one kernel parameter; insertion = one atomic on the block's count (hands out the byte
//     position) + a byte store after the distance phase, so nothing waits on the atomic's round trip; the rare
//     overflowing blocks sit behind one warp vote;
//   * 8 lanes per accepted candidate for the distance, 16 candidates' code loads in flight;
//   * the next node to expand is decided from the unsorted distances BEFORE the sort/merge (it is the
//     smaller of the best admitted new candidate and the first unvisited worklist entry — exactly what
//     the merge would produce) and its adjacency row is requested at once, so the merge overlaps the
//     DRAM latency of the next hop;
//   * in-place merge, 32 entries at a time from the tail, only from the first insertion point on.
// Where a hop's ~930 warp instructions and its stalls go: profiles/r2_hop_diet.md.
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
#ifdef BANG_PHASE_TIMERS
#include <cstdio>  // (the phase-clock build pins loads in place with never-taken printf branches)
#endif


namespace bang {

constexpr int kThreads = 32;            // threads per query = one warp (see the header comment)
constexpr int kMaxWarpsPerCta = 32;  // PQ modes: one CTA of up to 32 query warps per SM (kernels compiled for 16 / 24 / 32, see bang_search_kernel)
constexpr int kMaxR = 64;               // MAX_R, bang_search.cu:35
constexpr int kListCap = kMaxR + 8;     // medoid + R neighbours (65), padded to whole 16-byte groups (68); the last two id slots hold statistics
constexpr uint32_t kVisitedBit = 0x80000000u;  // worklist entries carry their visited flag in the top bit of the id word (ids < 2^31, checked at load)
constexpr uint32_t kBfEntries = 399887u;  // BF_ENTRIES, bang_search.cu:48
// The visited filter has the reference's semantics — a 399887-slot bit array addressed by two hashes — but is
// stored sparsely: 1569 blocks of 255 slots, each block = 8 bytes holding up to 7 one-byte offsets of its set
// slots (0xFF = empty) plus a count byte.  A typical search sets a few thousand slots (0.8 % of the array: two per
// block on average), so 12.5 KB replace the 50 KB bitmap (and the reference's 400 KB byte array) with identical
// answers; the filters of all resident queries (32 per SM: 59 MB) then stay in L2 next to the PQ codes
// (profiles/r1_l2_footprint.md, r2_*).  A block that receives an 8th slot spills into its own 255-bit bitmap
// (32 bytes, in a separate region that is cleared lazily at the moment of the spill and is otherwise never
// touched), so the filter stays exact and O(1) for any number of insertions — hard queries on 10^7+ point
// graphs insert 15-20 thousand slots (profiles/r1_c5.md).
constexpr uint32_t kVisBlocks = (kBfEntries + 254u) / 255u;        // 1569
constexpr uint32_t kVisSlotsPerBlock = 7;                           // offsets a block holds before it spills
constexpr uint32_t kVisBlockBytes = ((kVisBlocks * 8u + 127u) / 128u) * 128u;    // 12672: the 8-byte blocks of one query
constexpr uint32_t kVisBitmapBytes = ((kVisBlocks * 32u + 127u) / 128u) * 128u;  // 50304: the spill bitmaps of one query
// One allocation holds [block areas of all resident warps][bitmap areas of all resident warps]; kBloomWords is
// the per-warp allocation unit in 32-bit words.
constexpr uint32_t kBloomWords = (kVisBlockBytes + kVisBitmapBytes) / 4u;
constexpr uint32_t kNoNbr = 0xFFFFFFFFu;  // padding id in the HBM adjacency rows
constexpr int kAdjBytes = kMaxR * 4;    // 256 B adjacency block at the head of each HBM row
// Optional second block (PQ modes, `PH` kernels): the two visited-filter slots of every neighbour, precomputed at load —
// they are a pure function of the neighbour's id (hashFn1_d / hashFn2_d, bang_search.cu:1168-1189), and evaluating the two
// 64-bit hashes + the modulus per neighbour is 13 % of a hop's instructions.  8 bytes per neighbour: two words
// (block byte offset | offset in the block << 24), see vis_slot_word.  A byte-for-instruction trade: the row grows by 512 B.
constexpr int kSlotBytes = kMaxR * 8;
template <bool PH> __host__ __device__ constexpr int adj_bytes() { return PH ? kAdjBytes + kSlotBytes : kAdjBytes; }
constexpr int kMaxShards = 8;
#define BANG_B200_KERNEL_MAX_L 512  // MAX_L, bang.h:20

enum Mode : int { kBase = 0, kInmemory = 1, kExact = 2 };

struct SearchArgs {
  // ---- index (HBM layout) ----
  const uint8_t* rows[kMaxShards];  // shard s holds ids with id % n_shards == s at local row id / n_shards
  uint32_t n_shards;
  uint32_t shard_shift;     // log2(n_shards) when n_shards is a power of two above 1, else 0
  uint32_t row_stride;      // bytes, multiple of 32: [ R x u32 adjacency | vector padded to 16 B ]
  const uint8_t* codes;     // [N][code_stride], bytes permuted per 32-chunk group (see repack_codes)
  uint32_t code_stride;
  uint32_t n_chunks;
  const float* pivT;        // [D][256]   pivots transposed as the reference does at load (:281-285); stage-1 kernel only
  const float* piv;         // [256][piv_row]  the search kernel's pivot table (copied into shared memory once per CTA): file order
                            //            [256][D], or, for the CS = 4 kernels, every chunk zero-padded to 4 dimensions ([256][32][4])
  uint32_t piv_row;         // floats per row of `piv`: D, or 128 for the padded table
  uint32_t chunk4;          // 1..4 when there are exactly 32 chunks of that many dimensions each (selects the CS = 4 kernel), else 0
  const float* centroid;    // [D]
  const uint32_t* chunk_off;  // [n_chunks+1]
  uint32_t D;               // dims of the index
  uint32_t vec_units;       // 16-byte units per vector (D*sizeof(T) rounded up / 16)
  uint32_t medoid;
  uint32_t med_blk8[2], med_off[2], med_slots;  // the medoid's visited-filter slots (block byte offset, offset in the block), see set_medoid
  // ---- search ----
  uint32_t L, k, Q;
  uint32_t q_dim;           // elements per query row (D, or D-1 for MIPS)
  uint32_t max_iter;        // L+50 (Base) / L+120 (Inmemory) / 4L+20 (Exact)
  uint32_t cand_cap;        // max_iter + 1
  const void* queries;      // device T[Q][q_dim]
  uint64_t* out_ids;        // device [Q][k]
  float* out_dists;         // device [Q][k] (query-major)
  uint32_t* bloom;          // device: one sparse visited filter (kBloomWords words) per resident query warp
  uint32_t* cand_log;       // device: expanded-node log, cand_cap ids per resident query warp (PQ modes; read back by the re-rank)
  uint32_t piv_global;      // 1: the pivot table stays in global memory (it does not fit in shared memory: D above ~215)
  uint32_t code_prefetch;   // 1: request the PQ codes of every neighbour towards L2 while the filter is consulted
  uint32_t stop_on_empty_hop;  // 1 (Exactdistance searches): a hop without new neighbours ends the query, see the kernel; 0 for the index builder
  uint32_t* counter;        // device work counter (zeroed before launch)
  uint32_t* st_hops;        // device [Q] or null
  uint32_t* st_sumdeg;
  uint32_t* st_npass;
  // optional dump of the expanded-node log (used by the index builder): ids [Q][dump_stride], count [Q]
  uint32_t* dump_ids;
  uint32_t* dump_n;
  uint32_t dump_stride;
  long long* st_phase;      // debug builds (-DBANG_PHASE_TIMERS): [Q][16] per-phase SM clocks, else null
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t hash1(uint32_t x) {  // hashFn1_d, bang_search.cu:1168-1178
  uint64_t h = 0xcbf29ce4ull;
  h = (h ^ (x & 0xff)) * 0x01000193ull;
  h = (h ^ ((x >> 8) & 0xff)) * 0x01000193ull;
  h = (h ^ ((x >> 16) & 0xff)) * 0x01000193ull;
  h = (h ^ ((x >> 24) & 0xff)) * 0x01000193ull;
  return (uint32_t)(h % kBfEntries);
}
__host__ __device__ __forceinline__ uint32_t hash2(uint32_t x) {  // hashFn2_d, bang_search.cu:1179-1189
  uint64_t h = 0x84222325ull;
  h = (h ^ (x & 0xff)) * 0x1B3ull;
  h = (h ^ ((x >> 8) & 0xff)) * 0x1B3ull;
  h = (h ^ ((x >> 16) & 0xff)) * 0x1B3ull;
  h = (h ^ ((x >> 24) & 0xff)) * 0x1B3ull;
  return (uint32_t)(h % kBfEntries);
}

// host side: the entry point and its visited-filter slots (a pure function of the id, so computed once per launch here)
inline void set_medoid(SearchArgs& a, uint32_t medoid, bool one_hash /* Exactdistance */) {
  a.medoid = medoid;
  const uint32_t p[2] = {hash1(medoid), hash2(medoid)};
  for (int h = 0; h < 2; ++h) { a.med_blk8[h] = p[h] / 255u * 8u; a.med_off[h] = p[h] % 255u; }
  a.med_slots = (one_hash || p[0] == p[1]) ? 1u : 2u;
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
  // local HBM or a peer mapping over NVLink; 2, 4 or 8 shards: mask and shift (shard_shift > 0) instead of a division
  const uint32_t s = a.shard_shift ? (id & (a.n_shards - 1u)) : id % a.n_shards;
  const uint32_t r = a.shard_shift ? (id >> a.shard_shift) : id / a.n_shards;
  return a.rows[s] + (size_t)r * a.row_stride;
}

// L2 residency control.  The per-query visited filters (25 KB of blocks each, re-read every hop) are the only
// data with reuse; graph rows and PQ codes are touched once per query.  Streaming loads therefore carry an
// evict-first L2 policy and bypass L1, filter accesses an evict-last policy, so the gathers do not push the
// filters out of the 126 MB L2 (ncu: profiles/r1_*: DRAM bytes vs algorithmic bytes).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// createpolicy is not free (a few instructions per use): the kernel creates both policies once and keeps them in
// registers; helpers without an explicit policy argument are for the cold paths (builder, re-rank).
__device__ __forceinline__ uint4 ld_nc_u4(const void* p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint4 ld_nc_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_nc_u32(const void* p, uint64_t pol) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}
// Speculative L2 prefetch of one 32-byte sector: issued for the PQ codes of ALL neighbours of the expanded node
// as soon as their ids arrive, i.e. in parallel with the visited-filter round trip.  The real code loads (only
// for the candidates that pass the filter) then hit L2 instead of paying a second dependent DRAM access.  Off by
// default since the instruction diet: the wait it covered is hidden by the other warps now, and it cost 47 % more
// DRAM bytes at 10^7 points (SearchArgs::code_prefetch, profiles/r2u_dram_bytes_prefetch_l2fetch.log).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// ---- sparse visited filter (see kVisBlocks) -------------------------------------------------------
// A slot is addressed by the byte offset of its 8-byte block from SearchArgs::bloom (32 bits: the filters of a whole
// grid are a few hundred MB) and its offset inside the block; every access is  bloom + o  — one 64-bit add on a
// kernel parameter — instead of a pointer rebuilt from the warp's number.
struct VisSlot { uint32_t o, off; };
struct VisBase { uint32_t blocks_o, bitmaps_o; };  // byte offsets of this warp's block area / spill-bitmap area from SearchArgs::bloom
__device__ __forceinline__ VisSlot vis_slot(const VisBase& vb, uint32_t pos) {
  const uint32_t blk = __umulhi(pos, 0x80808081u) >> 7;  // pos / 255 (exact for pos < 2^31)
  VisSlot v;
  v.off = pos - blk * 255u;
  v.o = vb.blocks_o + blk * 8u;
  return v;
}
// the precomputed form of a slot (rows with the slot block): block byte offset in the low half, offset in the block in the top byte
__host__ __device__ __forceinline__ uint32_t vis_slot_word(uint32_t pos) { return (pos / 255u * 8u) | ((pos % 255u) << 24); }
__device__ __forceinline__ VisSlot vis_slot_from_word(const VisBase& vb, uint32_t w) {
  VisSlot v;
  v.off = w >> 24;
  v.o = vb.blocks_o + (w & 0xFFFFu);
  return v;
}
__device__ __forceinline__ uint2 vis_ld_block(const uint8_t* bloom, uint32_t o, uint64_t pol_keep) {
  uint2 r;
  asm volatile("ld.global.cg.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(bloom + o), "l"(pol_keep));
  return r;
}
// is the slot among the block's (up to 7) offset bytes?  Bytes 0..6 hold offsets of set slots or 0xFF; byte 7 counts insertions.
__device__ __forceinline__ bool vis_test(uint2 blk, uint32_t off) {
  const uint32_t pat = off * 0x01010101u;
  // "does any byte equal off": (x - 0x01..) & ~x & 0x80.. is non-zero iff x has a zero byte (exact for the any-test)
  const uint32_t x0 = blk.x ^ pat, x1 = (blk.y ^ pat) | 0xFF000000u;  // byte 7 = count
  return ((((x0 - 0x01010101u) & ~x0) | ((x1 - 0x01010101u) & ~x1)) & 0x80808080u) != 0;
}
// a block whose count exceeds 7 spilled: its 255-bit bitmap (in the warp's bitmap area) holds the 8th and later slots
__device__ __forceinline__ bool vis_block_spilled(uint2 blk) { return blk.y >= ((kVisSlotsPerBlock + 1u) << 24); }
__device__ __forceinline__ uint32_t* vis_bitmap_word(uint8_t* bloom, const VisBase& vb, const VisSlot& v) {
  return reinterpret_cast<uint32_t*>(bloom + vb.bitmaps_o + (v.o - vb.blocks_o) * 4u) + (v.off >> 5);
}
__device__ __forceinline__ bool vis_test_spilled(uint8_t* bloom, const VisBase& vb, const VisSlot& v) {
  return (__ldcg(vis_bitmap_word(bloom, vb, v)) >> (v.off & 31u)) & 1u;
}
// set a slot (not currently set), in two steps so that nothing waits for the atomic's round trip:
// vis_reserve bumps the block's count byte with one L2 atomic and returns the old count word; vis_commit, called
// after the distance computations of the hop, stores the offset byte into the reserved position.  Reservations
// 7 and up belong to the spill bitmap (vis_spill_*): the one lane that drew number 7 clears the block's bitmap,
// then, after a warp barrier, every lane with a number >= 7 sets its bit.
__device__ __forceinline__ uint32_t vis_reserve(uint8_t* bloom, uint32_t o) {
  return atomicAdd(reinterpret_cast<uint32_t*>(bloom + o + 4), 1u << 24);
}
// the same under a predicate (bit `bit` of `mask`), without a branch; 0 when the slot is not inserted
__device__ __forceinline__ uint32_t vis_reserve_if(uint8_t* bloom, uint32_t o, uint32_t mask, uint32_t bit) {
  uint32_t r = 0;
  asm volatile("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 q, t, 0;\n\t@q atom.global.add.u32 %0, [%1+4], 16777216;\n\t}"
               : "+r"(r) : "l"(bloom + o), "r"(mask), "r"(bit) : "memory");
  return r;
}
__device__ __forceinline__ void vis_spill_clear(uint8_t* bloom, const VisBase& vb, const VisSlot& v) {
  uint4* b = reinterpret_cast<uint4*>(bloom + vb.bitmaps_o + (v.o - vb.blocks_o) * 4u);
  b[0] = make_uint4(0u, 0u, 0u, 0u);
  b[1] = make_uint4(0u, 0u, 0u, 0u);
}
__device__ __forceinline__ void vis_spill_set(uint8_t* bloom, const VisBase& vb, const VisSlot& v) {
  atomicOr(vis_bitmap_word(bloom, vb, v), 1u << (v.off & 31u));
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

// Two rows at once (same arithmetic per row as l2_row_8lane: lane t accumulates its units in ascending order).
// Loads are issued four units ahead per row, i.e. eight 16-byte requests in flight per lane (4 KB per warp), which
// is what the long rows of the Exactdistance mode (3840 B at D = 960) need to keep HBM busy.
template <typename T>
__device__ __forceinline__ void l2_two_rows_8lane(const uint8_t* va, const uint8_t* vb, const float* q_f, uint32_t units,
                                                  uint32_t t, float* da, float* db) {
  constexpr int E = Elem<T>::kPerUnit;
  float acc_a = 0.0f, acc_b = 0.0f;
  uint32_t u = t;
  for (; u + 24 < units; u += 32) {
    uint4 ra[4], rb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ra[i] = ld_nc_u4(va + (size_t)(u + 8 * i) * 16); rb[i] = ld_nc_u4(vb + (size_t)(u + 8 * i) * 16); }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float f[E];
      const float* qq = q_f + (u + 8 * i) * E;
      Elem<T>::unpack(ra[i], f);
#pragma unroll
      for (int e = 0; e < E; ++e) { const float d = __fsub_rn(f[e], qq[e]); acc_a = __fmaf_rn(d, d, acc_a); }
      Elem<T>::unpack(rb[i], f);
#pragma unroll
      for (int e = 0; e < E; ++e) { const float d = __fsub_rn(f[e], qq[e]); acc_b = __fmaf_rn(d, d, acc_b); }
    }
  }
  for (; u < units; u += 8) {
    const uint4 ra = ld_nc_u4(va + (size_t)u * 16);
    const uint4 rb = ld_nc_u4(vb + (size_t)u * 16);
    float f[E];
    Elem<T>::unpack(ra, f);
#pragma unroll
    for (int e = 0; e < E; ++e) { const float d = __fsub_rn(f[e], q_f[u * E + e]); acc_a = __fmaf_rn(d, d, acc_a); }
    Elem<T>::unpack(rb, f);
#pragma unroll
    for (int e = 0; e < E; ++e) { const float d = __fsub_rn(f[e], q_f[u * E + e]); acc_b = __fmaf_rn(d, d, acc_b); }
  }
  *da = tree8(acc_a);
  *db = tree8(acc_b);
}

// ------------------------------------------------------------------------------------------------
// shared memory: one CTA-wide read-only region (pivot table, chunk offsets) + one private region per warp
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kFull = 0xffffffffu;

struct QState {
  uint32_t lane;            // %laneid, read once per kernel (through volatile asm: ptxas otherwise re-reads the special register all over the hot loop)
  uint32_t piv_sa, qc_sa;   // shared-space byte addresses of the pivot table and of this warp's query residual (CS = 4 ADC path)
  const float* piv_s;       // [256][piv_row] pivots: the CTA-shared copy, or the global table when it does not fit (PQ modes)
  const uint32_t* coff_s;   // [n_chunks+1] chunk offsets (CTA-shared, PQ modes)
  float* q_f;        // [vec_units * E] query as fp32, zero padded (PQ modes: only during the re-rank, in place of qc)
  float* qc;         // [piv_row] query - centroid, laid out like a pivot-table row (PQ modes, during the traversal)
  uint2* w;          // worklist [w_cap], sorted by distance: .x = distance bits, .y = id | kVisitedBit if expanded
  uint32_t* n_id;    // [kListCap] filtered neighbours of this hop, unordered
  float* n_d;
  uint32_t* s_id;    // the admitted ones, sorted by (dist, id): compacted and sorted in place, i.e. the same two arrays
  float* s_d;
  float* cd;         // [cand_cap] exact distances of the re-rank (PQ modes; over the dead worklist)
  uint32_t cand_o;   // expanded-node log of this warp: SearchArgs::cand_log + cand_o, [cand_cap] ids in global memory (PQ modes; one store per hop)
  uint64_t pol_stream, pol_keep;  // L2 policies: evict-first (one-touch gathers), evict-last (visited filter)
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// CTA-shared bytes
__host__ __device__ inline size_t cta_shared_bytes(int mode, uint32_t piv_row, uint32_t n_chunks, bool piv_global) {
  if (mode == kExact) return 0;
  return (piv_global ? 0 : align_up((size_t)256 * piv_row * 4, 16)) + align_up((size_t)(n_chunks + 1) * 4, 16);
}
// floats per row of the search kernel's pivot table: 32 uniform chunks of up to 4 dimensions are zero-padded to 4 each
// (one 16-byte shared-memory load per table entry; a zero dimension adds fmaf(0, 0, acc) = acc, so the sums keep their bits)
__host__ __device__ inline uint32_t pivot_row_floats(uint32_t D, uint32_t chunk4) { return chunk4 ? 128u : D; }
// Private bytes per warp (= per resident query): [query block][worklist | neighbour lists].  PQ modes keep
// query - centroid in the query block during the traversal and the fp32 query during the re-rank, whose exact
// distances and candidate ids reuse the (then dead) worklist + list block; the expanded-node log itself lives in
// global memory — 3152 B per query at D = 128, L = 176, so that 32 queries fit next to the 128 KB pivot table.
// Exactdistance keeps the fp32 query throughout.
template <typename T>
__host__ __device__ inline size_t query_block_bytes(int mode, uint32_t piv_row, uint32_t vec_units) {
  size_t qf = align_up((size_t)vec_units * Elem<T>::kPerUnit * 4, 16);  // >= D * 4
  if (mode != kExact && qf < (size_t)piv_row * 4) qf = (size_t)piv_row * 4;  // (padded residual)
  return qf;
}
template <typename T>
__host__ __device__ inline size_t warp_private_bytes(int mode, uint32_t piv_row, uint32_t vec_units, uint32_t L, uint32_t cand_cap) {
  const size_t qf = query_block_bytes<T>(mode, piv_row, vec_units);
  size_t walk = align_up(L, 16) * 8 + (size_t)kListCap * 4 * 2;  // worklist (8 bytes per entry) + n_id/n_d
  if (mode != kExact && walk < (size_t)cand_cap * 4) walk = (size_t)cand_cap * 4;  // (the re-rank's distances; never the case for cand_cap <= L + 121)
  return align_up(qf + walk, 16);
}

template <typename T>
__device__ __forceinline__ void carve(QState& s, uint8_t* base, int mode, const SearchArgs& a, uint32_t warp) {
  size_t o = 0;
  const bool pg = a.piv_global != 0;
  s.piv_s = pg ? a.piv : (const float*)(base + o);
  s.coff_s = (const uint32_t*)(base + (pg ? 0 : align_up((size_t)256 * a.piv_row * 4, 16)));
  o += cta_shared_bytes(mode, a.piv_row, a.n_chunks, pg);
  o += (size_t)warp * warp_private_bytes<T>(mode, a.piv_row, a.vec_units, a.L, a.cand_cap);
  s.qc = (float*)(base + o);
  s.q_f = (float*)(base + o);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(base);
  s.piv_sa = sbase;
  s.qc_sa = sbase + (uint32_t)o;
  o += query_block_bytes<T>(mode, a.piv_row, a.vec_units);
  s.cd = (float*)(base + o);
  const size_t wcap = align_up(a.L, 16);
  s.w = (uint2*)(base + o); o += wcap * 8;
  s.n_id = (uint32_t*)(base + o); o += (size_t)kListCap * 4;
  s.n_d = (float*)(base + o);
  s.s_id = s.n_id;
  s.s_d = s.n_d;
  s.cand_o = (blockIdx.x * (blockDim.x >> 5) + warp) * a.cand_cap;  // (a few thousand warps x a few hundred entries)
}

// ------------------------------------------------------------------------------------------------
// stage 1 (standalone kernel only): the reference's per-query PQ distance table, written to global.
// tbl[c][k] = sum_{j in chunk c} (pivT[j][k] - (q[j] - centroid[j]))^2, j ascending, fmaf
// (populate_pqDist_par, bang_search.cu:1118-1129).  The search kernel does not materialise the table:
// adc_entry() below evaluates the same chain on demand.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void build_pq_table(const SearchArgs& a, const float* q_f, float* tbl /*[m][256]*/) {
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t c = 0; c < a.n_chunks; ++c) {
    const uint32_t j0 = a.chunk_off[c], j1 = a.chunk_off[c + 1];
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
#pragma unroll 4
    for (uint32_t j = j0; j < j1; ++j) {
      const float qc = __fsub_rn(q_f[j], __ldg(a.centroid + j));
      const float4* row = reinterpret_cast<const float4*>(a.pivT + (size_t)j * 256);
      const float4 p0 = __ldg(row + lane), p1 = __ldg(row + 32 + lane);
      const float v[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = __fsub_rn(v[i], qc); acc[i] = __fmaf_rn(d, d, acc[i]); }
    }
    float4* out = reinterpret_cast<float4*>(tbl + (size_t)c * 256);
    out[lane] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    out[32 + lane] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// One table entry on demand: the identical operation sequence as build_pq_table for (chunk c, centre `code`),
// reading the pivot row from the CTA-shared table — hence bit-identical to the reference's tbl[c][code].
// General layout (any chunk offsets, any number of chunks); the 32-uniform-chunk kernels use adc4 below.
__device__ __forceinline__ float adc_entry(const QState& s, uint32_t D, uint32_t c, uint32_t code) {
  float acc = 0.0f;
  const float* p = s.piv_s + (size_t)code * D;
  const uint32_t j0 = s.coff_s[c], j1 = s.coff_s[c + 1];
  for (uint32_t j = j0; j < j1; ++j) { const float d = __fsub_rn(p[j], s.qc[j]); acc = __fmaf_rn(d, d, acc); }
  return acc;
}

// partial ADC sum of one 32-chunk group for lane t: chunks base+t, base+t+8, base+t+16, base+t+24 (ascending).
__device__ __forceinline__ float adc_group(const QState& s, const SearchArgs& a, uint32_t word, uint32_t base, uint32_t t, float sum) {
  float e[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {  // four independent chains, then the ordered sum
    const uint32_t c = base + t + 8 * b;
    e[b] = c < a.n_chunks ? adc_entry(s, a.D, c, (word >> (8 * b)) & 0xff) : 0.0f;
  }
#pragma unroll
  for (int b = 0; b < 4; ++b)
    if (base + t + 8 * b < a.n_chunks) sum = __fadd_rn(sum, e[b]);
  return sum;
}

// ---- the CS = 4 ADC path: 32 chunks, every chunk padded to 4 dimensions -----------------------------------
// The shared pivot table is [256][32][4] floats (512 bytes per centre), the query residual one such row.  Lane t of an
// 8-lane group owns chunks t, t+8, t+16, t+24: its four residual slices (16 floats) are loaded once per hop, a table
// entry is one 16-byte shared-memory load at  table + code * 512 + chunk * 16  and the fmaf chain of build_pq_table.
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
template <int OFF>
__device__ __forceinline__ float4 lds_f4_off(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
  return v;
}
// pa + 512 * (byte B of word): one byte permute + one multiply-add
template <int B>
__device__ __forceinline__ uint32_t entry_addr(uint32_t pa, uint32_t word) {
  uint32_t code, addr;
  asm("prmt.b32 %0, %1, 0, %2;" : "=r"(code) : "r"(word), "n"(0x4440 + B));
  asm("mad.lo.u32 %0, %1, 512, %2;" : "=r"(addr) : "r"(code), "r"(pa));
  return addr;
}
__device__ __forceinline__ float adc4_chain(const float4 p, const float4 q) {
#ifndef BANG_ADC_SCALAR_SUB
  // the four subtractions as two packed fp32x2 operations (round-to-nearest per element: the same bits)
  unsigned long long p01, p23, q01, q23, d01, d23;
  asm("mov.b64 %0, {%1,%2};" : "=l"(p01) : "f"(p.x), "f"(p.y));
  asm("mov.b64 %0, {%1,%2};" : "=l"(p23) : "f"(p.z), "f"(p.w));
  asm("mov.b64 %0, {%1,%2};" : "=l"(q01) : "f"(q.x), "f"(q.y));
  asm("mov.b64 %0, {%1,%2};" : "=l"(q23) : "f"(q.z), "f"(q.w));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d01) : "l"(p01), "l"(q01));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d23) : "l"(p23), "l"(q23));
  float d0, d1, d2, d3;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d01));
  asm("mov.b64 {%0,%1}, %2;" : "=f"(d2), "=f"(d3) : "l"(d23));
#else
  const float d0 = __fsub_rn(p.x, q.x), d1 = __fsub_rn(p.y, q.y), d2 = __fsub_rn(p.z, q.z), d3 = __fsub_rn(p.w, q.w);
#endif
  float acc = __fmaf_rn(d0, d0, 0.0f);
  acc = __fmaf_rn(d1, d1, acc);
  acc = __fmaf_rn(d2, d2, acc);
  return __fmaf_rn(d3, d3, acc);
}
// lane t's share of one candidate's PQ distance: entries of chunks t, t+8, t+16, t+24 (the four bytes of `word`), summed
// in ascending chunk order from 0.0f (0.0f + e = e exactly: every entry is >= +0)
__device__ __forceinline__ float adc4_word(uint32_t pa /* table + 16 t */, uint32_t word, const float4& q0, const float4& q1, const float4& q2,
                                           const float4& q3) {
  // the four table loads are issued back to back (volatile asm keeps them in this order), then the four chains
  const float4 p0 = lds_f4_off<0>(entry_addr<0>(pa, word)), p1 = lds_f4_off<128>(entry_addr<1>(pa, word));
  const float4 p2 = lds_f4_off<256>(entry_addr<2>(pa, word)), p3 = lds_f4_off<384>(entry_addr<3>(pa, word));
  const float e0 = adc4_chain(p0, q0), e1 = adc4_chain(p1, q1), e2 = adc4_chain(p2, q2), e3 = adc4_chain(p3, q3);
  return __fadd_rn(__fadd_rn(__fadd_rn(e0, e1), e2), e3);
}

// address of lane t's code word of point `id` (32-byte code rows): one multiply-add on the 64-bit base
__device__ __forceinline__ const uint8_t* code_row(const uint8_t* cbase, uint32_t id) {
  const uint8_t* p;
  asm("mad.wide.u32 %0, %1, 32, %2;" : "=l"(p) : "r"(id), "l"(cbase));
  return p;
}

template <typename T>
__device__ __forceinline__ void load_query(const SearchArgs& a, uint32_t q, float* q_f) {
  const T* src = reinterpret_cast<const T*>(a.queries) + (size_t)q * a.q_dim;
  const uint32_t n = a.vec_units * Elem<T>::kPerUnit;
  for (uint32_t i = threadIdx.x & 31; i < n; i += 32) q_f[i] = i < a.q_dim ? (float)src[i] : 0.0f;
}
// query - centroid (populate_pqDist_par's `query[j] - centroid[j]`, bang_search.cu:1118-1129); a MIPS query is
// padded with one zero dimension (bang_search.cu:1099-1113)
template <typename T, int CS>
__device__ __forceinline__ void load_query_residual(const SearchArgs& a, uint32_t q, float* qc) {
  const T* src = reinterpret_cast<const T*>(a.queries) + (size_t)q * a.q_dim;
  if (CS == 4) {  // laid out like a row of the padded pivot table: chunk c at floats 4c .. 4c+3, zero beyond the chunk's dimensions
    for (uint32_t i = threadIdx.x & 31; i < 128u; i += 32) {
      const uint32_t e = i & 3u, j = (i >> 2) * a.chunk4 + e;
      qc[i] = e < a.chunk4 ? __fsub_rn(j < a.q_dim ? (float)src[j] : 0.0f, __ldg(a.centroid + j)) : 0.0f;
    }
    return;
  }
  for (uint32_t j = threadIdx.x & 31; j < a.D; j += 32) qc[j] = __fsub_rn(j < a.q_dim ? (float)src[j] : 0.0f, __ldg(a.centroid + j));
}

// adjacency prefetch: lane l requests neighbours 2l and 2l+1 of `node`'s HBM row (one 256-byte request per warp) and, from rows
// that carry them, the four filter-slot words of the two (one 512-byte request)
template <bool PH> struct Adj;
template <> struct Adj<false> { uint2 ids; };
template <> struct Adj<true> { uint2 ids; uint4 slots; };
template <bool PH>
__device__ __forceinline__ Adj<PH> fetch_adj(const SearchArgs& a, uint32_t node, uint64_t pol_stream, uint32_t lane) {
  Adj<PH> r;
  const uint8_t* p = row_ptr(a, node) + 8 * lane;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
               : "=r"(r.ids.x), "=r"(r.ids.y) : "l"(p), "l"(pol_stream));
  if constexpr (PH) r.slots = ld_nc_u4(p + kAdjBytes + 8 * lane, pol_stream);
  return r;
}

// Debug-only phase timers (-DBANG_PHASE_TIMERS): every lane keeps SM-clock deltas per phase; lane 0's are
// written per query to SearchArgs::st_phase[q][16].  Compiled out of the product build (empty struct).
enum { PT_SETUP = 0, PT_ADJWAIT, PT_HASH, PT_BLOOM, PT_COMPACT, PT_CODEWAIT, PT_LUT, PT_SCAN, PT_DECIDE, PT_MERGE, PT_UNVIS,
       PT_RERANK, PT_TOPK, PT_HOPS, PT_MERGES, PT_COUNT = 16 };
#ifdef BANG_PHASE_TIMERS
struct Prof {
  long long t, acc[PT_COUNT];
  __device__ __forceinline__ static long long wall_ns() { long long v; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v)); return v; }
  __device__ __forceinline__ void start() { for (int i = 0; i < PT_COUNT; ++i) acc[i] = 0; acc[PT_TOPK] = wall_ns(); t = clock64(); }
  __device__ __forceinline__ void stop() { acc[PT_COUNT - 1] = wall_ns() - acc[PT_TOPK]; }  // slot 12: start (ns), slot 15: duration (ns)
  __device__ __forceinline__ void tick(int i) { const long long n = clock64(); acc[i] += n - t; t = n; }
  __device__ __forceinline__ void count(int i) { acc[i] += 1; }
};
#else
struct Prof {
  __device__ __forceinline__ void start() {}
  __device__ __forceinline__ void stop() {}
  __device__ __forceinline__ void tick(int) {}
  __device__ __forceinline__ void count(int) {}
};
#endif

// what one lane has to store into the filter after its reservations: its four slots, the count words the atomics
// returned, and a mask of the slots it inserts
struct FilterIns { VisSlot v[4]; uint32_t r[4]; uint32_t ins; };
__device__ __forceinline__ void commit_filter(uint8_t* bloom, const VisBase& vb, const FilterIns& f) {
  bool spill = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t idx = f.r[i] >> 24;
    const bool in = (f.ins >> i) & 1u;
    if (in && idx < kVisSlotsPerBlock) bloom[f.v[i].o + idx] = (uint8_t)f.v[i].off;
    spill |= in && idx >= kVisSlotsPerBlock;
  }
  if (__any_sync(kFull, spill)) {  // rare: blocks with more than 7 slots
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (((f.ins >> i) & 1u) && (f.r[i] >> 24) == kVisSlotsPerBlock) vis_spill_clear(bloom, vb, f.v[i]);
    __syncwarp();                  // the freshly spilled blocks' bitmaps are cleared
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (((f.ins >> i) & 1u) && (f.r[i] >> 24) >= kVisSlotsPerBlock) vis_spill_set(bloom, vb, f.v[i]);
  }
}

struct VisPos { uint32_t p1, p2; };
template <int MODE>
__device__ __forceinline__ VisPos vis_pos(uint32_t id) {
  VisPos p;
  p.p1 = hash1(id);
  p.p2 = (MODE == kExact) ? p.p1 : hash2(id);  // BANG_Exactdistance tests hash 1 only (parANN.cu:1040-1066)
  return p;
}

// The medoid enters the filter after the tests of the first hop (bang_init seeds it as the first candidate, :455-462; the
// first list is [medoid] ++ adj(medoid), tested against the empty filter).  One lane, once per query; slots from set_medoid.
__device__ __forceinline__ void insert_medoid(const SearchArgs& a, uint8_t* bloom, const VisBase& vb) {
  for (uint32_t h = 0; h < a.med_slots; ++h) {
    VisSlot v;
    v.o = vb.blocks_o + a.med_blk8[h];
    v.off = a.med_off[h];
    const uint32_t idx = vis_reserve(bloom, v.o) >> 24;
    if (idx < kVisSlotsPerBlock) { bloom[v.o + idx] = (uint8_t)v.off; continue; }
    if (idx == kVisSlotsPerBlock) vis_spill_clear(bloom, vb, v);
    vis_spill_set(bloom, vb, v);
  }
}

// per-query statistics, in registers: degrees seen by this lane (summed over the warp at the end), accepted candidates
struct HopStats { uint32_t deg, npass; };

// ------------------------------------------------------------------------------------------------
// expansion of one node = stages 4a + 3.  The adjacency row is already in registers (fetch_adj).
//   filter   neighbor_filtering_new + hashFn1_d/2_d   bang_search.cu:1140-1189 — snapshot semantics: all
//            tests of a list precede all insertions (the lock-step outcome of the reference's kernel).
//            On the first hop the filter is empty, so [medoid] ++ adj(medoid) is accepted wholesale: the caller
//            puts the medoid into n_id[0] (pre = 1) and inserts it into the filter afterwards (insert_medoid).
//   PQ dist  compute_neighborDist_par                 bang_search.cu:1201-1241: lane t of an 8-lane group
//            owns chunks t, t+8, ... ascending, partials combined by the 8-lane tree.  The HBM code rows
//            are permuted at load so lane t's chunks 32g+t, 32g+8+t, 32g+16+t, 32g+24+t are one aligned
//            32-bit word: one 32-byte sector per candidate per 32 chunks, fully used.
//   exact    compute_neighborDist_par                 BANG_Exactdistance/parANN.cu:1139-1179
// Returns the number of accepted candidates (pre included); n_id/n_d hold them unordered.
// ------------------------------------------------------------------------------------------------
template <typename T, int MODE, int CS, bool PH>
__device__ __forceinline__ uint32_t expand(const SearchArgs& a, const QState& s, uint8_t* bloom, const VisBase& vb, const Adj<PH>& nb, uint32_t pre,
                                           HopStats& st, Prof& pf) {
  const uint32_t lane = s.lane, lt = (1u << lane) - 1u;
  const uint32_t id0 = nb.ids.x, id1 = nb.ids.y;
  const bool v0 = id0 != kNoNbr, v1 = id1 != kNoNbr;
#ifdef BANG_PHASE_TIMERS
  if (__any_sync(kFull, id0 == 0x12345678u && id1 == 0x9abcdef0u)) printf("");  // forces the adjacency load to complete here
  pf.tick(PT_ADJWAIT);
#endif
  if (MODE != kExact && a.code_prefetch) {  // codes of every neighbour towards L2 while the filter is being consulted
    if (v0) prefetch_l2(a.codes + (size_t)id0 * a.code_stride);
    if (v1) prefetch_l2(a.codes + (size_t)id1 * a.code_stride);
  }
  // the four slots of this lane: (id0, hash 1), (id0, hash 2), (id1, hash 1), (id1, hash 2).  A padding id hashes to a
  // valid slot as well, so the loads and tests need no guard; v0 / v1 enter at the accept decision.
  FilterIns fi;
  bool same0, same1;  // the id's two hashes name one slot
  if constexpr (PH) {
    fi.v[0] = vis_slot_from_word(vb, nb.slots.x); fi.v[1] = vis_slot_from_word(vb, nb.slots.y);
    fi.v[2] = vis_slot_from_word(vb, nb.slots.z); fi.v[3] = vis_slot_from_word(vb, nb.slots.w);
    same0 = nb.slots.x == nb.slots.y; same1 = nb.slots.z == nb.slots.w;
  } else {
    const VisPos p0 = vis_pos<MODE>(id0), p1 = vis_pos<MODE>(id1);
    fi.v[0] = vis_slot(vb, p0.p1); fi.v[1] = vis_slot(vb, p0.p2); fi.v[2] = vis_slot(vb, p1.p1); fi.v[3] = vis_slot(vb, p1.p2);
    same0 = p0.p2 == p0.p1; same1 = p1.p2 == p1.p1;
  }
#ifdef BANG_PHASE_TIMERS
  if (__any_sync(kFull, fi.v[0].o == 0xFFFFFFFFu)) printf("");
  pf.tick(PT_HASH);
#endif
  uint32_t nf;  // bit i: slot i is not set in the filter (Exactdistance: bits 0 and 2 only)
  {
    const uint2 b01 = vis_ld_block(bloom, fi.v[0].o, s.pol_keep), b11 = vis_ld_block(bloom, fi.v[2].o, s.pol_keep);
    uint2 b02 = b01, b12 = b11;
    if (MODE != kExact) { b02 = vis_ld_block(bloom, fi.v[1].o, s.pol_keep); b12 = vis_ld_block(bloom, fi.v[3].o, s.pol_keep); }
    const bool f01 = vis_test(b01, fi.v[0].off), f11 = vis_test(b11, fi.v[2].off);
    const bool f02 = MODE == kExact || vis_test(b02, fi.v[1].off), f12 = MODE == kExact || vis_test(b12, fi.v[3].off);
    const bool sp01 = !f01 && vis_block_spilled(b01), sp02 = !f02 && vis_block_spilled(b02);
    const bool sp11 = !f11 && vis_block_spilled(b11), sp12 = !f12 && vis_block_spilled(b12);
    if (__builtin_expect(__any_sync(kFull, sp01 || sp02 || sp11 || sp12), 0)) {  // rare: a block with more than 7 slots, see its bitmap
      nf = ((f01 || (sp01 && vis_test_spilled(bloom, vb, fi.v[0]))) ? 0u : 1u) | ((f02 || (sp02 && vis_test_spilled(bloom, vb, fi.v[1]))) ? 0u : 2u) |
           ((f11 || (sp11 && vis_test_spilled(bloom, vb, fi.v[2]))) ? 0u : 4u) | ((f12 || (sp12 && vis_test_spilled(bloom, vb, fi.v[3]))) ? 0u : 8u);
    } else {
      nf = (f01 ? 0u : 1u) | (f02 ? 0u : 2u) | (f11 ? 0u : 4u) | (f12 ? 0u : 8u);
    }
  }
  // an id is accepted unless all its slots are set; its unset slots are inserted (one id whose two hashes coincide sets the slot once)
  const bool acc0 = v0 && (nf & 3u) != 0, acc1 = v1 && (nf & 12u) != 0;
  uint32_t ins = (acc0 ? (nf & 3u) : 0u) | (acc1 ? (nf & 12u) : 0u);
  if (MODE != kExact) {
    if (same0) ins &= ~2u;
    if (same1) ins &= ~8u;
  }
  fi.ins = ins;
  __syncwarp();  // every test precedes every insertion
#ifdef BANG_PHASE_TIMERS
  if (__any_sync(kFull, acc0 && id0 == 0x12345678u)) printf("");
  pf.tick(PT_BLOOM);
#endif
#pragma unroll
  for (int i = 0; i < 4; ++i) fi.r[i] = vis_reserve_if(bloom, fi.v[i].o, ins, 1u << i);
  const uint32_t m0 = __ballot_sync(kFull, acc0), m1 = __ballot_sync(kFull, acc1);
  const uint32_t c0 = pre + __popc(m0);
  const uint32_t n = c0 + __popc(m1);
  if (acc0) s.n_id[pre + __popc(m0 & lt)] = id0;
  if (acc1) s.n_id[c0 + __popc(m1 & lt)] = id1;
  st.deg += (v0 ? 1u : 0u) + (v1 ? 1u : 0u);
  st.npass += n;
  __syncwarp();
  pf.tick(PT_COMPACT);
  // The reserved filter bytes are stored after the distance computations of the hop: the atomics that hand out the
  // byte positions have long returned by then, so nothing waits for their round trip (storing them while the first
  // code words travel instead measured 8 % slower on the C2 shape, profiles/r2r_reorder_variants_ab.log).
  const uint32_t t = lane & 7, g = lane >> 3;  // 4 candidates per pass
  if (MODE == kExact) {
    for (uint32_t k0 = 0; k0 < n; k0 += 8) {  // two rows in flight per lane group
      const uint32_t ka = k0 + g, kb = k0 + 4 + g;
      const uint32_t ca = ka < n ? s.n_id[ka] : a.medoid, cb = kb < n ? s.n_id[kb] : a.medoid;
      float da, db;
      l2_two_rows_8lane<T>(row_ptr(a, ca) + adj_bytes<PH>(), row_ptr(a, cb) + adj_bytes<PH>(), s.q_f, a.vec_units, t, &da, &db);
      if (t == 0 && ka < n) s.n_d[ka] = da;
      if (t == 0 && kb < n) s.n_d[kb] = db;
    }
  } else if (CS > 0) {
    // 32 uniform chunks (C2 / C4 / C5: D = 128 or 96, 32 bytes per vector): lane t's chunks t, t+8, t+16, t+24 are one 32-bit
    // code word.  The code words of up to 16 candidates are requested at once; the table entries are evaluated four
    // candidates (one per 8-lane group) at a time.
    const uint32_t pa = s.piv_sa + 16u * t, qa = s.qc_sa + 16u * t;
    const float4 q0 = lds_f4_off<0>(qa), q1 = lds_f4_off<128>(qa), q2 = lds_f4_off<256>(qa), q3 = lds_f4_off<384>(qa);
    const uint8_t* cbase;  // codes + 4 t (32 chunks: 32-byte code rows); opaque, so that it is formed once per hop and not once per load
    asm volatile("add.u64 %0, %1, %2;" : "=l"(cbase) : "l"(a.codes), "l"((unsigned long long)(4u * t)));
    // the code words of the candidates k0 + g, + 4, + 8, + 12 of this lane group (beyond n: 0; the ids read there are stale
    // entries of the list block, never used)
    auto load_words = [&](uint32_t k0, uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3) {
      const uint32_t kg = k0 + g;
      const uint32_t* np = s.n_id + kg;
      const uint32_t i0 = np[0], i1 = np[4], i2 = np[8], i3 = np[12];
      x0 = x1 = x2 = x3 = 0;
      if (kg < n) x0 = ld_nc_u32(code_row(cbase, i0), s.pol_stream);
      if (kg + 4 < n) x1 = ld_nc_u32(code_row(cbase, i1), s.pol_stream);
      if (kg + 8 < n) x2 = ld_nc_u32(code_row(cbase, i2), s.pol_stream);
      if (kg + 12 < n) x3 = ld_nc_u32(code_row(cbase, i3), s.pol_stream);
    };
    uint32_t w0, w1, w2, w3;
    load_words(0, w0, w1, w2, w3);
    for (uint32_t k0 = 0; k0 < n; k0 += 16) {
      // the next 16 candidates' words are requested before this batch is evaluated (lists longer than 16 are the rule on the
      // DEEP shape at small L, where the wait for the second batch was 18 % of the stall samples: profiles/r2_hop_diet.md §7)
      uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;
      if (k0 + 16 < n) load_words(k0 + 16, x0, x1, x2, x3);
#ifdef BANG_PHASE_TIMERS
      if (__any_sync(kFull, (w0 ^ w1 ^ w2 ^ w3) == 0x12345678u)) printf("");
      pf.tick(PT_CODEWAIT);
#endif
      const uint32_t kend = min(n, k0 + 16);
#pragma unroll 1
      for (uint32_t kb = k0; kb < kend; kb += 4) {  // (not unrolled: the kernel's hot loop has to stay inside the instruction cache; two
        const float sum = tree8(adc4_word(pa, w0, q0, q1, q2, q3));  // candidates per pass measured +-1 %, profiles/r2r_reorder_variants_ab.log)
        if (t == 0 && kb + g < n) s.n_d[kb + g] = sum;
        w0 = w1; w1 = w2; w2 = w3;
      }
      w0 = x0; w1 = x1; w2 = x2; w3 = x3;
    }
  } else {
    // any chunk layout (uneven chunks, m != 32: the reference's SIFT1B build has 74, SIFT10K 128): 32 chunks per group
    const uint32_t groups = (a.n_chunks + 31) >> 5;
    for (uint32_t k0 = 0; k0 < n; k0 += 4) {
      const uint32_t k = k0 + g;
      const uint8_t* row = a.codes + (size_t)(k < n ? s.n_id[k] : 0u) * a.code_stride + 4 * t;
      float sum = 0.0f;
      for (uint32_t gg = 0; gg < groups; gg += 2) {
        const uint32_t wa = ld_nc_u32(row + gg * 32, s.pol_stream);
        const uint32_t wb = (gg + 1 < groups) ? ld_nc_u32(row + (gg + 1) * 32, s.pol_stream) : 0u;
        sum = adc_group(s, a, wa, gg * 32, t, sum);
        sum = adc_group(s, a, wb, (gg + 1) * 32, t, sum);
      }
      sum = tree8(sum);
      if (t == 0 && k < n) s.n_d[k] = sum;
    }
  }
  commit_filter(bloom, vb, fi);
  __syncwarp();
  pf.tick(PT_LUT);
  return n;
}

__device__ __forceinline__ void write_stats(const SearchArgs& a, const QState& s, uint32_t q, const HopStats& st) {
  const uint32_t deg = __reduce_add_sync(kFull, st.deg);
  if (s.lane == 0) {
    if (a.st_sumdeg) a.st_sumdeg[q] = deg;
    if (a.st_npass) a.st_npass[q] = st.npass;
  }
}

// (dist, id)-minimum of the unsorted neighbour list (optionally skipping the medoid), the number of
// entries closer than `maxd`, and the medoid's distance if present.  Distances are >= +0, so their bit
// patterns order like the floats and redux.sync (min/add over the warp) does the reductions.
struct Best { float d; uint32_t id; uint32_t below; float med_d; bool med_in; };
__device__ __forceinline__ Best scan_neighbours(const QState& s, uint32_t n, uint32_t medoid, bool skip_medoid, float maxd) {
  const uint32_t lane = s.lane;
  uint32_t bd = 0x7F7FFFFFu /* FLT_MAX */, bid = kNone, below = 0, md = 0;
  bool mi = false;
  // a list holds at most 65 entries: lane, lane + 32 and (lane 0) entry 64; most hops have fewer than 32
  auto take = [&](uint32_t i, bool skip) {
    const float d = s.n_d[i];
    const uint32_t id = s.n_id[i], db = __float_as_uint(d);
    below += d < maxd ? 1u : 0u;
    if (skip && id == medoid) { mi = true; md = db; }
    else if (db < bd || (db == bd && id < bid)) { bd = db; bid = id; }
  };
  if (skip_medoid) {  // the first hop: the list starts with the medoid itself, which is not a candidate for expansion
    if (lane < n) take(lane, true);
    if (lane + 32 < n) take(lane + 32, true);
    if (lane + 64 < n) take(lane + 64, true);
  } else {            // every later hop (med_in / med_d are only read on the first)
    if (lane < n) take(lane, false);
    if (n > 32) {
      if (lane + 32 < n) take(lane + 32, false);
      if (lane + 64 < n) take(lane + 64, false);
    }
  }
  Best b;
  const uint32_t dmin = __reduce_min_sync(kFull, bd);
  b.id = __reduce_min_sync(kFull, bd == dmin ? bid : kNone);
  b.d = __uint_as_float(dmin);
  b.below = __reduce_add_sync(kFull, below);
  const uint32_t mm = __ballot_sync(kFull, mi);
  b.med_in = mm != 0;
  b.med_d = __uint_as_float(__shfl_sync(kFull, md, mm ? __ffs(mm) - 1 : 0));
  return b;
}

// first unvisited worklist entry at or after `start`
__device__ __forceinline__ uint32_t scan_unvisited(const QState& s, uint32_t start, uint32_t ws) {
  const uint32_t lane = s.lane;
  for (uint32_t b0 = start; b0 < ws; b0 += 32) {
    const uint32_t j = b0 + lane;
    const uint32_t m = __ballot_sync(kFull, j < ws && (s.w[j].y & kVisitedBit) == 0);
    if (m) return b0 + (uint32_t)__ffs(m) - 1u;
  }
  return kNone;
}

// number of new entries the reference's merge admits (nbrsBound, bang_search.cu:1651-1656)
__device__ __forceinline__ uint32_t admit_count(uint32_t below, uint32_t n, uint32_t ws, uint32_t L) {
  return max(min(below, min(L, n)), min(L - ws, n));
}

// The nb admitted entries (the nb smallest of the n new ones by (dist, id)), sorted, into s_id/s_d.
// Common case (worklist full): the admitted set is exactly the entries closer than the current tail, a
// handful per hop — compacted with ballots and ranked through shuffles.  Otherwise: rank among all n.
__device__ __forceinline__ void select_admitted(const QState& s, uint32_t n, uint32_t nb, uint32_t below, float maxd) {
  const uint32_t lane = s.lane, lt = (1u << lane) - 1u;
  if (nb == below && nb <= 32) {
    uint32_t base = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
      const uint32_t i = i0 + lane;
      float d = 0.0f;
      uint32_t id = 0;
      if (i < n) { d = s.n_d[i]; id = s.n_id[i]; }
      const bool in = i < n && d < maxd;
      const uint32_t m = __ballot_sync(kFull, in);  // (in place: every lane has read its entry before any lane writes)
      if (in) { const uint32_t p = base + __popc(m & lt); s.s_d[p] = d; s.s_id[p] = id; }
      base += __popc(m);
    }
    __syncwarp();
    uint32_t kd = 0xFFFFFFFFu, ki = kNone;
    if (lane < nb) { kd = __float_as_uint(s.s_d[lane]); ki = s.s_id[lane]; }
    uint32_t r = 0;
    for (uint32_t j = 0; j < nb; ++j) {
      const uint32_t od = __shfl_sync(kFull, kd, j), oi = __shfl_sync(kFull, ki, j);
      r += (od < kd || (od == kd && oi < ki)) ? 1u : 0u;
    }
    __syncwarp();
    if (lane < nb) { s.s_d[r] = __uint_as_float(kd); s.s_id[r] = ki; }
  } else {
    // rank of every entry among all n by the 64-bit key (distance bits, id) — distances are >= +0, so their bit patterns
    // order like the floats.  The list is read four entries at a time (16-byte broadcast loads) against the lane's two
    // entries (i = lane, lane + 32); the tail of the last group of four is padded with +inf keys.
    if (lane < 3 && n + lane < ((n + 3u) & ~3u)) s.n_d[n + lane] = __uint_as_float(0x7F800000u);
    __syncwarp();
    const uint32_t ia = lane, ib = lane + 32;
    const unsigned long long ka = ia < n ? ((unsigned long long)__float_as_uint(s.n_d[ia]) << 32 | s.n_id[ia]) : ~0ull;
    const unsigned long long kb = ib < n ? ((unsigned long long)__float_as_uint(s.n_d[ib]) << 32 | s.n_id[ib]) : ~0ull;
    uint32_t ra = 0, rb = 0;
    for (uint32_t j0 = 0; j0 < n; j0 += 4) {
      const uint4 dj = *reinterpret_cast<const uint4*>(s.n_d + j0), ij = *reinterpret_cast<const uint4*>(s.n_id + j0);
      const unsigned long long k0 = (unsigned long long)dj.x << 32 | ij.x, k1 = (unsigned long long)dj.y << 32 | ij.y;
      const unsigned long long k2 = (unsigned long long)dj.z << 32 | ij.z, k3 = (unsigned long long)dj.w << 32 | ij.w;
      ra += (k0 < ka ? 1u : 0u) + (k1 < ka ? 1u : 0u) + (k2 < ka ? 1u : 0u) + (k3 < ka ? 1u : 0u);
      rb += (k0 < kb ? 1u : 0u) + (k1 < kb ? 1u : 0u) + (k2 < kb ? 1u : 0u) + (k3 < kb ? 1u : 0u);
    }
    float d64 = 0.0f;
    uint32_t id64 = 0, r64 = kNone;
    if (n > 64 && lane == 0) {  // (a 65th entry: the medoid's list on the first hop)
      d64 = s.n_d[64];
      id64 = s.n_id[64];
      r64 = 0;
      for (uint32_t j = 0; j < n; ++j) r64 += key_less(s.n_d[j], s.n_id[j], d64, id64) ? 1u : 0u;
    }
    __syncwarp();  // the sorted list replaces the unsorted one in place: every rank is known before the first store
    if (ia < n && ra < nb) { s.s_d[ra] = __uint_as_float((uint32_t)(ka >> 32)); s.s_id[ra] = (uint32_t)ka; }
    if (ib < n && rb < nb) { s.s_d[rb] = __uint_as_float((uint32_t)(kb >> 32)); s.s_id[rb] = (uint32_t)kb; }
    if (r64 < nb) { s.s_d[r64] = d64; s.s_id[r64] = id64; }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// stage 4b: (dist, id) sort of the new neighbours + merge into the worklist, in place.
// compute_BestLSets_par_sort_msort + compute_BestLSets_par_merge (bang_search.cu:1533-1585, 1605-1715):
// only the `nb` closest new entries take part (nb as the reference computes nbrsBound); a new entry
// goes before old entries of equal distance; the list is truncated to L.  New entries are unvisited,
// except `flag_id` (the node just chosen for expansion / the reference's d_mark) and, in the first
// merge, the medoid.  Old entries move to index + upper_bound(new, d), 32 at a time from the tail and
// only from the first insertion point on, so nothing is overwritten before it is read.
// Returns the new size; *pos0 = position of the closest new entry.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float w_dist(const QState& s, uint32_t j) { return __uint_as_float(s.w[j].x); }
__device__ __forceinline__ uint32_t merge_worklist(const SearchArgs& a, const QState& s, uint32_t n, uint32_t nb, uint32_t below,
                                                   float maxd, uint32_t ws, bool first, uint32_t flag_id, uint32_t* pos0) {
  const uint32_t lane = s.lane;
  select_admitted(s, n, nb, first ? kNone : below, maxd);
  if (first) {  // iter == 1 branch (:1636-1646): the worklist is the head of the sorted list
    for (uint32_t i = lane; i < nb; i += 32) {
      const uint32_t id = s.s_id[i];
      s.w[i] = make_uint2(__float_as_uint(s.s_d[i]), id | ((id == a.medoid || id == flag_id) ? kVisitedBit : 0u));
    }
    __syncwarp();
    *pos0 = 0;
    return nb;
  }
  const uint32_t newsize = min(ws + nb, a.L);
  if (nb <= 4) {
    // The common case once the worklist is full: a handful of new entries.  Their distances sit in registers of every
    // lane (+inf beyond nb); one pass over the old entries from the tail, 32 at a time, moves each by the number of new
    // entries at or below it (new before old on ties) and counts, per new entry, the old entries that stay in front of
    // it (= its lower bound) with one ballot each — no dependent shared-memory searches.  The pass stops at the first
    // chunk that lies entirely below the closest new entry.
    const float inf = __uint_as_float(0x7F800000u);
    const float4 sd = *reinterpret_cast<const float4*>(s.s_d);  // (the lists are 16-byte aligned)
    const float nd0 = sd.x, nd1 = nb > 1 ? sd.y : inf, nd2 = nb > 2 ? sd.z : inf, nd3 = nb > 3 ? sd.w : inf;
    uint32_t lb0 = 0, lb1 = 0, lb2 = 0, lb3 = 0;
    for (int c = (int)((ws - 1) >> 5); c >= 0; --c) {
      const uint32_t j = (uint32_t)c * 32 + lane;
      const bool live = j < ws;
      uint2 e = make_uint2(0x7F800000u, 0u);
      if (live) e = s.w[j];
      const float d = __uint_as_float(e.x);
      const bool l0 = d < nd0, l1 = d < nd1, l2 = d < nd2, l3 = d < nd3;
      const uint32_t b0 = __ballot_sync(kFull, l0);
      lb0 += __popc(b0);
      lb1 += __popc(__ballot_sync(kFull, l1));
      lb2 += __popc(__ballot_sync(kFull, l2));
      lb3 += __popc(__ballot_sync(kFull, l3));
      const uint32_t pos = j + (l0 ? 0u : 1u) + (l1 ? 0u : 1u) + (l2 ? 0u : 1u) + (l3 ? 0u : 1u);  // (dead lanes: >= ws + 4, never stored... see below)
      __syncwarp();
      if (live && pos < newsize) s.w[pos] = e;
      __syncwarp();
      if (b0 == kFull) {  // everything from here down is closer than every new entry and stays where it is
        const uint32_t rest = (uint32_t)c * 32;
        lb0 += rest; lb1 += rest; lb2 += rest; lb3 += rest;
        break;
      }
    }
    if (lane < nb) {
      const uint32_t lb = lane == 0 ? lb0 : (lane == 1 ? lb1 : (lane == 2 ? lb2 : lb3));
      const uint32_t np = lane + lb;
      if (np < newsize) {
        const uint32_t id = s.s_id[lane];
        s.w[np] = make_uint2(__float_as_uint(s.s_d[lane]), id | (id == flag_id ? kVisitedBit : 0u));
      }
    }
    __syncwarp();
    *pos0 = lb0;
    return newsize;
  }
  // general case (the worklist is still filling up): position = index + lower_bound(W, d)  (new before old on ties);
  // nb <= 65 -> at most 3 per lane
  uint32_t npos[3];
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    const uint32_t i = lane + 32 * e;
    npos[e] = kNone;
    if (i < nb) {
      const float d = s.s_d[i];
      uint32_t lo = 0, hi = ws;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (d <= w_dist(s, mid)) hi = mid; else lo = mid + 1;
      }
      npos[e] = lo + i;
    }
  }
  const uint32_t p0 = __shfl_sync(kFull, npos[0], 0);
  __syncwarp();
  for (int c = (int)((ws - 1) >> 5); c >= (int)(p0 >> 5); --c) {  // old entries at or after the insertion point, tail first
    const uint32_t j = (uint32_t)c * 32 + lane;
    const bool live = j < ws && j >= p0;
    uint2 e = make_uint2(0u, 0u);
    uint32_t pos = kNone;
    if (live) {
      e = s.w[j];
      const float d = __uint_as_float(e.x);
      uint32_t lo = 0, hi = nb;
      while (lo < hi) {  // upper_bound over the admitted new entries
        const uint32_t mid = (lo + hi) >> 1;
        if (d >= s.s_d[mid]) lo = mid + 1; else hi = mid;
      }
      pos = j + lo;
    }
    __syncwarp();
    if (pos < newsize) s.w[pos] = e;
    __syncwarp();
  }
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    const uint32_t i = lane + 32 * e;
    if (npos[e] < newsize) {
      const uint32_t id = s.s_id[i];
      s.w[npos[e]] = make_uint2(__float_as_uint(s.s_d[i]), id | (id == flag_id ? kVisitedBit : 0u));
    }
  }
  __syncwarp();
  *pos0 = p0;
  return newsize;
}

// ------------------------------------------------------------------------------------------------
// stage 5: exact distances of the expanded nodes (compute_L2Dist, bang_search.cu:1254-1299) with
// coalesced 16-byte loads, 8 lanes per row, two rows in flight per lane group; then the k smallest by
// (exact distance, id) (compute_NearestNeighbours, :1312-1368) by k rounds of warp-wide min extraction.
// ------------------------------------------------------------------------------------------------
template <typename T, bool PH>
__device__ __forceinline__ void rerank_and_write(const SearchArgs& a, const QState& s, uint32_t q, uint32_t n) {
  const uint32_t lane = s.lane, t = lane & 7, g = lane >> 3;
  float* cd = s.cd;       // the worklist block is dead by now: exact distances + a copy of the log go there,
  const uint32_t* cid = a.cand_log + s.cand_o;  // the expanded-node log in global memory (L2): ids are re-read from there
  __syncwarp();
  load_query<T>(a, q, s.q_f);
  __syncwarp();
  for (uint32_t b0 = 0; b0 < n; b0 += 8) {
    const uint32_t i0 = b0 + g, i1 = b0 + 4 + g;
    const uint32_t id0 = i0 < n ? __ldcg(cid + i0) : a.medoid, id1 = i1 < n ? __ldcg(cid + i1) : a.medoid;
    float d0, d1;
    l2_two_rows_8lane<T>(row_ptr(a, id0) + adj_bytes<PH>(), row_ptr(a, id1) + adj_bytes<PH>(), s.q_f, a.vec_units, t, &d0, &d1);
    if (t == 0 && i0 < n) cd[i0] = d0;
    if (t == 0 && i1 < n) cd[i1] = d1;
  }
  __syncwarp();
  uint32_t last_d = 0, last_id = 0;
  bool have_last = false;
  for (uint32_t r = 0; r < a.k; ++r) {
    uint32_t bd = 0xFFFFFFFFu, bid = kNone;
    for (uint32_t i = lane; i < n; i += 32) {
      const uint32_t db = __float_as_uint(cd[i]), id = __ldcg(cid + i);
      const bool after = !have_last || db > last_d || (db == last_d && id > last_id);
      if (after && (db < bd || (db == bd && id < bid))) { bd = db; bid = id; }
    }
    const uint32_t dmin = __reduce_min_sync(kFull, bd);
    const uint32_t idmin = __reduce_min_sync(kFull, bd == dmin ? bid : kNone);
    const bool found = idmin != kNone;
    if (lane == 0) {
      a.out_ids[(size_t)q * a.k + r] = found ? (uint64_t)idmin : 0xFFFFFFFFull;
      a.out_dists[(size_t)q * a.k + r] = found ? __uint_as_float(dmin) : 3.402823466e+38f;
    }
    if (!found) { last_d = 0xFFFFFFFFu; last_id = kNone; }  // nothing left: the remaining ranks are filler too
    else { last_d = dmin; last_id = idmin; }
    have_last = true;
  }
}


// ------------------------------------------------------------------------------------------------
// the kernel: blockDim.x = 32 * (query warps per CTA)
// ------------------------------------------------------------------------------------------------
// WPC = query warps per CTA the kernel is compiled for.  PQ modes: one CTA per SM around one pivot table, compiled
// for up to 32 warps (64 registers per thread) and for up to 16 (128 registers; large D, where few queries fit) —
// the host runs as many warps as shared memory allows for the index's D and the search's L (launch_geometry),
// because the search is a chain of dependent memory round trips and throughput follows the number of resident
// queries (profiles/r1_concurrency.md, r2_concurrency.md).  Exactdistance: 2 CTAs of 16 warps per SM.
template <typename T, int MODE, int CS, int WPC, bool PH = false>
__global__ void __launch_bounds__(WPC * 32, (MODE == kExact) ? 2 : 1) bang_search_kernel(const SearchArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t warps = blockDim.x >> 5;
  uint32_t lane;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
  const uint32_t warp = __shfl_sync(kFull, threadIdx.x >> 5, 0);  // (through a shuffle: the compiler then knows it is warp-uniform)
  if (MODE != kExact) {
    // the pivot table (unless it stays in global memory) and the chunk offsets, once per CTA, shared by all its query warps
    uint8_t* p = smem_raw;
    if (!a.piv_global) {
      float4* dst = reinterpret_cast<float4*>(smem_raw);
      const float4* src = reinterpret_cast<const float4*>(a.piv);
      const uint32_t n4 = 256u * a.piv_row / 4u;  // piv_row*256 floats; 256*piv_row*4 bytes is a multiple of 16
      for (uint32_t i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
      p += align_up((size_t)256 * a.piv_row * 4, 16);
    }
    uint32_t* coff = reinterpret_cast<uint32_t*>(p);
    for (uint32_t i = threadIdx.x; i <= a.n_chunks; i += blockDim.x) coff[i] = a.chunk_off[i];
    __syncthreads();  // the only CTA barrier; from here on the warps never meet again
  }
  QState s;
  s.lane = lane;
  carve<T>(s, smem_raw, MODE, a, warp);
  const uint64_t pol_stream = l2_policy_evict_first();
  s.pol_stream = pol_stream;
  s.pol_keep = l2_policy_evict_last();
  // [block areas of all warps of the grid][spill bitmap areas of all warps of the grid]
  uint8_t* const bloom = reinterpret_cast<uint8_t*>(a.bloom);
  VisBase vb;
  {
    const uint32_t gw = blockIdx.x * warps + warp;
    vb.blocks_o = gw * kVisBlockBytes;
    vb.bitmaps_o = gridDim.x * warps * kVisBlockBytes + gw * kVisBitmapBytes;
  }

  for (;;) {
    uint32_t q = 0;
    if (lane == 0) q = atomicAdd(a.counter, 1u);
    q = __shfl_sync(kFull, q, 0);
    if (q >= a.Q) break;

    Prof pf;
    pf.start();
    // ---- per-query setup: query -> smem, bloom filter cleared ----
    Adj<PH> my_nb = fetch_adj<PH>(a, a.medoid, pol_stream, lane);  // the first hop's adjacency row travels during the setup
    __syncwarp();
    if (MODE == kExact) load_query<T>(a, q, s.q_f);
    else load_query_residual<T, CS>(a, q, s.qc);
    {  // empty filter: every offset byte 0xFF, count 0 (two 8-byte blocks per store; the padding blocks are never read)
      uint4* b4 = reinterpret_cast<uint4*>(bloom + vb.blocks_o);
      for (uint32_t i = lane; i < (kVisBlocks + 1) / 2; i += 32) b4[i] = make_uint4(0xFFFFFFFFu, 0x00FFFFFFu, 0xFFFFFFFFu, 0x00FFFFFFu);
    }
    if (MODE != kExact && lane == 0) a.cand_log[s.cand_o] = a.medoid;  // bang_init: the medoid is every query's first candidate (:455-462)
    __syncwarp();
    __threadfence_block();  // the cleared filter is ordered before this query's tests and insertions
    pf.tick(PT_SETUP);

    uint32_t ws = 0, fu = kNone, ncand = 1, iter = 1, pos0 = 0;
    HopStats st{0u, 0u};
    auto log_parent = [&](uint32_t node) {
      if (lane == 0) {
        if (MODE != kExact && ncand < a.cand_cap) a.cand_log[s.cand_o + ncand] = node;
        if (MODE == kExact && a.dump_ids && ncand < a.dump_stride) a.dump_ids[(size_t)q * a.dump_stride + ncand] = node;  // (index builder)
      }
      if (ncand < a.cand_cap) ++ncand;
    };

    if (MODE == kBase) {
      // ---- BANG_Base (A.1, A.2): seed, then { merge(previous) ; expand(parent) ; compute_parent2 } ----
      if (lane == 0) s.n_id[0] = a.medoid;
      uint32_t n = expand<T, MODE, CS, PH>(a, s, bloom, vb, my_nb, 1u, st, pf);
      if (lane == 0) insert_medoid(a, bloom, vb);
      Best b = scan_neighbours(s, n, a.medoid, true, 0.0f);
      bool have = b.id != kNone;  // compute_parent1 (:1464-1521): closest seeded neighbour, medoid excluded
      uint32_t parent = b.id, mark = have ? b.id : 0x01010101u;
      if (have) log_parent(parent);
      uint32_t pend_n = n, pend_nb = min(n, a.L), pend_below = 0, scan_from = 0;
      float pend_maxd = 0.0f;
      while (have || pend_n > 0) {
        if (have) my_nb = fetch_adj<PH>(a, parent, pol_stream, lane);  // in flight during the merge
        pf.tick(PT_DECIDE);
        if (pend_n > 0 && pend_nb > 0) {          // sort + merge of the previous neighbours (:726,:738), mark (:1711-1714)
          ws = merge_worklist(a, s, pend_n, pend_nb, pend_below, pend_maxd, ws, iter == 1, mark, &pos0);
          scan_from = min(scan_from, pos0);
          pf.count(PT_MERGES);
        }
        pf.tick(PT_MERGE);
        fu = scan_unvisited(s, scan_from, ws);
        scan_from = fu == kNone ? ws : fu;
        pf.tick(PT_UNVIS);
        pf.count(PT_HOPS);
        n = 0;
        if (have) n = expand<T, MODE, CS, PH>(a, s, bloom, vb, my_nb, 0u, st, pf);
        ++iter;
        // compute_parent2 (:1403-1458)
        const float maxd = ws > 0 ? w_dist(s, ws - 1) : 0.0f;
        b = scan_neighbours(s, n, a.medoid, true, maxd);
        pf.tick(PT_SCAN);
        const bool hasx = b.id != kNone;
        have = false;
        if (fu != kNone) {
          have = true;
          if (hasx && b.d < w_dist(s, fu)) { parent = b.id; mark = b.id; }
          else { parent = s.w[fu].y; if (lane == 0) s.w[fu].y = parent | kVisitedBit; scan_from = fu + 1; }  // (fu is unvisited: no flag in the id word)
        } else if (ws > 0 && hasx && b.d < maxd) {
          have = true; parent = b.id; mark = b.id;
        }
        if (have) log_parent(parent);
        pend_n = n;
        pend_below = b.below;
        pend_maxd = maxd;
        pend_nb = n ? admit_count(b.below, n, ws, a.L) : 0u;
        if (iter == a.max_iter - 1) break;
      }
      write_stats(a, s, q, st);
      pf.tick(PT_DECIDE);
      rerank_and_write<T, PH>(a, s, q, ncand);
      pf.tick(PT_RERANK);
    } else {
      // ---- BANG_Inmemory / BANG_Exactdistance (A.2', A.2''): { expand(parent) ; merge ; first unvisited } ----
      // The first unvisited entry after the merge is decided before it: the closest new entry if it is
      // admitted and not farther than the first unvisited old entry (new goes before old on ties).
      uint32_t parent = a.medoid;
      for (;;) {
        const bool first = iter == 1;
        if (first && lane == 0) s.n_id[0] = a.medoid;
        const uint32_t n = expand<T, MODE, CS, PH>(a, s, bloom, vb, my_nb, first ? 1u : 0u, st, pf);
        if (first && lane == 0) insert_medoid(a, bloom, vb);
        // Exactdistance: a hop whose neighbours are all filtered out ends the query — what the reference's fused
        // kernel does when built for sm_100a (it scans the worklist up to a size it only sets when there are new
        // neighbours, BANG_Exactdistance/parANN.cu:1593,1600,1671; pinned by tests/golden/ref_forks_golden.npz).
        if (MODE == kExact && a.stop_on_empty_hop && !first && n == 0) break;
        const float maxd = ws > 0 ? w_dist(s, ws - 1) : 0.0f;
        const Best b = scan_neighbours(s, n, a.medoid, first, maxd);
        pf.tick(PT_SCAN);
        pf.count(PT_HOPS);
        uint32_t nb;
        bool have = false, from_new = false;
        if (first) {
          nb = min(n, a.L);
          const uint32_t rank_x = (b.med_in && key_less(b.med_d, a.medoid, b.d, b.id)) ? 1u : 0u;
          if (b.id != kNone && rank_x < nb) { have = true; from_new = true; parent = b.id; }
        } else {
          nb = n ? admit_count(b.below, n, ws, a.L) : 0u;
          if (nb > 0 && (fu == kNone || b.d <= w_dist(s, fu))) { have = true; from_new = true; parent = b.id; }
          else if (fu != kNone) { have = true; parent = s.w[fu].y; }  // (unvisited: no flag in the id word)
        }
        if (!have) {  // nothing unvisited and nothing admitted: the merge would be a no-op — except on a first hop whose
          if (first && nb > 0) {  // sorted list starts with the medoid and offers nothing else: the worklist is [medoid, visited]
            if (lane == 0) s.w[0] = make_uint2(__float_as_uint(b.med_d), a.medoid | kVisitedBit);
            ws = 1;
          }
          break;
        }
        uint32_t scan_from = fu == kNone ? ws : fu;
        if (!from_new) { if (lane == 0) s.w[fu].y = parent | kVisitedBit; scan_from = fu + 1; }
        log_parent(parent);  // thread 0, Inmemory parANN.cu:1399-1418
        const bool capped = iter == a.max_iter - 1;
        if (!capped) my_nb = fetch_adj<PH>(a, parent, pol_stream, lane);  // in flight during the merge
        pf.tick(PT_DECIDE);
        if (nb > 0) {
          ws = merge_worklist(a, s, n, nb, b.below, maxd, ws, first, from_new ? parent : kNone, &pos0);
          scan_from = min(scan_from, pos0);
          pf.count(PT_MERGES);
        }
        __syncwarp();
        pf.tick(PT_MERGE);
        fu = scan_unvisited(s, scan_from, ws);
        pf.tick(PT_UNVIS);
        if (capped) break;
        ++iter;
      }
      write_stats(a, s, q, st);
      if (MODE == kExact) {
        // top-k = head of the worklist (Exact parANN.cu:1273-1276)
        __syncwarp();
        for (uint32_t r = lane; r < a.k; r += 32) {
          a.out_ids[(size_t)q * a.k + r] = r < ws ? (uint64_t)(s.w[r].y & ~kVisitedBit) : 0xFFFFFFFFull;
          a.out_dists[(size_t)q * a.k + r] = r < ws ? w_dist(s, r) : 3.402823466e+38f;
        }
      } else {
        rerank_and_write<T, PH>(a, s, q, ncand);
        pf.tick(PT_RERANK);
      }
    }
#ifdef BANG_PHASE_TIMERS
    pf.stop();
    if (lane == 0 && a.st_phase) for (int i = 0; i < PT_COUNT; ++i) a.st_phase[(size_t)q * PT_COUNT + i] = pf.acc[i];
#endif
    if (lane == 0) {
      if (MODE == kExact && a.dump_ids) {
        a.dump_ids[(size_t)q * a.dump_stride] = a.medoid;
        a.dump_n[q] = min(ncand, a.dump_stride);
      }
      if (a.st_hops) a.st_hops[q] = ncand;
    }
    __syncwarp();
  }
}

// Standalone stage-1 kernel (parity test of populate_pqDist_par): one warp per query, table to global.
template <typename T>
__global__ void __launch_bounds__(32) pq_table_kernel(const SearchArgs a, float* tables) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* q_f = reinterpret_cast<float*>(smem_raw);
  const uint32_t q = blockIdx.x;
  load_query<T>(a, q, q_f);
  __syncwarp();
  build_pq_table(a, q_f, tables + (size_t)q * a.n_chunks * 256);
}

// ------------------------------------------------------------------------------------------------
// host-side launch geometry: how many query warps fit in one CTA / SM
// ------------------------------------------------------------------------------------------------
struct LaunchGeom { int warps_per_cta; int ctas_per_sm; size_t smem; };
// the kernel variant (WPC) that runs `warps` query warps per CTA
inline int wpc_variant(int mode, int warps) { return (mode == kExact || warps <= 16) ? 16 : (warps <= 24 ? 24 : 32); }
template <typename T>
inline LaunchGeom launch_geometry(int mode, uint32_t piv_row, uint32_t n_chunks, uint32_t vec_units, uint32_t L, uint32_t cand_cap,
                                  size_t smem_optin_per_block, size_t smem_per_sm, int max_warps_per_sm, bool piv_global = false,
                                  bool keep_l1 = true) {
  const size_t shared = cta_shared_bytes(mode, piv_row, n_chunks, piv_global), per = warp_private_bytes<T>(mode, piv_row, vec_units, L, cand_cap);
  LaunchGeom g{0, 0, 0};
  if (shared + per > smem_optin_per_block) return g;
  int w = (int)((smem_optin_per_block - shared) / per);
  const int cap = mode == kExact ? 16 : kMaxWarpsPerCta;
  if (w > cap) w = cap;
  if (max_warps_per_sm > 0 && w > max_warps_per_sm) w = max_warps_per_sm;
  // Shared memory and L1 share 256 KB per SM and the split moves in steps (... 164, 196, 228 KB of shared memory, 1 KB per
  // CTA reserved by the system).  The gathers of the traversal are scattered 8..32-byte loads whose lines in flight live
  // in L1: crossing from the 196 KB step (60 KB of L1) to the 228 KB step (28 KB) costs 15 % on the C2 shape at the same
  // number of warps (C2 at 24 warps: 4.65 -> 3.85 ms when the state was cut to fit, profiles/r2_hop_diet.md), so the PQ modes give up a few
  // resident queries to stay below the step.
  if (keep_l1 && mode != kExact) {
    const size_t step = (size_t)196 * 1024 - 1024;
    if (shared + (size_t)w * per > step && shared + 16 * per <= step) w = (int)((step - shared) / per);
  }
  g.warps_per_cta = w;
  g.smem = shared + (size_t)w * per;
  // Exactdistance (no CTA-shared table): several CTAs per SM, bounded by shared memory and the warp budget
  int c = 1;
  if (mode == kExact) {
    c = (int)(smem_per_sm / (g.smem + 1024));
    if (c > 2) c = 2;  // the kernel's launch bound
    if (max_warps_per_sm > 0 && c * w > max_warps_per_sm) c = max_warps_per_sm / w;
    if (c < 1) c = 1;
  }
  g.ctas_per_sm = c;
  return g;
}

}  // namespace bang
