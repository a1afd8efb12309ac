// builder.cu — GPU Vamana index builder (R = 64 RobustPrune) producing the reference's `_disk.bin` graph.
//
// The reference has no builder: it consumes DiskANN's build_disk_index output (README.md:46-58,
// "-R 64 -L 200") converted by bang_preprocess.py.  DiskANN and the datasets are not available offline,
// so indices of the BASELINE.json shapes are built here.  Not on the timed search
// path.  Algorithm: batch-parallel Vamana — for each batch of points: greedy search from the medoid on
// the current graph (this repo's own fused traversal kernel in exact-distance mode, dumping the expanded
// nodes), RobustPrune(alpha) over expanded ∪ current neighbours, then reverse-edge insertion with
// re-prune on overflow.  Batches double in size up to 2 % of N so early points see a dense graph.
// The graph lives in the search kernel's own HBM row layout while it is built.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>

#include "bang_b200.h"
#include "search_kernel.cuh"
#include "search_inst.cuh"

using namespace bang;

namespace {

constexpr int kPruneThreads = 128;
constexpr int kMaxCand = 640;  // expanded-node log (<= 4L+21) + 64 neighbours (+ incoming reverse edges)

__device__ __forceinline__ bool key3_less(float da, uint32_t ia, uint32_t xa, float db, uint32_t ib, uint32_t xb) {
  return da < db || (da == db && (ia < ib || (ia == ib && xa < xb)));
}

// RobustPrune for node p over the candidate ids in c_id[0..n): writes <= 64 neighbours to rows[p].
// Shared arrays are sized kMaxCand.  All threads of the CTA take part.
template <typename T>
__device__ void robust_prune(uint8_t* rows, uint32_t row_stride, uint32_t vec_units, uint32_t p, uint32_t n, float alpha,
                             uint32_t* c_id, float* c_d, uint32_t* s_id, float* s_d, uint8_t* dead, float* pv /*fp32 vec*/,
                             uint32_t* sel, uint32_t* scal) {
  const uint32_t tid = threadIdx.x, t = tid & 7, slot = tid >> 3;
  constexpr int E = Elem<T>::kPerUnit;
  const uint32_t nf = vec_units * E;
  // p's vector -> smem fp32
  {
    const uint8_t* v = rows + (size_t)p * row_stride + kAdjBytes;
    for (uint32_t u = tid; u < vec_units; u += kPruneThreads) {
      float f[E];
      Elem<T>::unpack(*reinterpret_cast<const uint4*>(v + (size_t)u * 16), f);
#pragma unroll
      for (int e = 0; e < E; ++e) pv[u * E + e] = f[e];
    }
  }
  __syncthreads();
  for (uint32_t b0 = 0; b0 < n; b0 += kPruneThreads / 8) {
    const uint32_t i = b0 + slot;
    const uint32_t id = i < n ? c_id[i] : p;
    const float d = l2_row_8lane<T>(rows + (size_t)id * row_stride + kAdjBytes, pv, vec_units, t);
    if (t == 0 && i < n) c_d[i] = d;
  }
  __syncthreads();
  // rank sort by (d, id, index); duplicates of an id end up adjacent
  for (uint32_t i = tid; i < n; i += kPruneThreads) {
    const float d = c_d[i];
    const uint32_t id = c_id[i];
    uint32_t r = 0;
    for (uint32_t j = 0; j < n; ++j) r += key3_less(c_d[j], c_id[j], j, d, id, i) ? 1u : 0u;
    s_d[r] = d;
    s_id[r] = id;
  }
  __syncthreads();
  for (uint32_t i = tid; i < n; i += kPruneThreads)
    dead[i] = (s_id[i] == p || (i > 0 && s_id[i] == s_id[i - 1])) ? 1 : 0;
  if (tid == 0) scal[0] = 0;
  __syncthreads();
  uint32_t cnt = 0;
  for (uint32_t i = 0; i < n && cnt < (uint32_t)kMaxR; ++i) {
    if (dead[i]) continue;  // uniform: dead[] only changes between barriers
    const uint32_t ci = s_id[i];
    if (tid == 0) sel[cnt] = ci;
    ++cnt;
    // selected neighbour's vector -> smem fp32 (reuse c_d region? no: separate buffer after pv)
    float* cv = pv + nf;
    {
      const uint8_t* v = rows + (size_t)ci * row_stride + kAdjBytes;
      for (uint32_t u = tid; u < vec_units; u += kPruneThreads) {
        float f[E];
        Elem<T>::unpack(*reinterpret_cast<const uint4*>(v + (size_t)u * 16), f);
#pragma unroll
        for (int e = 0; e < E; ++e) cv[u * E + e] = f[e];
      }
    }
    __syncthreads();
    for (uint32_t b0 = i + 1; b0 < n; b0 += kPruneThreads / 8) {
      const uint32_t j = b0 + slot;
      const bool live = j < n && !dead[j];
      const uint32_t id = live ? s_id[j] : ci;
      const float dij = l2_row_8lane<T>(rows + (size_t)id * row_stride + kAdjBytes, cv, vec_units, t);
      if (t == 0 && live && alpha * dij <= s_d[j]) dead[j] = 1;
    }
    __syncthreads();
  }
  uint32_t* adj = reinterpret_cast<uint32_t*>(rows + (size_t)p * row_stride);
  __syncthreads();
  for (uint32_t i = tid; i < (uint32_t)kMaxR; i += kPruneThreads) adj[i] = i < cnt ? sel[i] : kNoNbr;
  if (tid == 0) scal[1] = cnt;  // number of neighbours kept
}

struct PruneSmem {
  uint32_t c_id[kMaxCand];
  float c_d[kMaxCand];
  uint32_t s_id[kMaxCand];
  float s_d[kMaxCand];
  uint8_t dead[kMaxCand];
  uint32_t sel[kMaxR];
  uint32_t scal[4];
};

// phase 1: prune each batch point over (expanded nodes of its search) ∪ (its current neighbours)
template <typename T>
__global__ void __launch_bounds__(kPruneThreads) prune_batch_kernel(uint8_t* rows, uint32_t row_stride, uint32_t vec_units,
                                                                    const uint32_t* batch_ids, uint32_t B, const uint32_t* dump_ids,
                                                                    const uint32_t* dump_n, uint32_t dump_stride, float alpha,
                                                                    uint32_t* deg_arr, const uint32_t* ovf, uint32_t* ovf_cnt) {
  extern __shared__ __align__(16) uint8_t raw[];
  PruneSmem* sm = reinterpret_cast<PruneSmem*>(raw);
  float* pv = reinterpret_cast<float*>(raw + align_up(sizeof(PruneSmem), 16));
  const uint32_t b = blockIdx.x;
  if (b >= B) return;
  const uint32_t p = batch_ids[b];
  const uint32_t tid = threadIdx.x;
  const uint32_t extra = min(ovf_cnt[p], 32u);  // pending reverse edges of p (kSlack)
  uint32_t nv = min(dump_n[b], (uint32_t)(kMaxCand - kMaxR - 32));
  for (uint32_t i = tid; i < nv; i += kPruneThreads) sm->c_id[i] = dump_ids[(size_t)b * dump_stride + i];
  const uint32_t* adj = reinterpret_cast<const uint32_t*>(rows + (size_t)p * row_stride);
  uint32_t mine = kNoNbr;
  if (tid < kMaxR) mine = adj[tid];
  const uint32_t deg = __syncthreads_count(tid < kMaxR && mine != kNoNbr);
  if (tid < kMaxR && mine != kNoNbr) sm->c_id[nv + tid] = mine;  // valid entries are the leading ones
  if (tid < extra) sm->c_id[nv + deg + tid] = ovf[(size_t)p * 32 + tid];
  __syncthreads();
  robust_prune<T>(rows, row_stride, vec_units, p, nv + deg + extra, alpha, sm->c_id, sm->c_d, sm->s_id, sm->s_d, sm->dead, pv, sm->sel,
                  sm->scal);
  __syncthreads();
  if (tid == 0) { deg_arr[p] = sm->scal[1]; ovf_cnt[p] = 0; }
}

// ---- reverse edges with slack (DiskANN keeps up to 1.3 R edges before it re-prunes a node) ---------------------
// A node's row holds at most 64 neighbours; reverse edges that do not fit wait in a 32-entry side list and the
// node is re-pruned over (row ∪ side list) only when that list fills up, i.e. once per 32 insertions instead of
// once per insertion.  deg[j] counts the row's valid slots (values > 64 mean "row full").
constexpr int kSlack = 32;

// phase 2a: one thread per new edge p -> j: append p to j's row if it has room, else to j's side list
__global__ void reverse_append_kernel(uint8_t* rows, uint32_t row_stride, const uint32_t* batch_ids, uint32_t B, uint32_t* deg,
                                      uint32_t* ovf, uint32_t* ovf_cnt, uint32_t* prune_list, uint32_t* prune_n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * kMaxR) return;
  const uint32_t p = batch_ids[i / kMaxR];
  const uint32_t j = reinterpret_cast<const uint32_t*>(rows + (size_t)p * row_stride)[i % kMaxR];
  if (j == kNoNbr || j == p) return;
  uint32_t* adj = reinterpret_cast<uint32_t*>(rows + (size_t)j * row_stride);
  const uint4* a4 = reinterpret_cast<const uint4*>(adj);
  bool present = false;
#pragma unroll 4
  for (int k = 0; k < kMaxR / 4; ++k) {
    const uint4 v = a4[k];
    present |= v.x == p || v.y == p || v.z == p || v.w == p;
  }
  if (present) return;
  const uint32_t old = atomicAdd(&deg[j], 1u);
  if (old < (uint32_t)kMaxR) { adj[old] = p; return; }
  const uint32_t slot = atomicAdd(&ovf_cnt[j], 1u);
  if (slot < (uint32_t)kSlack) ovf[(size_t)j * kSlack + slot] = p;
  if (slot == (uint32_t)kSlack - 1) prune_list[atomicAdd(prune_n, 1u)] = j;  // the list just filled up
}

// phase 2b: re-prune the listed nodes over row ∪ side list (persistent grid; `all` = final flush over every node)
template <typename T>
__global__ void __launch_bounds__(kPruneThreads) slack_prune_kernel(uint8_t* rows, uint32_t row_stride, uint32_t vec_units,
                                                                    uint32_t* deg, const uint32_t* ovf, uint32_t* ovf_cnt,
                                                                    const uint32_t* prune_list, const uint32_t* prune_n, uint64_t N,
                                                                    bool all, float alpha) {
  extern __shared__ __align__(16) uint8_t raw[];
  PruneSmem* sm = reinterpret_cast<PruneSmem*>(raw);
  float* pv = reinterpret_cast<float*>(raw + align_up(sizeof(PruneSmem), 16));
  const uint32_t tid = threadIdx.x;
  const uint64_t total = all ? N : (uint64_t)*prune_n;
  for (uint64_t it = blockIdx.x; it < total; it += gridDim.x) {
    const uint32_t j = all ? (uint32_t)it : prune_list[it];
    const uint32_t extra = min(ovf_cnt[j], (uint32_t)kSlack);
    if (extra == 0) continue;  // block-uniform
    const uint32_t* adj = reinterpret_cast<const uint32_t*>(rows + (size_t)j * row_stride);
    uint32_t mine = kNoNbr;
    if (tid < kMaxR) mine = adj[tid];
    const uint32_t d = __syncthreads_count(tid < kMaxR && mine != kNoNbr);
    if (tid < kMaxR && mine != kNoNbr) sm->c_id[tid] = mine;  // valid entries are the leading ones
    if (tid < extra) sm->c_id[d + tid] = ovf[(size_t)j * kSlack + tid];
    __syncthreads();
    robust_prune<T>(rows, row_stride, vec_units, j, d + extra, alpha, sm->c_id, sm->c_d, sm->s_id, sm->s_d, sm->dead, pv, sm->sel,
                    sm->scal);
    __syncthreads();
    if (tid == 0) { deg[j] = sm->scal[1]; ovf_cnt[j] = 0; }
    __syncthreads();
  }
}

// vectors T[N][D] -> HBM rows (adjacency cleared)
__global__ void init_rows_kernel(const uint8_t* vec, uint32_t vec_bytes, uint8_t* rows, uint32_t row_stride, uint64_t N) {
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= N) return;
  uint8_t* d = rows + warp * row_stride;
  for (uint32_t i = lane; i < (uint32_t)kMaxR; i += 32) reinterpret_cast<uint32_t*>(d)[i] = kNoNbr;
  for (uint32_t i = lane; i < row_stride - kAdjBytes; i += 32) d[kAdjBytes + i] = i < vec_bytes ? vec[warp * vec_bytes + i] : (uint8_t)0;
}

__global__ void gather_queries_kernel(const uint8_t* vec, uint32_t vec_bytes, const uint32_t* ids, uint32_t B, uint8_t* out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= (uint64_t)B * vec_bytes) return;
  const uint32_t b = (uint32_t)(i / vec_bytes), o = (uint32_t)(i % vec_bytes);
  out[i] = vec[(size_t)ids[b] * vec_bytes + o];
}

// final: per node, neighbour ids ascending (bang_preprocess.py:102-104), degree, and a fallback edge for isolated nodes
__global__ void finalize_kernel(const uint8_t* rows, uint32_t row_stride, uint64_t N, uint32_t medoid, uint32_t* deg_out,
                                uint32_t* nbr_out, uint32_t pad_value) {
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= N) return;
  const uint32_t* adj = reinterpret_cast<const uint32_t*>(rows + warp * row_stride);
  const uint32_t a = adj[lane], b = adj[lane + 32];  // entries `lane` and `lane + 32`
  // rank of each entry among the 64 by (value, index): padding (0xFFFFFFFF) sorts last, ranks are unique
  uint32_t ra = 0, rb = 0;
  for (int s = 0; s < 32; ++s) {
    const uint32_t oa = __shfl_sync(0xffffffffu, a, s), ob = __shfl_sync(0xffffffffu, b, s);
    ra += (oa < a || (oa == a && s < (int)lane)) ? 1u : 0u;
    ra += (ob < a) ? 1u : 0u;
    rb += (oa <= b) ? 1u : 0u;
    rb += (ob < b || (ob == b && s < (int)lane)) ? 1u : 0u;
  }
  uint32_t* out = nbr_out + warp * kMaxR;
  const uint32_t d = __popc(__ballot_sync(0xffffffffu, a != kNoNbr)) + __popc(__ballot_sync(0xffffffffu, b != kNoNbr));
  out[ra] = a == kNoNbr ? pad_value : a;
  out[rb] = b == kNoNbr ? pad_value : b;
  __syncwarp();
  if (lane == 0) {
    if (d == 0) out[0] = (uint32_t)warp == medoid ? (uint32_t)((medoid + 1) % N) : medoid;
    if (deg_out) deg_out[warp] = d == 0 ? 1u : d;
  }
}

thread_local std::string b_err;
#define B_TRY(expr)                                                                                   \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) {                                                                          \
      b_err = std::string(#expr) + ": " + cudaGetErrorString(_e) + " (builder.cu:" + std::to_string(__LINE__) + ")"; \
      return BANG_E_CUDA;                                                                             \
    }                                                                                                 \
  } while (0)

template <typename T>
int build_impl(const void* d_vectors, uint64_t N, uint32_t D, uint32_t L, float alpha_first, uint64_t n_first, float alpha_rest,
               const uint32_t* d_order, uint64_t n_order, uint32_t medoid, uint32_t max_batch, uint32_t* h_deg, uint32_t* h_nbrs,
               float* stats_out, bool device_out) {
  const uint32_t vec_bytes = D * sizeof(T);
  const uint32_t vec_units = (vec_bytes + 15) / 16;
  const uint32_t row_stride = (uint32_t)align_up(kAdjBytes + (size_t)vec_units * 16, 32);
  const uint32_t max_iter = 4 * L + 20, cand_cap = max_iter + 1;
  if (cand_cap + kMaxR > (uint32_t)kMaxCand) { b_err = "L_build too large for the prune buffers (max 138)"; return BANG_E_ARG; }
  int dev = 0, sms = 0;
  B_TRY(cudaGetDevice(&dev));
  B_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  uint8_t* rows = nullptr;
  B_TRY(cudaMalloc(&rows, (size_t)N * row_stride));
  init_rows_kernel<<<(unsigned)((N * 32 + 255) / 256), 256>>>((const uint8_t*)d_vectors, vec_bytes, rows, row_stride, N);
  B_TRY(cudaGetLastError());

  if (max_batch == 0) max_batch = (uint32_t)std::max<uint64_t>(1024, std::min<uint64_t>(N / 50, 65536));
  const uint32_t MB = max_batch;
  // search scratch
  // the Exactdistance instantiation of the search kernel (search_inst_*.cu)
  search_fn_t kern = sizeof(T) == 4 ? search_kernel_f32(kExact, 0, 16, false) : (std::is_signed<T>::value ? search_kernel_i8(kExact, 0, 16, false) : search_kernel_u8(kExact, 0, 16, false));
  int max_optin = 0, per_sm = 0;
  B_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  B_TRY(cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
  const LaunchGeom geom = launch_geometry<T>(kExact, D, 0, vec_units, L, cand_cap, (size_t)max_optin, (size_t)per_sm, 16);
  if (geom.warps_per_cta < 1) { b_err = "vector too large for the search kernel's shared memory"; return BANG_E_UNSUPPORTED; }
  const size_t smem = geom.smem;
  const int wpc = geom.warps_per_cta;
  B_TRY(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int ctas = 0;
  B_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, (const void*)kern, wpc * 32, smem));
  ctas = std::max(1, std::min(ctas, geom.ctas_per_sm));
  const int grid_max = ctas * sms;
  uint8_t* d_q = nullptr; uint64_t* d_ids = nullptr; float* d_dd = nullptr; uint32_t *d_bloom = nullptr, *d_counter = nullptr;
  uint32_t *d_dump = nullptr, *d_dump_n = nullptr, *d_degc = nullptr, *d_ovf = nullptr, *d_ovfc = nullptr, *d_plist = nullptr,
           *d_pn = nullptr;
  B_TRY(cudaMalloc(&d_q, (size_t)MB * vec_bytes));
  B_TRY(cudaMalloc(&d_ids, (size_t)MB * 8));
  B_TRY(cudaMalloc(&d_dd, (size_t)MB * 4));
  B_TRY(cudaMalloc(&d_bloom, (size_t)grid_max * wpc * kBloomWords * 4));
  B_TRY(cudaMalloc(&d_counter, 4));
  B_TRY(cudaMalloc(&d_dump, (size_t)MB * cand_cap * 4));
  B_TRY(cudaMalloc(&d_dump_n, (size_t)MB * 4));
  B_TRY(cudaMalloc(&d_degc, N * 4));
  B_TRY(cudaMalloc(&d_ovf, N * (size_t)kSlack * 4));
  B_TRY(cudaMalloc(&d_ovfc, N * 4));
  B_TRY(cudaMalloc(&d_plist, N * 4));  // a node enters the list at most once between two of its prunes
  B_TRY(cudaMalloc(&d_pn, 4));
  B_TRY(cudaMemset(d_degc, 0, N * 4));
  B_TRY(cudaMemset(d_ovfc, 0, N * 4));
  const size_t prune_smem = align_up(sizeof(PruneSmem), 16) + (size_t)2 * vec_units * Elem<T>::kPerUnit * 4;
  B_TRY(cudaFuncSetAttribute((const void*)prune_batch_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prune_smem));
  B_TRY(cudaFuncSetAttribute((const void*)slack_prune_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prune_smem));

  SearchArgs a;
  memset(&a, 0, sizeof(a));
  a.rows[0] = rows; a.n_shards = 1; a.row_stride = row_stride; a.D = D; a.vec_units = vec_units; set_medoid(a, medoid, true);
  a.L = L; a.k = 1; a.q_dim = D; a.max_iter = max_iter; a.cand_cap = cand_cap; a.queries = d_q; a.out_ids = d_ids; a.out_dists = d_dd;
  a.bloom = d_bloom; a.counter = d_counter; a.dump_ids = d_dump; a.dump_n = d_dump_n; a.dump_stride = cand_cap;

  uint64_t done = 0;
  uint32_t bs = 1, n_batches = 0;
  cudaEvent_t ev[4];
  for (auto& e : ev) cudaEventCreate(&e);
  float t_search = 0.f, t_prune = 0.f, t_rev = 0.f;
  const bool timed = stats_out != nullptr && getenv("BANG_B200_BUILD_TIMERS") != nullptr;
  while (done < n_order) {
    uint32_t B = (uint32_t)std::min<uint64_t>(bs, n_order - done);
    if (done < n_first) B = (uint32_t)std::min<uint64_t>(B, n_first - done);  // a batch never straddles the two passes
    const float alpha = done < n_first ? alpha_first : alpha_rest;
    const uint32_t* ids = d_order + done;
    gather_queries_kernel<<<(unsigned)(((size_t)B * vec_bytes + 255) / 256), 256>>>((const uint8_t*)d_vectors, vec_bytes, ids, B, d_q);
    B_TRY(cudaMemsetAsync(d_counter, 0, 4));
    a.Q = B;
    if (timed) cudaEventRecord(ev[0]);
    kern<<<std::min<int>((B + wpc - 1) / wpc, grid_max), wpc * 32, smem>>>(a);
    if (timed) cudaEventRecord(ev[1]);
    prune_batch_kernel<T><<<B, kPruneThreads, prune_smem>>>(rows, row_stride, vec_units, ids, B, d_dump, d_dump_n, cand_cap, alpha,
                                                            d_degc, d_ovf, d_ovfc);
    if (timed) cudaEventRecord(ev[2]);
    const uint32_t np = B * kMaxR;
    B_TRY(cudaMemsetAsync(d_pn, 0, 4));
    reverse_append_kernel<<<(np + 255) / 256, 256>>>(rows, row_stride, ids, B, d_degc, d_ovf, d_ovfc, d_plist, d_pn);
    slack_prune_kernel<T><<<sms * 8, kPruneThreads, prune_smem>>>(rows, row_stride, vec_units, d_degc, d_ovf, d_ovfc, d_plist, d_pn,
                                                                  N, false, alpha);
    B_TRY(cudaGetLastError());
    if (timed) {
      cudaEventRecord(ev[3]);
      cudaEventSynchronize(ev[3]);
      float ms;
      cudaEventElapsedTime(&ms, ev[0], ev[1]); t_search += ms;
      cudaEventElapsedTime(&ms, ev[1], ev[2]); t_prune += ms;
      cudaEventElapsedTime(&ms, ev[2], ev[3]); t_rev += ms;
    }
    done += B;
    ++n_batches;
    if (bs < MB) bs = std::min<uint32_t>(MB, bs * 2);
  }
  // final flush: every node with pending reverse edges is pruned over row ∪ side list
  slack_prune_kernel<T><<<sms * 8, kPruneThreads, prune_smem>>>(rows, row_stride, vec_units, d_degc, d_ovf, d_ovfc, d_plist, d_pn, N,
                                                                true, alpha_rest);
  B_TRY(cudaGetLastError());
  B_TRY(cudaDeviceSynchronize());
  // the per-batch scratch is no longer needed: release it before the outputs are materialised
  cudaFree(d_q); cudaFree(d_ids); cudaFree(d_dd); cudaFree(d_bloom); cudaFree(d_counter);
  cudaFree(d_dump); cudaFree(d_dump_n); cudaFree(d_degc); cudaFree(d_ovf); cudaFree(d_ovfc); cudaFree(d_plist); cudaFree(d_pn);
  if (device_out) {
    // outputs are DEVICE arrays: neighbours ascending, unused slots 0xFFFFFFFF (the search kernel's padding)
    finalize_kernel<<<(unsigned)((N * 32 + 255) / 256), 256>>>(rows, row_stride, N, medoid, h_deg, h_nbrs, kNoNbr);
    B_TRY(cudaGetLastError());
    B_TRY(cudaDeviceSynchronize());
  } else {
    uint32_t *d_deg = nullptr, *d_nbr = nullptr;
    B_TRY(cudaMalloc(&d_deg, N * 4));
    B_TRY(cudaMalloc(&d_nbr, N * kMaxR * 4));
    finalize_kernel<<<(unsigned)((N * 32 + 255) / 256), 256>>>(rows, row_stride, N, medoid, d_deg, d_nbr, 0u);
    B_TRY(cudaGetLastError());
    B_TRY(cudaMemcpy(h_deg, d_deg, N * 4, cudaMemcpyDeviceToHost));
    B_TRY(cudaMemcpy(h_nbrs, d_nbr, N * kMaxR * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_deg); cudaFree(d_nbr);
  }
  if (stats_out) { stats_out[0] = (float)n_batches; stats_out[1] = t_search; stats_out[2] = t_prune; stats_out[3] = t_rev; }
  for (auto& e : ev) cudaEventDestroy(e);
  cudaFree(rows);
  return BANG_OK;
}

}  // namespace

extern "C" const char* bang_b200_builder_last_error(void) { return b_err.c_str(); }

// d_vectors: device T[N][D]; d_order: device u32[n_order] insertion order (may repeat ids for a second pass);
// medoid: entry point; outputs are host arrays: deg u32[N], nbrs u32[N][64] ascending.
// The first n_first insertions use alpha_first (DiskANN's first pass runs with alpha = 1), the rest alpha_rest.
extern "C" int bang_b200_build_vamana(int dtype, const void* d_vectors, uint64_t N, uint32_t D, uint32_t L_build, float alpha_first,
                                      uint64_t n_first, float alpha_rest, const uint32_t* d_order, uint64_t n_order, uint64_t medoid,
                                      uint32_t max_batch, uint32_t* h_deg, uint32_t* h_nbrs, float* stats_out) {
  // max_batch bit 31: the outputs are DEVICE arrays (h_deg may then be null; unused neighbour slots = 0xFFFFFFFF)
  const bool device_out = (max_batch & 0x80000000u) != 0;
  max_batch &= 0x7FFFFFFFu;
  if (!d_vectors || !d_order || (!h_deg && !device_out) || !h_nbrs || N < 2 || medoid >= N) { b_err = "bad argument"; return BANG_E_ARG; }
  switch (dtype) {
    case BANG_DT_FLOAT: return build_impl<float>(d_vectors, N, D, L_build, alpha_first, n_first, alpha_rest, d_order, n_order, (uint32_t)medoid, max_batch, h_deg, h_nbrs, stats_out, device_out);
    case BANG_DT_INT8: return build_impl<int8_t>(d_vectors, N, D, L_build, alpha_first, n_first, alpha_rest, d_order, n_order, (uint32_t)medoid, max_batch, h_deg, h_nbrs, stats_out, device_out);
    case BANG_DT_UINT8: return build_impl<uint8_t>(d_vectors, N, D, L_build, alpha_first, n_first, alpha_rest, d_order, n_order, (uint32_t)medoid, max_batch, h_deg, h_nbrs, stats_out, device_out);
  }
  b_err = "bad dtype";
  return BANG_E_ARG;
}
