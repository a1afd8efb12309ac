// prep_kernels.cu — the two data-preparation steps the reference leaves to DiskANN's tools (README.md:46-58), as
// hand-written sm_100a kernels behind the C ABI (include/bang_b200.h):
//
//   * exact k-nearest-neighbour ground truth (`compute_groundtruth`): brute force over all base points, results
//     ordered by (distance, id) — the truthset the recall report reads (test_driver.cpp:238-272);
//   * PQ training and encoding (`build_disk_index`'s PQ stage): Lloyd k-means with 256 centres per chunk on a
//     training sample, then one code byte per (point, chunk) = the closest centre.  The outputs are the arrays of
//     `_pq_pivots.bin` / `_pq_compressed.bin`.
//
// Ground truth.  A query tile x base tile kernel computes all pair distances with 4x4 register blocking (u8/i8:
// packed dp4a dot products, distance = |q|^2 + |b|^2 - 2 q.b in exact integers; float: fmaf of differences) and
// appends only the pairs that come before the query's current k-th (distance, id) to a per-query candidate buffer; a
// per-query select kernel (bitonic sort by (distance, id)) folds the buffer into the running top-k and tightens
// the bound.  Base ranges grow geometrically, so a query appends O(k) candidates per pass.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "bang_b200.h"

namespace {

thread_local std::string p_err;
#define P_TRY(expr)                                                                                               \
  do {                                                                                                            \
    cudaError_t _e = (expr);                                                                                      \
    if (_e != cudaSuccess) {                                                                                      \
      p_err = std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"; \
      return BANG_E_CUDA;                                                                                         \
    }                                                                                                             \
  } while (0)

constexpr int kTile = 64;        // queries per CTA tile = base points per CTA tile
constexpr int kStepWords = 32;   // 32-bit words of one row consumed per shared-memory stage (u8/i8: 128 dims, float: 32 dims)
constexpr int kPad = 4;          // row padding of the transposed stage (keeps the 16-byte reads aligned)
constexpr uint32_t kMaxK = 128;
constexpr uint32_t kSlots = 4096;                 // entries the select kernel sorts: candidates + running top-k
constexpr uint32_t kCandCap = kSlots - 2 * kMaxK;  // 3840 candidates per query and pass

template <typename T> struct GT;
template <> struct GT<float> {
  typedef float acc_t;
  __device__ static __forceinline__ void mac(float& acc, uint32_t q, uint32_t b) {
    const float d = __fsub_rn(__uint_as_float(q), __uint_as_float(b));
    acc = __fmaf_rn(d, d, acc);
  }
  // sortable 32-bit key of a distance (>= 0) and the distance as float
  __device__ static __forceinline__ uint32_t key(float acc, int, int) { return __float_as_uint(acc); }
};
template <> struct GT<uint8_t> {
  typedef int acc_t;
  __device__ static __forceinline__ void mac(int& acc, uint32_t q, uint32_t b) { acc = (int)__dp4a(q, b, (uint32_t)acc); }
  __device__ static __forceinline__ uint32_t key(int dot, int qn, int bn) { return (uint32_t)(qn + bn - 2 * dot); }
};
template <> struct GT<int8_t> {
  typedef int acc_t;
  __device__ static __forceinline__ void mac(int& acc, uint32_t q, uint32_t b) { acc = __dp4a((int)q, (int)b, acc); }
  __device__ static __forceinline__ uint32_t key(int dot, int qn, int bn) { return (uint32_t)(qn + bn - 2 * dot); }
};

// squared norms of the rows (integer types only)
template <typename T>
__global__ void norms_kernel(const T* __restrict__ x, uint64_t n, uint32_t D, int* __restrict__ out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = 0;
  for (uint32_t j = 0; j < D; ++j) { const int v = (int)x[i * D + j]; s += v * v; }
  out[i] = s;
}

// one 32-bit word of row `r` (zero beyond the row's D elements / beyond n rows)
template <typename T>
__device__ __forceinline__ uint32_t row_word(const T* __restrict__ x, uint64_t n, uint32_t D, uint64_t r, uint32_t w) {
  if (r >= n) return 0u;
  if (sizeof(T) == 4) return w < D ? reinterpret_cast<const uint32_t*>(x)[r * D + w] : 0u;
  const uint8_t* p = reinterpret_cast<const uint8_t*>(x) + r * D;
  if ((D & 3u) == 0) return 4 * w < D ? reinterpret_cast<const uint32_t*>(p)[w] : 0u;  // rows are 4-byte aligned
  uint32_t v = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) { const uint32_t j = 4 * w + b; if (j < D) v |= (uint32_t)p[j] << (8 * b); }
  return v;
}

// grid = (base tiles of the pass, query tiles); 256 threads; thread (ty, tx) owns queries ty*4.. and base points tx*4..
template <typename T>
__global__ void __launch_bounds__(256) gt_pairs_kernel(const T* __restrict__ base, uint64_t n_base, uint64_t first, uint64_t count,
                                                       uint64_t id_offset, const T* __restrict__ queries, uint32_t nq, uint32_t D,
                                                       const int* __restrict__ base_norm, const int* __restrict__ query_norm,
                                                       const unsigned long long* __restrict__ bound /*[nq] key<<32 | id of the k-th best*/,
                                                       unsigned long long* __restrict__ cand /*[nq][kCandCap] key<<32 | id*/,
                                                       uint32_t* __restrict__ cand_n, uint64_t id_base_for_key) {
  __shared__ __align__(16) uint32_t qs[kStepWords][kTile + kPad];
  __shared__ __align__(16) uint32_t bs[kStepWords][kTile + kPad];
  typedef typename GT<T>::acc_t acc_t;
  const uint32_t tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const uint64_t b0 = first + (uint64_t)blockIdx.x * kTile;
  const uint32_t q0 = blockIdx.y * kTile;
  const uint32_t words = sizeof(T) == 4 ? D : (D + 3) / 4;
  acc_t acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;
  const uint64_t b_end = first + count;
  for (uint32_t w0 = 0; w0 < words; w0 += kStepWords) {
    // stage: 64 rows x 32 words of the queries and of the base tile, transposed to [word][row]
    for (uint32_t i = threadIdx.x; i < kTile * kStepWords; i += 256) {
      const uint32_t r = i / kStepWords, w = i % kStepWords;
      qs[w][r] = row_word<T>(queries, nq, D, (uint64_t)q0 + r, w0 + w);
      const uint64_t br = b0 + r;
      bs[w][r] = br < b_end ? row_word<T>(base, n_base, D, br, w0 + w) : 0u;
    }
    __syncthreads();
    const uint32_t wn = min((uint32_t)kStepWords, words - w0);
    for (uint32_t w = 0; w < wn; ++w) {
      const uint4 qv = *reinterpret_cast<const uint4*>(&qs[w][ty * 4]);
      const uint4 bv = *reinterpret_cast<const uint4*>(&bs[w][tx * 4]);
      const uint32_t qa[4] = {qv.x, qv.y, qv.z, qv.w}, ba[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) GT<T>::mac(acc[i][j], qa[i], ba[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t q = q0 + ty * 4 + i;
    if (q >= nq) continue;
    const unsigned long long bnd = bound[q];
    const int qn = sizeof(T) == 4 ? 0 : query_norm[q];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t b = b0 + tx * 4 + j;
      if (b >= b_end) continue;
      const uint32_t key = GT<T>::key(acc[i][j], qn, sizeof(T) == 4 ? 0 : base_norm[b]);
      const unsigned long long ent = ((unsigned long long)key << 32) | (uint32_t)(b + id_offset - id_base_for_key);
      if (ent < bnd) {   // strictly before the current k-th (distance, id): ties with larger ids never enter
        const uint32_t pos = atomicAdd(cand_n + q, 1u);
        if (pos < kCandCap) cand[(size_t)q * kCandCap + pos] = ent;
      }
    }
  }
}

// one CTA per query: fold the candidate buffer into the running top-k (both as key<<32 | id, ascending), drop repeated
// ids, tighten the bound.  overflow[0] is set when a buffer had more than kCandCap entries (the pass is then repeated
// with the tightened bounds; the entries that did fit are all genuine, repeated ids are dropped here).
__global__ void __launch_bounds__(1024) gt_select_kernel(unsigned long long* __restrict__ cand, uint32_t* __restrict__ cand_n,
                                                         unsigned long long* __restrict__ topk /*[nq][kMaxK]*/, uint32_t* __restrict__ topk_n,
                                                         uint32_t k, unsigned long long* __restrict__ bound, uint32_t* __restrict__ overflow) {
  __shared__ unsigned long long s[kSlots];  // 32 KB
  const uint32_t q = blockIdx.x;
  uint32_t n_c = cand_n[q];
  if (n_c == 0) return;
  if (n_c > kCandCap) { if (threadIdx.x == 0) atomicExch(overflow, 1u); n_c = kCandCap; }
  const uint32_t n_t = topk_n[q];
  const uint32_t n = n_c + n_t;   // <= kCandCap + kMaxK < kSlots
  uint32_t len = 1;
  while (len < n) len <<= 1;
  for (uint32_t i = threadIdx.x; i < len; i += blockDim.x)
    s[i] = i < n_c ? cand[(size_t)q * kCandCap + i] : (i < n ? topk[(size_t)q * kMaxK + (i - n_c)] : ~0ull);
  __syncthreads();
  for (uint32_t size = 2; size <= len; size <<= 1) {       // bitonic sort, ascending
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = threadIdx.x; i < len / 2; i += blockDim.x) {
        const uint32_t lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = s[lo], b = s[hi];
        if ((a > b) == up) { s[lo] = b; s[hi] = a; }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    uint32_t u = 0;
    unsigned long long prev = ~0ull;
    for (uint32_t i = 0; i < n && u < k; ++i) {
      const unsigned long long v = s[i];
      if (v == prev) continue;  // the same (distance, id) again: a repeated pass re-appends what is already in the top-k
      topk[(size_t)q * kMaxK + u++] = v;
      prev = v;
    }
    topk_n[q] = u;
    cand_n[q] = 0;
    if (u == k) bound[q] = topk[(size_t)q * kMaxK + k - 1];
  }
}

template <typename T>
__global__ void gt_write_kernel(const unsigned long long* __restrict__ topk, const uint32_t* __restrict__ topk_n, uint32_t nq, uint32_t k,
                                uint64_t id_base_for_key, uint32_t* __restrict__ ids, float* __restrict__ dists) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq * k) return;
  const uint32_t q = i / k, r = i % k;
  if (r < topk_n[q]) {
    const unsigned long long v = topk[(size_t)q * kMaxK + r];
    ids[i] = (uint32_t)((uint32_t)v + id_base_for_key);
    dists[i] = sizeof(T) == 4 ? __uint_as_float((uint32_t)(v >> 32)) : (float)(uint32_t)(v >> 32);
  } else {
    ids[i] = 0xFFFFFFFFu;
    dists[i] = 3.402823466e+38f;
  }
}

template <typename T>
int gt_run(const T* d_base, uint64_t n, uint32_t D, const T* d_queries, uint32_t nq, uint32_t k, uint64_t id_offset, uint32_t* d_ids,
           float* d_dists, cudaStream_t st) {
  if (k == 0 || k > kMaxK) { p_err = "k must be in 1..128"; return BANG_E_ARG; }
  if (id_offset + n > 0xFFFFFFFFull) { p_err = "ids exceed 32 bits"; return BANG_E_ARG; }
  int *bn = nullptr, *qn = nullptr;
  uint32_t *cand_n = nullptr, *topk_n = nullptr, *overflow = nullptr;
  unsigned long long *cand = nullptr, *topk = nullptr, *bound = nullptr;
  int rc = BANG_OK;
  auto body = [&]() -> int {
    if (sizeof(T) == 1) {
      P_TRY(cudaMalloc(&bn, n * sizeof(int)));
      P_TRY(cudaMalloc(&qn, (size_t)nq * sizeof(int)));
      norms_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_base, n, D, bn);
      norms_kernel<T><<<(nq + 255) / 256, 256, 0, st>>>(d_queries, nq, D, qn);
    }
    P_TRY(cudaMalloc(&bound, (size_t)nq * 8));
    P_TRY(cudaMalloc(&cand_n, (size_t)nq * 4));
    P_TRY(cudaMalloc(&topk_n, (size_t)nq * 4));
    P_TRY(cudaMalloc(&overflow, 4));
    P_TRY(cudaMalloc(&cand, (size_t)nq * kCandCap * 8));
    P_TRY(cudaMalloc(&topk, (size_t)nq * kMaxK * 8));
    P_TRY(cudaMemsetAsync(bound, 0xFF, (size_t)nq * 8, st));   // no bound yet: everything is a candidate
    P_TRY(cudaMemsetAsync(cand_n, 0, (size_t)nq * 4, st));
    P_TRY(cudaMemsetAsync(topk_n, 0, (size_t)nq * 4, st));
    // the first pass holds at most kCandCap points (every one of them is a candidate), later passes grow 4x
    uint64_t done = 0, pass = kCandCap / kTile * kTile;
    const dim3 qt((nq + kTile - 1) / kTile);
    while (done < n) {
      const uint64_t cnt = std::min<uint64_t>(pass, n - done);
      for (int attempt = 0;; ++attempt) {
        P_TRY(cudaMemsetAsync(overflow, 0, 4, st));
        const dim3 grid((unsigned)((cnt + kTile - 1) / kTile), qt.x);
        gt_pairs_kernel<T><<<grid, 256, 0, st>>>(d_base, n, done, cnt, id_offset, d_queries, nq, D, bn, qn, bound, cand, cand_n, 0);
        gt_select_kernel<<<nq, 1024, 0, st>>>(cand, cand_n, topk, topk_n, k, bound, overflow);
        P_TRY(cudaGetLastError());
        uint32_t ovf = 0;
        P_TRY(cudaMemcpyAsync(&ovf, overflow, 4, cudaMemcpyDeviceToHost, st));
        P_TRY(cudaStreamSynchronize(st));
        if (!ovf) break;
        if (attempt == 16) { p_err = "ground truth: candidate buffers keep overflowing (base rows not visited in ascending id order?)"; return BANG_E_UNSUPPORTED; }
      }
      done += cnt;
      pass = std::min<uint64_t>(pass * 4, 1ull << 26);
    }
    gt_write_kernel<T><<<(nq * k + 255) / 256, 256, 0, st>>>(topk, topk_n, nq, k, 0, d_ids, d_dists);
    P_TRY(cudaGetLastError());
    P_TRY(cudaStreamSynchronize(st));
    return BANG_OK;
  };
  rc = body();
  cudaFree(bn); cudaFree(qn); cudaFree(bound); cudaFree(cand_n); cudaFree(topk_n); cudaFree(overflow); cudaFree(cand); cudaFree(topk);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// PQ: k-means (256 centres per chunk) and encoding
// ------------------------------------------------------------------------------------------------
constexpr int kMaxChunkDims = 32;
constexpr double kFix = 1048576.0;  // fixed-point scale of the centre sums: integer atomics make the update order-independent

template <typename T>
__device__ __forceinline__ float elem_f(const T* x, uint64_t i) { return (float)x[i]; }

// closest of the 256 centres of a chunk for the residual r[j] = x[j] - centroid[j]; first minimum wins.
// CS > 0: chunk size known at compile time (residual in registers); CS = 0: any size up to kMaxChunkDims.
template <int CS>
__device__ __forceinline__ uint32_t closest_centre(const float* __restrict__ piv_chunk /*smem [256][cs]*/, uint32_t cs, const float* r) {
  float best = 3.402823466e+38f;
  uint32_t arg = 0;
  for (uint32_t c = 0; c < 256; ++c) {
    float acc = 0.0f;
    if (CS > 0) {
#pragma unroll
      for (int j = 0; j < CS; ++j) { const float d = __fsub_rn(r[j], piv_chunk[c * CS + j]); acc = __fmaf_rn(d, d, acc); }
    } else {
      for (uint32_t j = 0; j < cs; ++j) { const float d = __fsub_rn(r[j], piv_chunk[c * cs + j]); acc = __fmaf_rn(d, d, acc); }
    }
    if (acc < best) { best = acc; arg = c; }
  }
  return arg;
}
__device__ __forceinline__ uint32_t closest_centre_any(const float* piv_chunk, uint32_t cs, const float* r) {
  switch (cs) {   // warp-uniform: one chunk per CTA
    case 1: return closest_centre<1>(piv_chunk, cs, r);
    case 2: return closest_centre<2>(piv_chunk, cs, r);
    case 3: return closest_centre<3>(piv_chunk, cs, r);
    case 4: return closest_centre<4>(piv_chunk, cs, r);
    case 8: return closest_centre<8>(piv_chunk, cs, r);
    default: return closest_centre<0>(piv_chunk, cs, r);
  }
}

// grid = (point blocks, chunks): codes[i][c] for 256 points of one chunk per CTA
template <typename T>
__global__ void __launch_bounds__(256) pq_encode_kernel(const T* __restrict__ x, uint64_t n, uint32_t D, const float* __restrict__ pivots,
                                                        const float* __restrict__ centroid, const uint32_t* __restrict__ chunk_off, uint32_t m,
                                                        uint8_t* __restrict__ codes) {
  __shared__ float piv_s[256 * kMaxChunkDims];
  const uint32_t c = blockIdx.y, j0 = chunk_off[c], cs = chunk_off[c + 1] - j0;
  for (uint32_t i = threadIdx.x; i < 256 * cs; i += 256) piv_s[i] = pivots[(size_t)(i / cs) * D + j0 + i % cs];
  __syncthreads();
  const uint64_t p = blockIdx.x * 256ull + threadIdx.x;
  if (p >= n) return;
  float r[kMaxChunkDims];
  for (uint32_t j = 0; j < cs; ++j) r[j] = __fsub_rn(elem_f(x, p * D + j0 + j), centroid[j0 + j]);
  codes[p * m + c] = (uint8_t)closest_centre_any(piv_s, cs, r);
}

// one Lloyd assignment pass over the (centred, float) training rows of one chunk: fixed-point sums and counts per centre
__global__ void __launch_bounds__(256) kmeans_assign_kernel(const float* __restrict__ train, uint64_t nt, uint32_t D, const float* __restrict__ pivots,
                                                            const uint32_t* __restrict__ chunk_off, long long* __restrict__ sums /*[m][256][kMaxChunkDims]*/,
                                                            unsigned long long* __restrict__ counts /*[m][256]*/) {
  __shared__ float piv_s[256 * kMaxChunkDims];
  const uint32_t c = blockIdx.y, j0 = chunk_off[c], cs = chunk_off[c + 1] - j0;
  for (uint32_t i = threadIdx.x; i < 256 * cs; i += 256) piv_s[i] = pivots[(size_t)(i / cs) * D + j0 + i % cs];
  __syncthreads();
  const uint64_t p = blockIdx.x * 256ull + threadIdx.x;
  if (p >= nt) return;
  float r[kMaxChunkDims];
  for (uint32_t j = 0; j < cs; ++j) r[j] = train[p * D + j0 + j];
  const uint32_t k = closest_centre_any(piv_s, cs, r);
  for (uint32_t j = 0; j < cs; ++j)
    atomicAdd(reinterpret_cast<unsigned long long*>(sums) + ((size_t)c * 256 + k) * kMaxChunkDims + j, (unsigned long long)(long long)llrint((double)r[j] * kFix));
  atomicAdd(counts + (size_t)c * 256 + k, 1ull);
}

__global__ void kmeans_update_kernel(float* __restrict__ pivots, uint32_t D, const uint32_t* __restrict__ chunk_off, uint32_t m,
                                     const long long* __restrict__ sums, const unsigned long long* __restrict__ counts) {
  const uint32_t c = blockIdx.x, k = threadIdx.x;
  const uint32_t j0 = chunk_off[c], cs = chunk_off[c + 1] - j0;
  const unsigned long long cnt = counts[(size_t)c * 256 + k];
  if (cnt == 0) return;  // an empty centre keeps its position
  for (uint32_t j = 0; j < cs; ++j)
    pivots[(size_t)k * D + j0 + j] = (float)((double)sums[((size_t)c * 256 + k) * kMaxChunkDims + j] / kFix / (double)cnt);
}

// initial centres: 256 training rows picked by a seeded hash (distinct rows when nt >= 256)
__global__ void kmeans_init_kernel(const float* __restrict__ train, uint64_t nt, uint32_t D, const uint32_t* __restrict__ chunk_off, uint32_t m,
                                   uint64_t seed, float* __restrict__ pivots) {
  const uint32_t c = blockIdx.x, k = threadIdx.x;
  const uint32_t j0 = chunk_off[c], cs = chunk_off[c + 1] - j0;
  unsigned long long h = seed ^ (0x9E3779B97F4A7C15ull * (c + 1));
  h ^= h >> 31; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 29;
  // a stride walk that visits distinct rows: start + k * step (mod nt), step odd and coprime enough for synthetic data
  const uint64_t step = nt >= 256 ? (nt / 256) | 1ull : 1ull;
  const uint64_t row = (h % nt + (uint64_t)k * step) % nt;
  for (uint32_t j = 0; j < cs; ++j) pivots[(size_t)k * D + j0 + j] = train[row * D + j0 + j];
}

template <typename T>
__global__ void centre_rows_kernel(const T* __restrict__ x, const uint64_t* __restrict__ rows, uint64_t nt, uint32_t D, const float* __restrict__ centroid,
                                   float* __restrict__ out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= nt * D) return;
  const uint64_t r = rows ? rows[i / D] : i / D;
  out[i] = __fsub_rn(elem_f(x, r * D + i % D), centroid[i % D]);
}

// column means in double (the PQ centroid), grid-stride over rows, one CTA per 32 columns
template <typename T>
__global__ void column_sum_kernel(const T* __restrict__ x, uint64_t n, uint32_t D, double* __restrict__ sums) {
  const uint32_t j = blockIdx.x * 32 + (threadIdx.x & 31);
  const uint32_t lanes_y = blockDim.x >> 5, y = threadIdx.x >> 5;
  double s = 0.0;
  if (j < D)
    for (uint64_t i = (uint64_t)blockIdx.y * lanes_y + y; i < n; i += (uint64_t)gridDim.y * lanes_y) s += (double)elem_f(x, i * D + j);
  if (j < D) atomicAdd(sums + j, s);
}

template <typename T>
int pq_train_run(const T* d_base, uint64_t n, uint32_t D, const uint32_t* h_chunk_off, uint32_t m, uint32_t iters, uint64_t max_train,
                 uint64_t seed, float* h_pivots, float* h_centroid, cudaStream_t st) {
  for (uint32_t c = 0; c < m; ++c)
    if (h_chunk_off[c + 1] < h_chunk_off[c] || h_chunk_off[c + 1] - h_chunk_off[c] > (uint32_t)kMaxChunkDims || h_chunk_off[c + 1] - h_chunk_off[c] == 0) {
      p_err = "PQ chunks must span 1..32 dimensions"; return BANG_E_UNSUPPORTED;
    }
  if (h_chunk_off[0] != 0 || h_chunk_off[m] != D) { p_err = "chunk offsets must cover [0, D)"; return BANG_E_ARG; }
  double* d_sum = nullptr; float *d_cen = nullptr, *d_train = nullptr, *d_piv = nullptr; uint32_t* d_off = nullptr; uint64_t* d_rows = nullptr;
  long long* d_sums = nullptr; unsigned long long* d_cnt = nullptr;
  const uint64_t nt = std::min<uint64_t>(n, max_train ? max_train : n);
  auto body = [&]() -> int {
    P_TRY(cudaMalloc(&d_sum, (size_t)D * 8));
    P_TRY(cudaMalloc(&d_cen, (size_t)D * 4));
    P_TRY(cudaMalloc(&d_off, (size_t)(m + 1) * 4));
    P_TRY(cudaMalloc(&d_piv, (size_t)256 * D * 4));
    P_TRY(cudaMalloc(&d_train, nt * (size_t)D * 4));
    P_TRY(cudaMalloc(&d_sums, (size_t)m * 256 * kMaxChunkDims * 8));
    P_TRY(cudaMalloc(&d_cnt, (size_t)m * 256 * 8));
    P_TRY(cudaMemsetAsync(d_sum, 0, (size_t)D * 8, st));
    P_TRY(cudaMemcpyAsync(d_off, h_chunk_off, (size_t)(m + 1) * 4, cudaMemcpyHostToDevice, st));
    column_sum_kernel<T><<<dim3((D + 31) / 32, 592), 256, 0, st>>>(d_base, n, D, d_sum);
    std::vector<double> hs(D);
    P_TRY(cudaMemcpyAsync(hs.data(), d_sum, (size_t)D * 8, cudaMemcpyDeviceToHost, st));
    P_TRY(cudaStreamSynchronize(st));
    for (uint32_t j = 0; j < D; ++j) h_centroid[j] = (float)(hs[j] / (double)n);
    P_TRY(cudaMemcpyAsync(d_cen, h_centroid, (size_t)D * 4, cudaMemcpyHostToDevice, st));
    // training sample: nt rows spread evenly over the base (row i * n / nt), centred
    std::vector<uint64_t> rows(nt);
    for (uint64_t i = 0; i < nt; ++i) rows[i] = (uint64_t)((unsigned __int128)i * n / nt);
    P_TRY(cudaMalloc(&d_rows, nt * 8));
    P_TRY(cudaMemcpyAsync(d_rows, rows.data(), nt * 8, cudaMemcpyHostToDevice, st));
    centre_rows_kernel<T><<<(unsigned)((nt * D + 255) / 256), 256, 0, st>>>(d_base, d_rows, nt, D, d_cen, d_train);
    kmeans_init_kernel<<<m, 256, 0, st>>>(d_train, nt, D, d_off, m, seed, d_piv);
    for (uint32_t it = 0; it < iters; ++it) {
      P_TRY(cudaMemsetAsync(d_sums, 0, (size_t)m * 256 * kMaxChunkDims * 8, st));
      P_TRY(cudaMemsetAsync(d_cnt, 0, (size_t)m * 256 * 8, st));
      kmeans_assign_kernel<<<dim3((unsigned)((nt + 255) / 256), m), 256, 0, st>>>(d_train, nt, D, d_piv, d_off, d_sums, d_cnt);
      kmeans_update_kernel<<<m, 256, 0, st>>>(d_piv, D, d_off, m, d_sums, d_cnt);
    }
    P_TRY(cudaGetLastError());
    P_TRY(cudaMemcpyAsync(h_pivots, d_piv, (size_t)256 * D * 4, cudaMemcpyDeviceToHost, st));
    P_TRY(cudaStreamSynchronize(st));
    return BANG_OK;
  };
  const int rc = body();
  cudaFree(d_sum); cudaFree(d_cen); cudaFree(d_off); cudaFree(d_piv); cudaFree(d_train); cudaFree(d_sums); cudaFree(d_cnt); cudaFree(d_rows);
  return rc;
}

template <typename T>
int pq_encode_run(const T* d_base, uint64_t n, uint32_t D, const float* h_pivots, const float* h_centroid, const uint32_t* h_chunk_off, uint32_t m,
                  uint8_t* d_codes, cudaStream_t st) {
  for (uint32_t c = 0; c < m; ++c)
    if (h_chunk_off[c + 1] <= h_chunk_off[c] || h_chunk_off[c + 1] - h_chunk_off[c] > (uint32_t)kMaxChunkDims) { p_err = "PQ chunks must span 1..32 dimensions"; return BANG_E_UNSUPPORTED; }
  float *d_piv = nullptr, *d_cen = nullptr; uint32_t* d_off = nullptr;
  auto body = [&]() -> int {
    P_TRY(cudaMalloc(&d_piv, (size_t)256 * D * 4));
    P_TRY(cudaMalloc(&d_cen, (size_t)D * 4));
    P_TRY(cudaMalloc(&d_off, (size_t)(m + 1) * 4));
    P_TRY(cudaMemcpyAsync(d_piv, h_pivots, (size_t)256 * D * 4, cudaMemcpyHostToDevice, st));
    P_TRY(cudaMemcpyAsync(d_cen, h_centroid, (size_t)D * 4, cudaMemcpyHostToDevice, st));
    P_TRY(cudaMemcpyAsync(d_off, h_chunk_off, (size_t)(m + 1) * 4, cudaMemcpyHostToDevice, st));
    pq_encode_kernel<T><<<dim3((unsigned)((n + 255) / 256), m), 256, 0, st>>>(d_base, n, D, d_piv, d_cen, d_off, m, d_codes);
    P_TRY(cudaGetLastError());
    P_TRY(cudaStreamSynchronize(st));
    return BANG_OK;
  };
  const int rc = body();
  cudaFree(d_piv); cudaFree(d_cen); cudaFree(d_off);
  return rc;
}

}  // namespace

extern "C" const char* bang_b200_prep_last_error(void) { return p_err.c_str(); }

extern "C" int bang_b200_bruteforce_gt(bang_dtype_t dtype, const void* d_base, uint64_t n, uint32_t D, const void* d_queries, uint32_t nq,
                                       uint32_t k, uint64_t id_offset, uint32_t* d_ids, float* d_dists, void* cuda_stream) {
  if (!d_base || !d_queries || !d_ids || !d_dists || n == 0 || nq == 0 || D == 0) { p_err = "bad argument"; return BANG_E_ARG; }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  switch (dtype) {
    case BANG_DT_FLOAT: return gt_run<float>((const float*)d_base, n, D, (const float*)d_queries, nq, k, id_offset, d_ids, d_dists, st);
    case BANG_DT_INT8: return gt_run<int8_t>((const int8_t*)d_base, n, D, (const int8_t*)d_queries, nq, k, id_offset, d_ids, d_dists, st);
    case BANG_DT_UINT8: return gt_run<uint8_t>((const uint8_t*)d_base, n, D, (const uint8_t*)d_queries, nq, k, id_offset, d_ids, d_dists, st);
    default: p_err = "bad dtype"; return BANG_E_ARG;
  }
}

extern "C" int bang_b200_pq_train(bang_dtype_t dtype, const void* d_base, uint64_t n, uint32_t D, const uint32_t* chunk_offsets, uint32_t n_chunks,
                                  uint32_t iters, uint64_t max_train, uint64_t seed, float* pivots, float* centroid, void* cuda_stream) {
  if (!d_base || !chunk_offsets || !pivots || !centroid || n == 0 || D == 0 || n_chunks == 0) { p_err = "bad argument"; return BANG_E_ARG; }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  switch (dtype) {
    case BANG_DT_FLOAT: return pq_train_run<float>((const float*)d_base, n, D, chunk_offsets, n_chunks, iters, max_train, seed, pivots, centroid, st);
    case BANG_DT_INT8: return pq_train_run<int8_t>((const int8_t*)d_base, n, D, chunk_offsets, n_chunks, iters, max_train, seed, pivots, centroid, st);
    case BANG_DT_UINT8: return pq_train_run<uint8_t>((const uint8_t*)d_base, n, D, chunk_offsets, n_chunks, iters, max_train, seed, pivots, centroid, st);
    default: p_err = "bad dtype"; return BANG_E_ARG;
  }
}

extern "C" int bang_b200_pq_encode(bang_dtype_t dtype, const void* d_base, uint64_t n, uint32_t D, const float* pivots, const float* centroid,
                                   const uint32_t* chunk_offsets, uint32_t n_chunks, uint8_t* d_codes, void* cuda_stream) {
  if (!d_base || !pivots || !centroid || !chunk_offsets || !d_codes || n == 0 || D == 0 || n_chunks == 0) { p_err = "bad argument"; return BANG_E_ARG; }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  switch (dtype) {
    case BANG_DT_FLOAT: return pq_encode_run<float>((const float*)d_base, n, D, pivots, centroid, chunk_offsets, n_chunks, d_codes, st);
    case BANG_DT_INT8: return pq_encode_run<int8_t>((const int8_t*)d_base, n, D, pivots, centroid, chunk_offsets, n_chunks, d_codes, st);
    case BANG_DT_UINT8: return pq_encode_run<uint8_t>((const uint8_t*)d_base, n, D, pivots, centroid, chunk_offsets, n_chunks, d_codes, st);
    default: p_err = "bad dtype"; return BANG_E_ARG;
  }
}
