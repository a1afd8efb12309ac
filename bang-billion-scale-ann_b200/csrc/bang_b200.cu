// bang_b200.cu — C ABI (include/bang_b200.h) + host orchestration for the fused sm_100a search kernel.
//
// Mirrors the reference's BANGSearchInner<T> life cycle (BANG_Base/bang_search.cu:139-1068): load ->
// set_searchparams -> alloc -> init -> query -> free -> unload.  Where the reference keeps the graph in
// host RAM and crosses PCIe three times per hop (bang_search.cu:709,827-838), the whole index lives in
// HBM here (optionally row-sharded over several GPUs) and one kernel launch runs the entire search.
#include "bang_b200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "loader.h"
#include "shard_mem.h"
#include "search_kernel.cuh"
#include "search_inst.cuh"

using namespace bang;

namespace bang {
// ------------------------------------------------------------------------------------------------
// load-time repack kernels: file layouts -> HBM layouts
// ------------------------------------------------------------------------------------------------
// `_disk.bin` entry: [T vec[D]][u32 degree][u32 nbr[R]] (entry_len bytes, unaligned)  ->
// HBM row: [u32 nbr[64], unused slots = kNoNbr][vec, zero padded to 16 B][pad to row_stride]
// Adjacency rows are validated while they are repacked (the reference only asserts the first and last entry of
// the file, bang_search.cu:335-345): a degree above R, a neighbour id >= N, or the same id twice in one row are
// counted in bad[0..2] and fail the load with BANG_E_FORMAT — the kernel's visited filter tests a whole row
// against the state before the row (snapshot semantics), so a duplicate would be admitted twice.
__device__ __forceinline__ void validate_row(uint32_t ida, uint32_t idb, uint64_t N, unsigned long long* bad) {
  // lane l holds slots l and l+32 (kNoNbr = unused)
  const uint32_t lane = threadIdx.x & 31;
  bool range = (ida != kNoNbr && ida >= N) || (idb != kNoNbr && idb >= N);
  bool dup = ida != kNoNbr && ida == idb;
  for (int j = 0; j < 32; ++j) {
    const uint32_t oa = __shfl_sync(0xffffffffu, ida, j), ob = __shfl_sync(0xffffffffu, idb, j);
    if ((uint32_t)j != lane) dup = dup || (ida != kNoNbr && (ida == oa || ida == ob)) || (idb != kNoNbr && (idb == oa || idb == ob));
  }
  if (__any_sync(0xffffffffu, range) && lane == 0) atomicAdd(bad + 1, 1ull);
  if (__any_sync(0xffffffffu, dup) && lane == 0) atomicAdd(bad + 2, 1ull);
}

// the slot block of a row (search_kernel.cuh kSlotBytes): neighbour i's two visited-filter slots as two words at 256 + 8 i
__device__ __forceinline__ void write_slot_pair(uint8_t* row, uint32_t i, uint32_t id) {
  uint2 w = make_uint2(0u, 0u);  // (an unused neighbour slot: any valid slot, never accepted)
  if (id != kNoNbr) w = make_uint2(vis_slot_word(hash1(id)), vis_slot_word(hash2(id)));
  reinterpret_cast<uint2*>(row + kAdjBytes)[i] = w;
}

__global__ void repack_rows_kernel(const uint8_t* __restrict__ src, uint64_t entry_len, uint32_t vec_bytes, uint32_t R,
                                   uint8_t* __restrict__ dst, uint32_t row_stride, uint64_t first_id, uint64_t n_ids,
                                   uint32_t shard, uint32_t n_shards, uint64_t N, unsigned long long* __restrict__ bad, uint32_t slot_block) {
  // one warp per node
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= n_ids) return;
  const uint64_t id = first_id + warp;
  if (id % n_shards != shard) return;
  const uint8_t* e = src + warp * entry_len;
  uint8_t* d = dst + (id / n_shards) * (uint64_t)row_stride;
  uint32_t deg = 0;
  for (int b = 0; b < 4; ++b) deg |= (uint32_t)e[vec_bytes + b] << (8 * b);
  if (deg > R) {
    if (lane == 0) atomicAdd(bad + 0, 1ull);
    deg = R;
  }
  uint32_t ids[2];
  for (uint32_t h = 0; h < 2; ++h) {
    const uint32_t i = lane + 32 * h;
    uint32_t v = kNoNbr;
    if (i < deg) {
      v = 0;
      for (int b = 0; b < 4; ++b) v |= (uint32_t)e[vec_bytes + 4 + 4 * i + b] << (8 * b);
    }
    reinterpret_cast<uint32_t*>(d)[i] = v;
    if (slot_block) write_slot_pair(d, i, v);
    ids[h] = v;
  }
  validate_row(ids[0], ids[1], N, bad);
  const uint32_t vo = kAdjBytes + (slot_block ? kSlotBytes : 0);
  for (uint32_t i = lane; i < row_stride - vo; i += 32) d[vo + i] = i < vec_bytes ? e[i] : (uint8_t)0;
}

// PQ code row [m] -> permuted row [code_stride]: byte (32g + 4t + b) = chunk (32g + 8b + t)
__global__ void repack_codes_kernel(const uint8_t* __restrict__ src, uint32_t m, uint8_t* __restrict__ dst,
                                    uint32_t code_stride, uint64_t n_rows) {
  const uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t row = idx / code_stride;
  const uint32_t p = (uint32_t)(idx % code_stride);
  if (row >= n_rows) return;
  const uint32_t g = p >> 5, t = (p & 31) >> 2, b = p & 3;
  const uint32_t c = 32 * g + 8 * b + t;
  dst[row * code_stride + p] = c < m ? src[row * m + c] : (uint8_t)0;
}

// the same permutation for rows with scattered destination ids (bang_b200_load_device_codes_at)
__global__ void repack_codes_at_kernel(const uint8_t* __restrict__ src, uint32_t m, uint8_t* __restrict__ dst, uint32_t code_stride,
                                       const uint32_t* __restrict__ ids, uint64_t n_rows, uint64_t N) {
  const uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t row = idx / code_stride;
  const uint32_t p = (uint32_t)(idx % code_stride);
  if (row >= n_rows) return;
  const uint64_t id = ids[row];
  if (id >= N) return;
  const uint32_t g = p >> 5, t = (p & 31) >> 2, b = p & 3;
  const uint32_t c = 32 * g + 8 * b + t;
  dst[id * code_stride + p] = c < m ? src[row * m + c] : (uint8_t)0;
}

// device arrays -> HBM rows (bang_b200_load_device_rows): one warp per node
__global__ void pack_rows_kernel(const uint8_t* __restrict__ vec, uint32_t vec_bytes, const uint32_t* __restrict__ adj,
                                 uint8_t* __restrict__ dst, uint32_t row_stride, uint64_t n, uint64_t N,
                                 unsigned long long* __restrict__ bad, uint32_t slot_block) {
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= n) return;
  uint8_t* d = dst + warp * (uint64_t)row_stride;
  const uint32_t ida = adj[warp * kMaxR + lane], idb = adj[warp * kMaxR + 32 + lane];
  reinterpret_cast<uint32_t*>(d)[lane] = ida;
  reinterpret_cast<uint32_t*>(d)[32 + lane] = idb;
  if (slot_block) { write_slot_pair(d, lane, ida); write_slot_pair(d, 32 + lane, idb); }
  validate_row(ida, idb, N, bad);
  const uint32_t vo = kAdjBytes + (slot_block ? kSlotBytes : 0);
  for (uint32_t i = lane; i < row_stride - vo; i += 32) d[vo + i] = i < vec_bytes ? vec[warp * vec_bytes + i] : (uint8_t)0;
}

}  // namespace bang

static thread_local std::string g_err;
extern "C" const char* bang_b200_last_error(void) { return g_err.c_str(); }

static int set_err(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      return set_err(BANG_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                                      std::to_string(__LINE__) + ")");                                   \
  } while (0)

struct bang_b200_ctx {
  int dtype = BANG_DT_UINT8, mode = BANG_MODE_BASE, device = 0;
  int shard = 0, n_shards = 1;
  bool loaded = false, allocated = false;
  // index
  uint64_t N = 0, medoid = 0, entry_len = 0;
  uint32_t D = 0, R = 0, n_chunks = 0;
  uint32_t vec_bytes = 0, vec_units = 0, row_stride = 0, code_stride = 0;
  uint64_t rows_local = 0;
  uint8_t* d_rows = nullptr;      // = rows_mem.ptr
  ShardMem rows_mem;              // this process's shard of the rows (cudaMalloc, or VMM when BANG_B200_SHARD_VMM=1)
  const uint8_t* rows[kMaxShards] = {nullptr};
  ShardMem imported[kMaxShards];  // peers' shards mapped into this process
  uint8_t* d_codes = nullptr;
  float* d_pivT = nullptr;
  float* d_piv = nullptr;
  uint32_t chunk4 = 0;
  float* d_centroid = nullptr;
  uint32_t* d_chunk_off = nullptr;
  uint64_t device_bytes = 0;
  unsigned long long* d_bad = nullptr;  // row validation counters: degree > R, id >= N, duplicate id (see validate_row)
  bool piv_global = false;              // the pivot table does not fit in shared memory: read it from global/L2
  bool code_prefetch = false;           // speculative L2 prefetch of every neighbour's PQ code row (BANG_B200_CODE_PREFETCH=1): no longer
                                        // pays (same time, +47 % DRAM bytes at 10^7 points, profiles/r2u_dram_bytes_prefetch_l2fetch.log)
  bool prehash = false;                 // rows carry their neighbours' visited-filter slots (kSlotBytes; decided at load, see set_row_geometry)
  // params
  int k = 0, L = 0, distfn = BANG_DIST_L2, dists_layout = BANG_DISTS_RANK_MAJOR;
  // per-alloc scratch
  int Qcap = 0;
  void* d_queries = nullptr;
  uint64_t* d_ids = nullptr;
  float* d_dists = nullptr;
  uint32_t* d_bloom = nullptr;
  uint32_t* d_counter = nullptr;
  uint32_t *d_hops = nullptr, *d_sumdeg = nullptr, *d_npass = nullptr;
  long long* d_phase = nullptr;
  float* h_dists = nullptr;  // pinned staging for the layout transpose
  cudaStream_t stream = nullptr;
  uint32_t* d_candlog = nullptr;  // expanded-node logs of the resident query warps (PQ modes)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_busy = nullptr;  // recorded after every launch: the next launch (on any stream) waits for it, because
  bool busy = false;              // the work counter, the visited filters and the candidate logs are shared scratch
  int grid = 0, ctas_per_sm = 0, warps_per_cta = 0, sm_count = 0;
  size_t smem = 0;
  int lastQ = 0;
  bang_b200_timing_t timing = {};
};

static size_t elem_size(int dtype) { return dtype == BANG_DT_FLOAT ? 4 : 1; }

// ------------------------------------------------------------------------------------------------
// kernel dispatch
// ------------------------------------------------------------------------------------------------
// The search kernels are instantiated per element type in search_inst_{u8,i8,f32}.cu (compiled in parallel); each
// of those files exports one lookup function (search_inst.cuh).
static search_fn_t pick_kernel(int dtype, int mode, uint32_t cs, int warps_per_cta, bool ph) {
  const int wpc = wpc_variant(mode, warps_per_cta);
  switch (dtype) {
    case BANG_DT_FLOAT: return search_kernel_f32(mode, cs, wpc, ph);
    case BANG_DT_INT8: return search_kernel_i8(mode, cs, wpc, ph);
    default: return search_kernel_u8(mode, cs, wpc, ph);
  }
}
static table_fn_t pick_table_kernel(int dtype) {
  switch (dtype) {
    case BANG_DT_FLOAT: return table_kernel_f32();
    case BANG_DT_INT8: return table_kernel_i8();
    default: return table_kernel_u8();
  }
}
static LaunchGeom geometry_for(const bang_b200_ctx* c, uint32_t L, uint32_t cand_cap, size_t optin, size_t per_sm, int max_warps,
                               bool piv_global, bool keep_l1) {
  const uint32_t piv_row = pivot_row_floats(c->D, piv_global ? 0 : c->chunk4);
  switch (c->dtype) {
    case BANG_DT_FLOAT: return launch_geometry<float>(c->mode, piv_row, c->n_chunks, c->vec_units, L, cand_cap, optin, per_sm, max_warps, piv_global, keep_l1);
    case BANG_DT_INT8: return launch_geometry<int8_t>(c->mode, piv_row, c->n_chunks, c->vec_units, L, cand_cap, optin, per_sm, max_warps, piv_global, keep_l1);
    default: return launch_geometry<uint8_t>(c->mode, piv_row, c->n_chunks, c->vec_units, L, cand_cap, optin, per_sm, max_warps, piv_global, keep_l1);
  }
}
static uint32_t max_iter_for(int mode, int L) {
  // bang_search.cu:53,603 (L+50) / BANG_Inmemory parANN.cu:30 (L+120) / BANG_Exactdistance parANN.cu:42 (4L+20)
  return mode == BANG_MODE_BASE ? L + 50 : (mode == BANG_MODE_INMEMORY ? L + 120 : 4 * L + 20);
}

// ------------------------------------------------------------------------------------------------
// create / destroy
// ------------------------------------------------------------------------------------------------
extern "C" int bang_b200_create(bang_handle_t* out, bang_dtype_t dtype, bang_mode_t mode, int device) {
  if (!out) return set_err(BANG_E_ARG, "null handle pointer");
  if (dtype < BANG_DT_INT8 || dtype > BANG_DT_FLOAT) return set_err(BANG_E_ARG, "bad dtype");
  if (mode < BANG_MODE_BASE || mode > BANG_MODE_EXACTDISTANCE) return set_err(BANG_E_ARG, "bad mode");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (ndev == 0) return set_err(BANG_E_CUDA, "no CUDA device: the sm_100a kernels are the only search path");
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  if (device >= ndev) return set_err(BANG_E_ARG, "device index out of range");
  bang_b200_ctx* c = new bang_b200_ctx();
  c->dtype = dtype;
  c->mode = mode;
  c->device = device;
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  *out = c;
  return BANG_OK;
}

extern "C" void bang_b200_destroy(bang_handle_t h) {
  if (!h) return;
  bang_b200_free(h);
  bang_b200_unload(h);
  delete h;
}

extern "C" int bang_b200_set_sharding(bang_handle_t h, int shard, int n_shards) {
  if (!h) return set_err(BANG_E_ARG, "null handle");
  if (h->loaded) return set_err(BANG_E_STATE, "set_sharding must precede load");
  if (n_shards < 1 || n_shards > kMaxShards || shard < 0 || shard >= n_shards) return set_err(BANG_E_ARG, "bad shard spec");
  h->shard = shard;
  h->n_shards = n_shards;
  return BANG_OK;
}

// ------------------------------------------------------------------------------------------------
// load
// ------------------------------------------------------------------------------------------------
static int upload_pq(bang_b200_ctx* c, const PQHost& pq) {
  const uint32_t D = c->D;
  // transpose pivots to [D][256] as the reference does at load (bang_search.cu:281-285)
  std::vector<float> pivT((size_t)D * 256);
  for (uint32_t row = 0; row < 256; ++row)
    for (uint32_t col = 0; col < D; ++col) pivT[(size_t)col * 256 + row] = pq.pivots[(size_t)row * D + col];
  if (pq.chunk_off.size() != c->n_chunks + 1) return set_err(BANG_E_FORMAT, "chunk offsets count != n_chunks+1");
  if (pq.chunk_off.back() != D) return set_err(BANG_E_FORMAT, "chunk offsets do not end at D");
  for (size_t i = 0; i + 1 < pq.chunk_off.size(); ++i)
    if (pq.chunk_off[i] > pq.chunk_off[i + 1]) return set_err(BANG_E_FORMAT, "chunk offsets not monotone");
  // 32 chunks of one size <= 4 (4: SIFT 128/32; 3: DEEP 96/32) select the CS = 4 kernels, whose pivot table has every
  // chunk zero-padded to 4 dimensions ([256][32][4], one 16-byte load per table entry); everything else the general kernel
  c->chunk4 = pq.chunk_off.size() > 1 ? pq.chunk_off[1] - pq.chunk_off[0] : 0;
  for (size_t i = 0; i + 1 < pq.chunk_off.size(); ++i)
    if (pq.chunk_off[i + 1] - pq.chunk_off[i] != c->chunk4) c->chunk4 = 0;
  if (c->chunk4 < 1 || c->chunk4 > 4 || c->n_chunks != 32) c->chunk4 = 0;  // the CS = 4 kernels assume one full 32-chunk group
  std::vector<float> padded;
  const std::vector<float>* table = &pq.pivots;
  if (c->chunk4) {
    padded.assign((size_t)256 * 128, 0.0f);
    for (uint32_t row = 0; row < 256; ++row)
      for (uint32_t ch = 0; ch < 32; ++ch)
        for (uint32_t e = 0; e < c->chunk4; ++e) padded[(size_t)row * 128 + ch * 4 + e] = pq.pivots[(size_t)row * D + ch * c->chunk4 + e];
    table = &padded;
  }
  CUDA_TRY(cudaMalloc(&c->d_piv, table->size() * 4));
  CUDA_TRY(cudaMemcpy(c->d_piv, table->data(), table->size() * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&c->d_pivT, pivT.size() * 4));
  CUDA_TRY(cudaMalloc(&c->d_centroid, (size_t)D * 4));
  CUDA_TRY(cudaMalloc(&c->d_chunk_off, pq.chunk_off.size() * 4));
  CUDA_TRY(cudaMemcpy(c->d_pivT, pivT.data(), pivT.size() * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->d_centroid, pq.centroid.data(), (size_t)D * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->d_chunk_off, pq.chunk_off.data(), pq.chunk_off.size() * 4, cudaMemcpyHostToDevice));
  c->device_bytes += (pivT.size() + table->size()) * 4 + (size_t)D * 4 + pq.chunk_off.size() * 4;
  return BANG_OK;
}

// streams a file region through a pinned buffer to the device and runs `consume(d_chunk, first_item, n_items)`
template <typename F>
static int stream_file(const std::string& path, uint64_t offset, uint64_t item_bytes, uint64_t n_items, F consume) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return set_err(BANG_E_IO, "Could not open " + path);
  if (fseeko(f, (off_t)offset, SEEK_SET) != 0) { fclose(f); return set_err(BANG_E_IO, "seek failed in " + path); }
  const uint64_t budget = 64ull << 20;
  const uint64_t per = std::max<uint64_t>(1, budget / item_bytes);
  uint8_t* h_buf = nullptr;
  uint8_t* d_buf = nullptr;
  cudaError_t e = cudaMallocHost(&h_buf, per * item_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&d_buf, per * item_bytes);
  int rc = BANG_OK;
  if (e != cudaSuccess) rc = set_err(BANG_E_CUDA, std::string("staging alloc: ") + cudaGetErrorString(e));
  for (uint64_t first = 0; rc == BANG_OK && first < n_items; first += per) {
    const uint64_t n = std::min(per, n_items - first);
    if (fread(h_buf, item_bytes, n, f) != n) { rc = set_err(BANG_E_FORMAT, "unexpected end of file in " + path); break; }
    e = cudaMemcpy(d_buf, h_buf, n * item_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { rc = set_err(BANG_E_CUDA, std::string("H2D: ") + cudaGetErrorString(e)); break; }
    rc = consume(d_buf, first, n);
    if (rc == BANG_OK) {
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) rc = set_err(BANG_E_CUDA, std::string("repack: ") + cudaGetErrorString(e));
    }
  }
  if (h_buf) cudaFreeHost(h_buf);
  if (d_buf) cudaFree(d_buf);
  fclose(f);
  return rc;
}

static int rows_check_begin(bang_b200_ctx* c) {
  if (!c->d_bad) CUDA_TRY(cudaMalloc(&c->d_bad, 3 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemset(c->d_bad, 0, 3 * sizeof(unsigned long long)));
  return BANG_OK;
}
static int rows_check_end(bang_b200_ctx* c) {
  unsigned long long bad[3] = {0, 0, 0};
  CUDA_TRY(cudaMemcpy(bad, c->d_bad, sizeof(bad), cudaMemcpyDeviceToHost));
  if (bad[0] || bad[1] || bad[2])
    return set_err(BANG_E_FORMAT, "graph rows are ill-formed: " + std::to_string(bad[0]) + " with degree > R, " + std::to_string(bad[1]) +
                                      " with a neighbour id >= N, " + std::to_string(bad[2]) + " with a repeated neighbour id");
  return BANG_OK;
}

// Row geometry.  PQ modes with 32 uniform chunks (the CS = 4 kernels) can store, next to the 64 neighbour ids, the two
// visited-filter slots of every neighbour (512 bytes per row): the kernel then reads them with the adjacency instead of
// evaluating two 64-bit hashes + modulus per neighbour (12 % of a hop's instructions).  Opt-in, BANG_B200_PREHASH=1:
// measured on one B200 it buys 2-4 % of time for twice the DRAM bytes and 1.8-2.3 x the row memory, because the hash
// arithmetic mostly fills slots in which the warp would otherwise wait for memory (profiles/r2_hop_diet.md).  Call after
// the PQ tables are uploaded (chunk4).
static void set_row_geometry(bang_b200_ctx* c) {
  c->vec_bytes = c->D * (uint32_t)elem_size(c->dtype);
  c->vec_units = (c->vec_bytes + 15) / 16;
  c->rows_local = (c->N + c->n_shards - 1 - c->shard) / c->n_shards;
  const uint32_t plain = (uint32_t)align_up(kAdjBytes + (size_t)c->vec_units * 16, 32);
  const uint32_t wide = (uint32_t)align_up(kAdjBytes + kSlotBytes + (size_t)c->vec_units * 16, 32);
  bool want = false;
  if (c->mode != BANG_MODE_EXACTDISTANCE && c->chunk4 != 0)
    if (const char* e = getenv("BANG_B200_PREHASH")) want = atoi(e) != 0;
  c->prehash = want;
  c->row_stride = want ? wide : plain;
}

static int load_graph(bang_b200_ctx* c, const std::string& disk_path) {
  uint64_t sz = 0;
  if (!file_size(disk_path, &sz, &g_err)) return BANG_E_IO;
  if (sz != c->N * c->entry_len)
    return set_err(BANG_E_FORMAT, "Graph Index File size " + std::to_string(sz) + " != N*entry_len " +
                                      std::to_string(c->N * c->entry_len));
  set_row_geometry(c);
  const size_t bytes = (size_t)c->rows_local * c->row_stride;
  {
    std::string why;
    const int rc = shard_alloc(&c->rows_mem, bytes, c->device, c->n_shards > 1 && shard_vmm_requested(), &why);
    if (rc != 0) return set_err(rc == -2 ? BANG_E_NOMEM : BANG_E_CUDA, "rows: " + why);
    c->d_rows = static_cast<uint8_t*>(c->rows_mem.ptr);
  }
  c->device_bytes += bytes;
  for (int s = 0; s < kMaxShards; ++s) c->rows[s] = nullptr;
  c->rows[c->shard] = c->d_rows;
  bang_b200_ctx* cc = c;
  int rc = rows_check_begin(c);
  if (rc != BANG_OK) return rc;
  rc = stream_file(disk_path, 0, c->entry_len, c->N, [cc](uint8_t* d_chunk, uint64_t first, uint64_t n) -> int {
    const uint64_t threads = n * 32;
    repack_rows_kernel<<<(unsigned)((threads + 255) / 256), 256>>>(d_chunk, cc->entry_len, cc->vec_bytes, cc->R, cc->d_rows,
                                                                   cc->row_stride, first, n, cc->shard, cc->n_shards, cc->N, cc->d_bad, cc->prehash ? 1u : 0u);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? BANG_OK : set_err(BANG_E_CUDA, std::string("repack_rows: ") + cudaGetErrorString(e));
  });
  return rc == BANG_OK ? rows_check_end(c) : rc;
}

static int load_codes(bang_b200_ctx* c, const std::string& path) {
  uint32_t n = 0, m = 0;
  if (!read_bin_header(path, 1, &n, &m, &g_err)) return BANG_E_FORMAT;
  if (n != c->N) return set_err(BANG_E_FORMAT, "PQ Compressed Vectors File holds " + std::to_string(n) + " points, graph has " + std::to_string(c->N));
  if (m == 0) return set_err(BANG_E_FORMAT, "PQ Compressed Vectors File has 0 chunks");
  c->n_chunks = m;
  c->code_stride = (m + 31) / 32 * 32;
  const size_t bytes = (size_t)c->N * c->code_stride;
  CUDA_TRY(cudaMalloc(&c->d_codes, bytes));
  c->device_bytes += bytes;
  bang_b200_ctx* cc = c;
  return stream_file(path, 8, m, c->N, [cc](uint8_t* d_chunk, uint64_t first, uint64_t nrows) -> int {
    const uint64_t total = nrows * cc->code_stride;
    repack_codes_kernel<<<(unsigned)((total + 255) / 256), 256>>>(d_chunk, cc->n_chunks,
                                                                  cc->d_codes + first * cc->code_stride, cc->code_stride, nrows);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? BANG_OK : set_err(BANG_E_CUDA, std::string("repack_codes: ") + cudaGetErrorString(e));
  });
}

static int check_common(bang_b200_ctx* c) {
  if (c->R != (uint32_t)kMaxR) return set_err(BANG_E_UNSUPPORTED, "graph degree bound must be 64 (MAX_R, bang_search.cu:35,190)");
  if (c->N == 0 || c->N > 0xFFFFFFFEull) return set_err(BANG_E_FORMAT, "bad dataset size");
  if (c->N > 0x80000000ull) return set_err(BANG_E_UNSUPPORTED, "more than 2^31 points (the worklist keeps its visited flag in the top bit of the id)");
  if (c->medoid >= c->N) return set_err(BANG_E_FORMAT, "medoid out of range");
  if (c->entry_len != (uint64_t)c->D * elem_size(c->dtype) + 4 + 4ull * c->R)
    return set_err(BANG_E_FORMAT, "index entry length does not match D*sizeof(T)+4+4R (wrong element type?)");
  return BANG_OK;
}

extern "C" int bang_b200_unload(bang_handle_t h);

static int finish_load(bang_b200_ctx* c, int rc) {
  if (rc != BANG_OK) {
    std::string keep = g_err;
    c->loaded = true;  // let unload release partial allocations
    bang_b200_unload(c);
    g_err = keep;
    return rc;
  }
  c->loaded = true;
  return BANG_OK;
}

extern "C" int bang_b200_load(bang_handle_t c, const char* prefix_c) {
  if (!c || !prefix_c) return set_err(BANG_E_ARG, "null argument");
  if (c->loaded) return set_err(BANG_E_STATE, "index already loaded");
  CUDA_TRY(cudaSetDevice(c->device));
  const std::string prefix(prefix_c);
  const std::string pivots = prefix + "_pq_pivots.bin", compressed = prefix + "_pq_compressed.bin";
  const std::string disk = prefix + "_disk.bin", meta = prefix + "_disk_metadata.bin";
  GraphMeta gm;
  if (!read_graph_meta(meta, &gm, &g_err)) return BANG_E_IO;
  c->N = gm.N; c->D = gm.D; c->R = gm.R; c->medoid = gm.medoid; c->entry_len = gm.entry_len;
  c->device_bytes = 0;
  // The metadata's datatype word follows the converter's numbering (bang_preprocess.py:42-51: 0 int8, 1 uint8, 2 float),
  // the same as bang_dtype_t.  The reference only prints it; here an index of another element type is refused —
  // int8 and uint8 entries have the same length, so nothing else would notice.
  if (gm.dtype >= BANG_DT_INT8 && gm.dtype <= BANG_DT_FLOAT && gm.dtype != c->dtype)
    return set_err(BANG_E_FORMAT, "index element type (metadata datatype " + std::to_string(gm.dtype) + ") differs from the handle's (" +
                                      std::to_string(c->dtype) + ")");
  int rc = check_common(c);
  if (rc != BANG_OK) return rc;
  if (c->mode != BANG_MODE_EXACTDISTANCE) {
    rc = load_codes(c, compressed);
    if (rc == BANG_OK) {
      PQHost pq;
      if (!read_pq_pivots_new(pivots, c->D, c->n_chunks, &pq, &g_err)) rc = BANG_E_IO;
      else rc = upload_pq(c, pq);
    }
    if (rc != BANG_OK) return finish_load(c, rc);
  }
  return finish_load(c, load_graph(c, disk));
}

extern "C" int bang_b200_load_files(bang_handle_t c, const char* pq_pivots_bin, const char* pq_compressed_bin,
                                    const char* disk_bin, const char* chunk_offsets_bin, const char* centroid_bin,
                                    uint64_t N, uint32_t D, uint64_t medoid) {
  if (!c || !disk_bin) return set_err(BANG_E_ARG, "null argument");
  if (c->loaded) return set_err(BANG_E_STATE, "index already loaded");
  CUDA_TRY(cudaSetDevice(c->device));
  c->N = N; c->D = D; c->R = kMaxR; c->medoid = medoid;
  c->entry_len = (uint64_t)D * elem_size(c->dtype) + 4 + 4ull * kMaxR;
  c->device_bytes = 0;
  int rc = check_common(c);
  if (rc != BANG_OK) return rc;
  if (c->mode != BANG_MODE_EXACTDISTANCE) {
    if (!pq_pivots_bin || !pq_compressed_bin || !chunk_offsets_bin || !centroid_bin)
      return set_err(BANG_E_ARG, "PQ file names are required in Base/Inmemory mode");
    rc = load_codes(c, pq_compressed_bin);
    if (rc == BANG_OK) {
      PQHost pq;
      if (!read_pq_pivots_old(pq_pivots_bin, centroid_bin, chunk_offsets_bin, D, &pq, &g_err)) rc = BANG_E_IO;
      else rc = upload_pq(c, pq);
    }
    if (rc != BANG_OK) return finish_load(c, rc);
  }
  return finish_load(c, load_graph(c, disk_bin));
}

// ---- device-resident load ---------------------------------------------------------------------
extern "C" int bang_b200_load_device_begin(bang_handle_t c, uint64_t N, uint32_t D, uint64_t medoid, uint32_t n_chunks,
                                           const float* pivots, const float* centroid, const uint32_t* chunk_offsets) {
  if (!c) return set_err(BANG_E_ARG, "null handle");
  if (c->loaded) return set_err(BANG_E_STATE, "index already loaded");
  CUDA_TRY(cudaSetDevice(c->device));
  c->N = N; c->D = D; c->R = kMaxR; c->medoid = medoid;
  c->entry_len = (uint64_t)D * elem_size(c->dtype) + 4 + 4ull * kMaxR;
  c->device_bytes = 0;
  int rc = check_common(c);
  if (rc != BANG_OK) return rc;
  c->loaded = true;  // from here on bang_b200_unload releases whatever was allocated
  if (c->mode != BANG_MODE_EXACTDISTANCE) {
    if (!pivots || !centroid || !chunk_offsets || n_chunks == 0) { bang_b200_unload(c); return set_err(BANG_E_ARG, "PQ arrays are required in Base/Inmemory mode"); }
    c->n_chunks = n_chunks;
    c->code_stride = (n_chunks + 31) / 32 * 32;
    PQHost pq;
    pq.pivots.assign(pivots, pivots + (size_t)256 * D);
    pq.centroid.assign(centroid, centroid + D);
    pq.chunk_off.assign(chunk_offsets, chunk_offsets + n_chunks + 1);
    rc = upload_pq(c, pq);
    if (rc == BANG_OK) {
      cudaError_t e = cudaMalloc(&c->d_codes, (size_t)N * c->code_stride);
      if (e != cudaSuccess) rc = set_err(BANG_E_NOMEM, std::string("codes: ") + cudaGetErrorString(e));
      else c->device_bytes += (size_t)N * c->code_stride;
    }
    if (rc != BANG_OK) { std::string k = g_err; bang_b200_unload(c); g_err = k; return rc; }
  }
  set_row_geometry(c);
  const size_t bytes = (size_t)c->rows_local * c->row_stride;
  {
    std::string why;
    const int rc = shard_alloc(&c->rows_mem, bytes, c->device, c->n_shards > 1 && shard_vmm_requested(), &why);
    if (rc != 0) { bang_b200_unload(c); return set_err(rc == -2 ? BANG_E_NOMEM : BANG_E_CUDA, "rows: " + why); }
    c->d_rows = static_cast<uint8_t*>(c->rows_mem.ptr);
  }
  c->device_bytes += bytes;
  for (int s = 0; s < kMaxShards; ++s) c->rows[s] = nullptr;
  c->rows[c->shard] = c->d_rows;
  rc = rows_check_begin(c);
  if (rc != BANG_OK) { std::string k = g_err; bang_b200_unload(c); g_err = k; return rc; }
  return BANG_OK;
}

extern "C" int bang_b200_load_device_rows(bang_handle_t c, uint64_t first_local_row, uint64_t n_rows, const void* d_vectors,
                                          const uint32_t* d_adj) {
  if (!c || !d_vectors || !d_adj) return set_err(BANG_E_ARG, "null argument");
  if (!c->loaded || !c->d_rows) return set_err(BANG_E_STATE, "bang_b200_load_device_begin first");
  if (first_local_row + n_rows > c->rows_local) return set_err(BANG_E_ARG, "row range exceeds this shard");
  if (n_rows == 0) return BANG_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  const uint64_t threads = n_rows * 32;
  pack_rows_kernel<<<(unsigned)((threads + 255) / 256), 256>>>((const uint8_t*)d_vectors, c->vec_bytes, d_adj,
                                                             c->d_rows + first_local_row * c->row_stride, c->row_stride, n_rows,
                                                             c->N, c->d_bad, c->prehash ? 1u : 0u);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaDeviceSynchronize());
  return BANG_OK;
}

extern "C" int bang_b200_load_device_codes(bang_handle_t c, uint64_t first_id, uint64_t n, const uint8_t* d_codes) {
  if (!c || !d_codes) return set_err(BANG_E_ARG, "null argument");
  if (!c->loaded || !c->d_codes) return set_err(BANG_E_STATE, "bang_b200_load_device_begin (PQ mode) first");
  if (first_id + n > c->N) return set_err(BANG_E_ARG, "id range exceeds N");
  if (n == 0) return BANG_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  const uint64_t total = n * c->code_stride;
  repack_codes_kernel<<<(unsigned)((total + 255) / 256), 256>>>(d_codes, c->n_chunks, c->d_codes + first_id * c->code_stride,
                                                                c->code_stride, n);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaDeviceSynchronize());
  return BANG_OK;
}

extern "C" int bang_b200_load_device_codes_at(bang_handle_t c, const uint32_t* d_ids, uint64_t n, const uint8_t* d_codes) {
  if (!c || !d_codes || !d_ids) return set_err(BANG_E_ARG, "null argument");
  if (!c->loaded || !c->d_codes) return set_err(BANG_E_STATE, "bang_b200_load_device_begin (PQ mode) first");
  if (n == 0) return BANG_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  const uint64_t total = n * c->code_stride;
  repack_codes_at_kernel<<<(unsigned)((total + 255) / 256), 256>>>(d_codes, c->n_chunks, c->d_codes, c->code_stride, d_ids, n, c->N);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaDeviceSynchronize());
  return BANG_OK;
}

extern "C" int bang_b200_load_device_end(bang_handle_t c) {
  if (!c) return set_err(BANG_E_ARG, "null handle");
  if (!c->loaded || !c->d_rows) return set_err(BANG_E_STATE, "bang_b200_load_device_begin first");
  CUDA_TRY(cudaDeviceSynchronize());
  const int rc = rows_check_end(c);
  if (rc != BANG_OK) { std::string k = g_err; bang_b200_unload(c); g_err = k; }
  return rc;
}

extern "C" int bang_b200_unload(bang_handle_t c) {
  if (!c) return set_err(BANG_E_ARG, "null handle");
  if (!c->loaded) return BANG_OK;
  cudaSetDevice(c->device);
  for (int s = 0; s < kMaxShards; ++s) shard_release(&c->imported[s]);
  shard_release(&c->rows_mem);
  cudaFree(c->d_codes); cudaFree(c->d_pivT); cudaFree(c->d_piv); cudaFree(c->d_centroid); cudaFree(c->d_chunk_off);
  cudaFree(c->d_bad); c->d_bad = nullptr;
  c->d_rows = nullptr; c->d_codes = nullptr; c->d_pivT = nullptr; c->d_piv = nullptr; c->d_centroid = nullptr; c->d_chunk_off = nullptr;
  c->loaded = false;
  c->device_bytes = 0;
  return BANG_OK;
}

extern "C" int bang_b200_export_shard(bang_handle_t c, void* out64) {
  if (!c || !out64) return set_err(BANG_E_ARG, "null argument");
  if (!c->loaded) return set_err(BANG_E_STATE, "load first");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  if (c->rows_mem.vmm) return set_err(BANG_E_STATE, "the rows were allocated with BANG_B200_SHARD_VMM=1: use bang_b200_export_shard_fd");
  cudaIpcMemHandle_t hnd;
  CUDA_TRY(cudaIpcGetMemHandle(&hnd, c->d_rows));
  memcpy(out64, &hnd, 64);
  return BANG_OK;
}

extern "C" int bang_b200_import_shard(bang_handle_t c, int shard, const void* in64) {
  if (!c || !in64) return set_err(BANG_E_ARG, "null argument");
  if (!c->loaded) return set_err(BANG_E_STATE, "load first");
  if (shard < 0 || shard >= c->n_shards || shard == c->shard) return set_err(BANG_E_ARG, "bad shard index");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaIpcMemHandle_t hnd;
  memcpy(&hnd, in64, 64);
  void* p = nullptr;
  CUDA_TRY(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
  shard_release(&c->imported[shard]);
  c->imported[shard].ptr = p;
  c->imported[shard].imported = true;
  c->rows[shard] = (const uint8_t*)p;
  return BANG_OK;
}

extern "C" int bang_b200_export_shard_fd(bang_handle_t c, int* fd_out, uint64_t* bytes_out) {
  if (!c || !fd_out || !bytes_out) return set_err(BANG_E_ARG, "null argument");
  if (!c->loaded) return set_err(BANG_E_STATE, "load first");
  std::string why;
  if (shard_export_fd(&c->rows_mem, fd_out, &why) != 0) return set_err(BANG_E_STATE, why);
  *bytes_out = c->rows_mem.bytes;
  return BANG_OK;
}

extern "C" int bang_b200_import_shard_fd(bang_handle_t c, int shard, int fd, uint64_t bytes) {
  if (!c || fd < 0) return set_err(BANG_E_ARG, "bad argument");
  if (!c->loaded) return set_err(BANG_E_STATE, "load first");
  if (shard < 0 || shard >= c->n_shards || shard == c->shard) return set_err(BANG_E_ARG, "bad shard index");
  CUDA_TRY(cudaSetDevice(c->device));
  shard_release(&c->imported[shard]);
  std::string why;
  if (shard_import_fd(&c->imported[shard], fd, (size_t)bytes, c->device, &why) != 0) return set_err(BANG_E_CUDA, why);
  c->rows[shard] = static_cast<const uint8_t*>(c->imported[shard].ptr);
  return BANG_OK;
}

extern "C" int bang_b200_info(bang_handle_t c, bang_b200_info_t* out) {
  if (!c || !out) return set_err(BANG_E_ARG, "null argument");
  if (!c->loaded) return set_err(BANG_E_STATE, "no index loaded");
  out->N = c->N; out->medoid = c->medoid; out->entry_len = c->entry_len;
  out->D = c->D; out->R = c->R; out->n_chunks = c->n_chunks;
  out->dtype = c->dtype; out->mode = c->mode; out->device_bytes = c->device_bytes;
  out->row_stride = c->row_stride; out->slot_block = c->prehash ? 1u : 0u;
  return BANG_OK;
}

// ------------------------------------------------------------------------------------------------
// search params / alloc / init / free
// ------------------------------------------------------------------------------------------------
extern "C" int bang_b200_set_searchparams(bang_handle_t c, int recall, int worklist_length, bang_distfn_t dist) {
  if (!c) return set_err(BANG_E_ARG, "null handle");
  if (recall <= 0) return set_err(BANG_E_ARG, "recall (k) must be positive");
  if (worklist_length < recall) return set_err(BANG_E_ARG, "WorkList Length must be at least recall_at");  // test_driver.cpp:394-398
  if (worklist_length > BANG_B200_MAX_L) return set_err(BANG_E_ARG, "worklist_length exceeds MAX_L (512)");   // bang_search.cu:439
  if (dist != BANG_DIST_L2 && dist != BANG_DIST_MIPS) return set_err(BANG_E_ARG, "bad distance function");
  if (c->allocated) return set_err(BANG_E_STATE, "set_searchparams must precede alloc (bang_search.cu:370,379-384)");
  c->k = recall;
  c->L = worklist_length;
  c->distfn = dist;
  return BANG_OK;
}

extern "C" int bang_b200_set_dists_layout(bang_handle_t c, bang_dists_layout_t layout) {
  if (!c) return set_err(BANG_E_ARG, "null handle");
  c->dists_layout = layout;
  return BANG_OK;
}

static int alloc_impl(bang_b200_ctx* c, int Q) {
  const uint32_t max_iter = max_iter_for(c->mode, c->L);
  int max_optin = 0, per_sm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  CUDA_TRY(cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, c->device));
  // Resident queries per SM.  PQ modes: one CTA per SM around one pivot table with up to 24 query warps (80
  // registers each) — 16 (128 registers) and 32 (64 registers) are within 2-7 % on C2 / DEEP shapes, 24 is the best
  // of the three on both; Exactdistance: 2 CTAs of 16.  BANG_B200_WARPS_PER_SM overrides (up to 32).
  int max_warps = c->mode == BANG_MODE_EXACTDISTANCE ? 32 : 24;   // measured: profiles/r2_concurrency.md
  bool keep_l1 = true;  // (an explicit warp count is taken as given)
  if (const char* e = getenv("BANG_B200_WARPS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v <= 64) { max_warps = v; keep_l1 = false; } }
  if (const char* e = getenv("BANG_B200_CODE_PREFETCH")) c->code_prefetch = atoi(e) != 0;
  // A pivot table that does not fit next to one query's state (256 x D floats: D above ~215) stays in global memory
  // (L2-resident, read with plain loads by the generic-chunk kernel) instead of being refused.
  c->piv_global = false;
  LaunchGeom g = geometry_for(c, c->L, max_iter + 1, (size_t)max_optin, (size_t)per_sm, max_warps, false, keep_l1);
  if (g.warps_per_cta < 1 && c->mode != BANG_MODE_EXACTDISTANCE) {
    c->piv_global = true;
    g = geometry_for(c, c->L, max_iter + 1, (size_t)max_optin, (size_t)per_sm, max_warps, true, keep_l1);
  }
  if (g.warps_per_cta < 1)
    return set_err(BANG_E_UNSUPPORTED, "one query's state (D = " + std::to_string(c->D) + ", L = " + std::to_string(c->L) + ") does not fit in " +
                                           std::to_string(max_optin) + " B of shared memory");
  search_fn_t fn = pick_kernel(c->dtype, c->mode, (c->piv_global || !c->chunk4) ? 0 : 4, g.warps_per_cta, c->prehash);
  c->smem = g.smem;
  c->warps_per_cta = g.warps_per_cta;
  CUDA_TRY(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem));
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)fn, c->warps_per_cta * 32, c->smem));
  if (occ < 1) return set_err(BANG_E_CUDA, "kernel does not fit on an SM");
  c->ctas_per_sm = std::min(occ, g.ctas_per_sm);
  c->grid = std::min((Q + c->warps_per_cta - 1) / c->warps_per_cta, c->ctas_per_sm * c->sm_count);
  const size_t qbytes = (size_t)Q * c->D * elem_size(c->dtype);
  const size_t slots = (size_t)c->grid * c->warps_per_cta;  // resident query warps of the whole grid
  CUDA_TRY(cudaMalloc(&c->d_queries, qbytes));
  CUDA_TRY(cudaMalloc(&c->d_ids, (size_t)Q * c->k * sizeof(uint64_t)));
  CUDA_TRY(cudaMalloc(&c->d_dists, (size_t)Q * c->k * sizeof(float)));
  // [filter block areas of all resident warps (hot, L2-resident)][spill bitmaps (touched only by blocks that overflow)]
  CUDA_TRY(cudaMalloc(&c->d_bloom, slots * kBloomWords * 4));
  if (c->mode != BANG_MODE_EXACTDISTANCE) CUDA_TRY(cudaMalloc(&c->d_candlog, slots * (size_t)(max_iter + 1) * 4));
  CUDA_TRY(cudaMalloc(&c->d_counter, 4));
  CUDA_TRY(cudaMalloc(&c->d_hops, (size_t)Q * 4));
  CUDA_TRY(cudaMalloc(&c->d_sumdeg, (size_t)Q * 4));
  CUDA_TRY(cudaMalloc(&c->d_npass, (size_t)Q * 4));
#ifdef BANG_PHASE_TIMERS
  CUDA_TRY(cudaMalloc(&c->d_phase, (size_t)Q * PT_COUNT * 8));
#endif
  CUDA_TRY(cudaMallocHost(&c->h_dists, (size_t)Q * c->k * sizeof(float)));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&c->ev0));
  CUDA_TRY(cudaEventCreate(&c->ev1));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_busy, cudaEventDisableTiming));
  c->Qcap = Q;
  return BANG_OK;
}

extern "C" int bang_b200_alloc(bang_handle_t c, int Q) {
  if (!c) return set_err(BANG_E_ARG, "null handle");
  if (!c->loaded) return set_err(BANG_E_STATE, "bang_load must precede bang_alloc");
  if (c->k <= 0 || c->L <= 0) return set_err(BANG_E_STATE, "bang_set_searchparams must precede bang_alloc");
  if (c->allocated) return set_err(BANG_E_STATE, "already allocated; call bang_free first");
  if (Q <= 0) return set_err(BANG_E_ARG, "numQueries must be positive");
  for (int s = 0; s < c->n_shards; ++s)
    if (!c->rows[s]) return set_err(BANG_E_STATE, "graph shard " + std::to_string(s) + " has not been imported");
  CUDA_TRY(cudaSetDevice(c->device));
  c->allocated = true;  // from here on bang_b200_free releases whatever a failed attempt got hold of
  const int rc = alloc_impl(c, Q);
  if (rc != BANG_OK) {
    const std::string keep = g_err;
    bang_b200_free(c);
    cudaGetLastError();
    g_err = keep;
  }
  return rc;
}

extern "C" int bang_b200_init(bang_handle_t c, int Q) {
  // The reference re-zeroes the bloom filters/worklists and re-seeds the medoid here (bang_search.cu:440-506).
  // The fused kernel initialises all per-query state itself, so only the argument checks remain.
  if (!c) return set_err(BANG_E_ARG, "null handle");
  if (!c->allocated) return set_err(BANG_E_STATE, "bang_alloc must precede bang_init");
  if (Q <= 0 || Q > c->Qcap) return set_err(BANG_E_ARG, "numQueries exceeds the allocated batch");
  return BANG_OK;
}

extern "C" int bang_b200_free(bang_handle_t c) {
  if (!c) return set_err(BANG_E_ARG, "null handle");
  if (!c->allocated) return BANG_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();  // launches on callers' streams (bang_b200_query_device) included
  cudaFree(c->d_queries); cudaFree(c->d_ids); cudaFree(c->d_dists); cudaFree(c->d_bloom); cudaFree(c->d_counter);
  cudaFree(c->d_candlog);
  cudaFree(c->d_hops); cudaFree(c->d_sumdeg); cudaFree(c->d_npass); cudaFree(c->d_phase); c->d_phase = nullptr;
  if (c->h_dists) cudaFreeHost(c->h_dists);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev_busy) cudaEventDestroy(c->ev_busy);
  cudaGetLastError();
  c->d_queries = nullptr; c->d_ids = nullptr; c->d_dists = nullptr; c->d_bloom = nullptr; c->d_counter = nullptr;
  c->d_candlog = nullptr; c->ev_busy = nullptr; c->busy = false;
  c->d_hops = c->d_sumdeg = c->d_npass = nullptr; c->h_dists = nullptr; c->stream = nullptr; c->ev0 = c->ev1 = nullptr;
  c->allocated = false;
  c->Qcap = 0;
  return BANG_OK;
}

// ------------------------------------------------------------------------------------------------
// query
// ------------------------------------------------------------------------------------------------
static void fill_args(const bang_b200_ctx* c, SearchArgs* a, const void* d_queries, int Q, uint64_t* d_ids, float* d_dists) {
  memset(a, 0, sizeof(*a));
  for (int s = 0; s < kMaxShards; ++s) a->rows[s] = c->rows[s];
  a->n_shards = c->n_shards;
  a->shard_shift = 0;
  if (c->n_shards > 1 && (c->n_shards & (c->n_shards - 1)) == 0) while ((1u << a->shard_shift) < (uint32_t)c->n_shards) ++a->shard_shift;
  a->row_stride = c->row_stride;
  a->codes = c->d_codes;
  a->code_stride = c->code_stride;
  a->n_chunks = c->n_chunks;
  a->pivT = c->d_pivT;
  a->piv = c->d_piv;
  a->piv_row = pivot_row_floats(c->D, c->piv_global ? 0 : c->chunk4);
  a->chunk4 = c->chunk4;
  a->centroid = c->d_centroid;
  a->chunk_off = c->d_chunk_off;
  a->D = c->D;
  a->vec_units = c->vec_units;
  set_medoid(*a, (uint32_t)c->medoid, c->mode == BANG_MODE_EXACTDISTANCE);
  a->L = c->L;
  a->k = c->k;
  a->Q = Q;
  a->q_dim = c->D - (c->distfn == BANG_DIST_MIPS ? 1 : 0);  // MIPS_EXTRA_DIM, bang.h:31; bang_search.cu:631
  a->max_iter = max_iter_for(c->mode, c->L);
  a->cand_cap = a->max_iter + 1;
  a->queries = d_queries;
  a->out_ids = d_ids;
  a->out_dists = d_dists;
  a->bloom = c->d_bloom;
  a->cand_log = c->d_candlog;
  a->piv_global = c->piv_global ? 1u : 0u;
  a->code_prefetch = c->code_prefetch ? 1u : 0u;
  a->stop_on_empty_hop = c->mode == BANG_MODE_EXACTDISTANCE ? 1u : 0u;  // BANG_Exactdistance/parANN.cu:1593-1671 as built
  a->counter = c->d_counter;
  a->st_hops = c->d_hops;
  a->st_sumdeg = c->d_sumdeg;
  a->st_npass = c->d_npass;
  a->st_phase = c->d_phase;
}

static int launch_search(bang_b200_ctx* c, const void* d_queries, int Q, uint64_t* d_ids, float* d_dists, cudaStream_t st) {
  SearchArgs a;
  fill_args(c, &a, d_queries, Q, d_ids, d_dists);
  // one search in flight per handle: a launch on another stream is ordered after the previous one
  if (c->busy) CUDA_TRY(cudaStreamWaitEvent(st, c->ev_busy, 0));
  CUDA_TRY(cudaMemsetAsync(c->d_counter, 0, 4, st));
  const int grid = std::min((Q + c->warps_per_cta - 1) / c->warps_per_cta, c->grid);
  search_fn_t fn = pick_kernel(c->dtype, c->mode, (c->piv_global || !c->chunk4) ? 0 : 4, c->warps_per_cta, c->prehash);
  fn<<<grid, c->warps_per_cta * 32, c->smem, st>>>(a);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(c->ev_busy, st));
  c->busy = true;
  c->lastQ = Q;
  c->timing.launches = 1;
  c->timing.grid = grid;
  c->timing.block = c->warps_per_cta * 32;
  c->timing.smem_bytes = (uint32_t)c->smem;
  c->timing.ctas_per_sm = c->ctas_per_sm;
  return BANG_OK;
}

static int check_query_state(bang_b200_ctx* c, const void* q, int Q, const void* ids) {
  if (!c || !q || !ids) return set_err(BANG_E_ARG, "null argument");
  if (!c->allocated) return set_err(BANG_E_STATE, "bang_alloc/bang_init must precede bang_query");
  if (Q <= 0 || Q > c->Qcap) return set_err(BANG_E_ARG, "num_queries exceeds the allocated batch");
  return BANG_OK;
}

extern "C" int bang_b200_query(bang_handle_t c, const void* queries, int Q, uint64_t* ids, float* dists) {
  int rc = check_query_state(c, queries, Q, ids);
  if (rc != BANG_OK) return rc;
  CUDA_TRY(cudaSetDevice(c->device));
  const uint32_t q_dim = c->D - (c->distfn == BANG_DIST_MIPS ? 1 : 0);
  const size_t qbytes = (size_t)Q * q_dim * elem_size(c->dtype);
  CUDA_TRY(cudaMemcpyAsync(c->d_queries, queries, qbytes, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
  rc = launch_search(c, c->d_queries, Q, c->d_ids, c->d_dists, c->stream);
  if (rc != BANG_OK) return rc;
  CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  const size_t n = (size_t)Q * c->k;
  CUDA_TRY(cudaMemcpyAsync(ids, c->d_ids, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
  uint64_t d2h = n * sizeof(uint64_t);
  if (dists) {
    float* dst = c->dists_layout == BANG_DISTS_QUERY_MAJOR ? dists : c->h_dists;
    CUDA_TRY(cudaMemcpyAsync(dst, c->d_dists, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    d2h += n * sizeof(float);
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (dists && c->dists_layout == BANG_DISTS_RANK_MAJOR) {
    // the reference hands back the first Q*k floats of its rank-major matrix (bang_search.cu:999,1297)
    for (int q = 0; q < Q; ++q)
      for (int j = 0; j < c->k; ++j) dists[(size_t)j * Q + q] = c->h_dists[(size_t)q * c->k + j];
  }
  CUDA_TRY(cudaEventElapsedTime(&c->timing.kernel_ms, c->ev0, c->ev1));
  c->timing.h2d_bytes = qbytes;
  c->timing.d2h_bytes = d2h;
  return BANG_OK;
}

extern "C" int bang_b200_query_device(bang_handle_t c, const void* d_queries, int Q, uint64_t* d_ids, float* d_dists,
                                      void* cuda_stream) {
  int rc = check_query_state(c, d_queries, Q, d_ids);
  if (rc != BANG_OK) return rc;
  if (!d_dists) return set_err(BANG_E_ARG, "null dists");
  CUDA_TRY(cudaSetDevice(c->device));
  c->timing.h2d_bytes = c->timing.d2h_bytes = 0;
  c->timing.kernel_ms = 0.f;
  return launch_search(c, d_queries, Q, d_ids, d_dists, (cudaStream_t)cuda_stream);
}

extern "C" int bang_b200_last_stats(bang_handle_t c, uint32_t* hops, uint32_t* sum_deg, uint32_t* n_cand) {
  if (!c) return set_err(BANG_E_ARG, "null handle");
  if (!c->allocated || c->lastQ <= 0) return set_err(BANG_E_STATE, "no query has run");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaDeviceSynchronize());
  const size_t b = (size_t)c->lastQ * 4;
  if (hops) CUDA_TRY(cudaMemcpy(hops, c->d_hops, b, cudaMemcpyDeviceToHost));
  if (sum_deg) CUDA_TRY(cudaMemcpy(sum_deg, c->d_sumdeg, b, cudaMemcpyDeviceToHost));
  if (n_cand) CUDA_TRY(cudaMemcpy(n_cand, c->d_npass, b, cudaMemcpyDeviceToHost));
  return BANG_OK;
}

extern "C" int bang_b200_last_timing(bang_handle_t c, bang_b200_timing_t* out) {
  if (!c || !out) return set_err(BANG_E_ARG, "null argument");
  *out = c->timing;
  return BANG_OK;
}

// ------------------------------------------------------------------------------------------------
// stage 1 on its own
// ------------------------------------------------------------------------------------------------
extern "C" int bang_b200_pq_table(bang_handle_t c, const void* queries, int Q, float* tables) {
  if (!c || !queries || !tables) return set_err(BANG_E_ARG, "null argument");
  if (!c->loaded) return set_err(BANG_E_STATE, "no index loaded");
  if (c->mode == BANG_MODE_EXACTDISTANCE) return set_err(BANG_E_UNSUPPORTED, "Exactdistance mode has no PQ table");
  if (Q <= 0) return set_err(BANG_E_ARG, "numQueries must be positive");
  CUDA_TRY(cudaSetDevice(c->device));
  const uint32_t q_dim = c->D - (c->distfn == BANG_DIST_MIPS ? 1 : 0);
  const size_t qbytes = (size_t)Q * q_dim * elem_size(c->dtype);
  const size_t tbytes = (size_t)Q * c->n_chunks * 256 * sizeof(float);
  void* d_q = nullptr;
  float* d_t = nullptr;
  CUDA_TRY(cudaMalloc(&d_q, qbytes));
  cudaError_t e = cudaMalloc(&d_t, tbytes);
  if (e != cudaSuccess) { cudaFree(d_q); return set_err(BANG_E_CUDA, cudaGetErrorString(e)); }
  SearchArgs a;
  bang_b200_ctx tmp = *c;
  tmp.L = tmp.L > 0 ? tmp.L : 1;
  tmp.k = tmp.k > 0 ? tmp.k : 1;
  fill_args(&tmp, &a, d_q, Q, nullptr, nullptr);
  const size_t smem = (size_t)c->vec_units * 16 * (c->dtype == BANG_DT_FLOAT ? 1 : 4);
  e = cudaMemcpy(d_q, queries, qbytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    pick_table_kernel(c->dtype)<<<Q, 32, smem>>>(a, d_t);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(tables, d_t, tbytes, cudaMemcpyDeviceToHost);
  cudaFree(d_q);
  cudaFree(d_t);
  if (e != cudaSuccess) return set_err(BANG_E_CUDA, std::string("pq_table: ") + cudaGetErrorString(e));
  return BANG_OK;
}

// Debug builds only (-DBANG_PHASE_TIMERS): per-query, per-phase SM clock totals of the last query call.
extern "C" int bang_b200_debug_phase_clocks(bang_handle_t c, long long* out /*[Q][16]*/) {
  if (!c || !out) return set_err(BANG_E_ARG, "null argument");
  if (!c->d_phase) return set_err(BANG_E_UNSUPPORTED, "library built without BANG_PHASE_TIMERS");
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out, c->d_phase, (size_t)c->lastQ * PT_COUNT * 8, cudaMemcpyDeviceToHost));
  return BANG_OK;
}
