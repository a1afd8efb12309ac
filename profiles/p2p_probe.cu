// p2p_probe: what does a random 256-byte row read from a PEER GPU's HBM cost, as a function of the footprint and of
// how the peer buffer was allocated and mapped?  Written after profiles/r1_c5.md finding 2 (the sharded search has a
// slow mode at some sizes in which every kernel phase stretches 2-6x).  The kernel mimics the traversal's access
// mix: per step one dependent peer row read (32 lanes x 8 B, like fetch_adj), plus `local_loads` random 16-byte
// loads from a local 60 MB region (the visited filters) and a few warp shuffles; clock64 buckets separate the peer
// wait from the local work, so "everything slows down" is visible directly.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o /tmp/p2p_probe profiles/p2p_probe.cu -lcuda
//   /tmp/p2p_probe <mode> <footprint MiB>[,<MiB>...] [warps_per_sm=16] [steps=2000] [local_loads=4] [fragment=0]
//   mode: local | peer (cudaMalloc + cudaDeviceEnablePeerAccess, one process) | vmm (cuMemCreate/cuMemMap, 2 MiB
//         granularity, one process) | ipc (forked owner process, cudaIpcGetMemHandle/OpenMemHandle — what the
//         search library does between ranks)
//   fragment=1: allocate and free a pile of odd-sized buffers on the owner first (what torch's caching allocator
//         leaves behind after the shard builds)
// Single process on 2 GPUs for local/peer/vmm, so `ncu` can profile it (ipc mode forks before CUDA starts).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/wait.h>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(2); } } while (0)
#define CU(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_ = nullptr; cuGetErrorString(r_, &s_); fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, s_ ? s_ : "?"); exit(2); } } while (0)

constexpr int kRowBytes = 384;

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__global__ void fill_kernel(uint32_t* p, size_t n_words) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) p[i] = mix((uint32_t)i);
}

// out[warp] = {cycles in peer waits, cycles in local work, checksum}
// LD: 0 = plain ld.global (LDG.E.64), 1 = the search kernel's form: ld.global.nc.L1::no_allocate.L2::cache_hint with an
// evict_first policy (fetch_adj in csrc/search_kernel.cuh)
template <int LD>
__global__ void __launch_bounds__(512) probe_kernel(const uint8_t* rows, uint64_t n_rows, const uint4* local, uint32_t n_local16,
                                                    int steps, int local_loads, unsigned long long* out) {
  const uint32_t lane = threadIdx.x & 31, warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  uint32_t state = mix(warp * 2654435761u + 12345u);
  long long t_peer = 0, t_local = 0;
  uint32_t acc = 0;
  uint64_t pol = 0;
  if (LD == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  for (int s = 0; s < steps; ++s) {
    const uint64_t row = (((uint64_t)mix(state) << 32) | mix(state ^ 0x9e3779b9u)) % n_rows;
    long long t0 = clock64();
    uint2 v;
    const uint8_t* p = rows + row * kRowBytes + lane * 8;
    if (LD == 1) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    else v = *reinterpret_cast<const uint2*>(p);
    uint32_t x = v.x ^ v.y;
    x = __reduce_xor_sync(0xffffffffu, x);   // the next row depends on the data: one request in flight per warp
    long long t1 = clock64();
    t_peer += t1 - t0;
    for (int j = 0; j < local_loads; ++j) {
      const uint4 w = local[mix(x + j * 0x85ebca6bu + lane * 0xc2b2ae35u) % n_local16];
      acc ^= w.x ^ w.y ^ w.z ^ w.w;
    }
    acc = __reduce_xor_sync(0xffffffffu, acc);
#pragma unroll 1
    for (int j = 0; j < 16; ++j) acc = mix(acc) ^ __shfl_xor_sync(0xffffffffu, acc, 1 << (j & 3));
    t_local += clock64() - t1;
    state = x ^ acc ^ (uint32_t)s;
  }
  if (lane == 0) { out[warp * 3] = t_peer; out[warp * 3 + 1] = t_local; out[warp * 3 + 2] = acc; }
}

static void fragment_heap(size_t total_free) {
  // many odd-sized allocations, every other one freed again, the rest freed at the end: leaves a chopped-up heap
  std::vector<void*> keep;
  size_t sz = 37u << 20, used = 0;
  while (used + sz < total_free / 2) {
    void* p = nullptr;
    if (cudaMalloc(&p, sz) != cudaSuccess) { cudaGetLastError(); break; }
    keep.push_back(p);
    used += sz;
    sz = (sz * 1103515245u + 12345u) % (900u << 20) + (3u << 20);
  }
  for (size_t i = 0; i < keep.size(); i += 2) cudaFree(keep[i]);
  for (size_t i = 1; i < keep.size(); i += 2) cudaFree(keep[i]);
}

struct Run { const uint8_t* rows; uint64_t n_rows; };

static int g_ldmode = 0;  // P2P_PROBE_LD=1: the search kernel's load form

static void measure(const char* mode, size_t mib, Run r, int warps_per_sm, int steps, int local_loads) {
  CK(cudaSetDevice(0));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int wpc = warps_per_sm > 16 ? 16 : warps_per_sm, ctas = sms * ((warps_per_sm + wpc - 1) / wpc);
  const uint32_t n_local16 = (60u << 20) / 16;
  uint4* local = nullptr;
  CK(cudaMalloc(&local, (size_t)n_local16 * 16));
  fill_kernel<<<1024, 256>>>(reinterpret_cast<uint32_t*>(local), (size_t)n_local16 * 4);
  unsigned long long* out = nullptr;
  const int n_warps = ctas * wpc;
  CK(cudaMalloc(&out, (size_t)n_warps * 3 * 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    if (g_ldmode) probe_kernel<1><<<ctas, wpc * 32>>>(r.rows, r.n_rows, local, n_local16, steps, local_loads, out);
    else probe_kernel<0><<<ctas, wpc * 32>>>(r.rows, r.n_rows, local, n_local16, steps, local_loads, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  std::vector<unsigned long long> h((size_t)n_warps * 3);
  CK(cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost));
  double tp = 0, tl = 0;
  for (int w = 0; w < n_warps; ++w) { tp += (double)h[w * 3]; tl += (double)h[w * 3 + 1]; }
  printf("{\"mode\": \"%s\", \"ld\": %d, \"footprint_mib\": %zu, \"warps_per_sm\": %d, \"steps\": %d, \"local_loads\": %d, \"ms\": %.3f, "
         "\"rows_per_s\": %.3e, \"peer_wait_cycles\": %.0f, \"local_work_cycles\": %.0f}\n",
         mode, g_ldmode, mib, warps_per_sm, steps, local_loads, best, (double)n_warps * steps / (best * 1e-3), tp / n_warps / steps, tl / n_warps / steps);
  fflush(stdout);
  cudaFree(local); cudaFree(out);
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s local|peer|vmm|ipc MiB[,MiB...] [warps_per_sm] [steps] [local_loads] [fragment]\n", argv[0]); return 1; }
  const std::string mode = argv[1];
  std::vector<size_t> sizes;
  for (char* tok = strtok(argv[2], ","); tok; tok = strtok(nullptr, ",")) sizes.push_back((size_t)atoll(tok));
  const int wps = argc > 3 ? atoi(argv[3]) : 16, steps = argc > 4 ? atoi(argv[4]) : 2000, ll = argc > 5 ? atoi(argv[5]) : 4;
  const bool frag = argc > 6 && atoi(argv[6]) != 0;
  if (const char* e = getenv("P2P_PROBE_LD")) g_ldmode = atoi(e) != 0;

  for (size_t mib : sizes) {
    const size_t bytes = mib << 20;
    const uint64_t n_rows = bytes / kRowBytes;
    if (mode == "ipc") {
      int to_parent[2], to_child[2];
      if (pipe(to_parent) || pipe(to_child)) return 3;
      const pid_t pid = fork();   // before any CUDA call in this process
      if (pid == 0) {
        CK(cudaSetDevice(1));
        if (frag) { size_t f, t; CK(cudaMemGetInfo(&f, &t)); fragment_heap(f); }
        void* p = nullptr;
        CK(cudaMalloc(&p, bytes));
        fill_kernel<<<2048, 256>>>(reinterpret_cast<uint32_t*>(p), bytes / 4);
        CK(cudaDeviceSynchronize());
        cudaIpcMemHandle_t hnd;
        CK(cudaIpcGetMemHandle(&hnd, p));
        if (write(to_parent[1], &hnd, sizeof(hnd)) != (ssize_t)sizeof(hnd)) _exit(4);
        char c;
        if (read(to_child[0], &c, 1) != 1) _exit(5);   // wait until the reader is done
        cudaFree(p);
        _exit(0);
      }
      cudaIpcMemHandle_t hnd;
      if (read(to_parent[0], &hnd, sizeof(hnd)) != (ssize_t)sizeof(hnd)) return 6;
      CK(cudaSetDevice(0));
      void* p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
      measure("ipc", mib, Run{static_cast<const uint8_t*>(p), n_rows}, wps, steps, ll);
      CK(cudaIpcCloseMemHandle(p));
      if (write(to_child[1], "x", 1) != 1) return 7;
      int st = 0;
      waitpid(pid, &st, 0);
      // a second CUDA-using fork from a process that already initialised CUDA is not allowed: one size per ipc run
      if (sizes.size() > 1) { fprintf(stderr, "ipc mode: one footprint per invocation\n"); break; }
    } else if (mode == "vmm") {
      CK(cudaSetDevice(1)); CK(cudaFree(0));
      CK(cudaSetDevice(0)); CK(cudaFree(0));
      CUmemAllocationProp prop;
      memset(&prop, 0, sizeof(prop));
      prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
      prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      prop.location.id = 1;
      size_t gran = 0;
      CU(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
      const size_t padded = (bytes + gran - 1) / gran * gran;
      CUmemGenericAllocationHandle h;
      CU(cuMemCreate(&h, padded, &prop, 0));
      CUdeviceptr va = 0;
      CU(cuMemAddressReserve(&va, padded, gran, 0, 0));
      CU(cuMemMap(va, padded, 0, h, 0));
      CUmemAccessDesc acc[2];
      for (int d = 0; d < 2; ++d) { acc[d].location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc[d].location.id = d; acc[d].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE; }
      CU(cuMemSetAccess(va, padded, acc, 2));
      CK(cudaSetDevice(1));
      fill_kernel<<<2048, 256>>>(reinterpret_cast<uint32_t*>(va), bytes / 4);
      CK(cudaDeviceSynchronize());
      fprintf(stderr, "vmm: granularity %zu\n", gran);
      measure("vmm", mib, Run{reinterpret_cast<const uint8_t*>(va), n_rows}, wps, steps, ll);
      CU(cuMemUnmap(va, padded)); CU(cuMemAddressFree(va, padded)); CU(cuMemRelease(h));
    } else {
      const int owner = mode == "local" ? 0 : 1;
      CK(cudaSetDevice(owner));
      if (frag) { size_t f, t; CK(cudaMemGetInfo(&f, &t)); fragment_heap(f); }
      void* p = nullptr;
      CK(cudaMalloc(&p, bytes));
      fill_kernel<<<2048, 256>>>(reinterpret_cast<uint32_t*>(p), bytes / 4);
      CK(cudaDeviceSynchronize());
      if (owner == 1) {
        CK(cudaSetDevice(0));
        cudaError_t e = cudaDeviceEnablePeerAccess(1, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        cudaGetLastError();
      }
      measure(mode.c_str(), mib, Run{static_cast<const uint8_t*>(p), n_rows}, wps, steps, ll);
      CK(cudaSetDevice(owner));
      CK(cudaFree(p));
    }
  }
  return 0;
}
