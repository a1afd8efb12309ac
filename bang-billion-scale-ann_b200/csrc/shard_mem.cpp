// shard_mem.cpp — see shard_mem.h.
#include "shard_mem.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace bang {

namespace {

struct DriverApi {
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  bool ok = false;
  std::string why;
};

template <typename F>
bool resolve(const char* name, F* fn, std::string* why) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult st;
  const cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st);
  if (e != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    *why = std::string("driver entry point ") + name + " not available";
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}

const DriverApi& driver() {
  static const DriverApi api = [] {
    DriverApi a;
    a.ok = resolve("cuGetErrorString", &a.GetErrorString, &a.why) &&
           resolve("cuMemGetAllocationGranularity", &a.MemGetAllocationGranularity, &a.why) &&
           resolve("cuMemCreate", &a.MemCreate, &a.why) && resolve("cuMemRelease", &a.MemRelease, &a.why) &&
           resolve("cuMemAddressReserve", &a.MemAddressReserve, &a.why) && resolve("cuMemAddressFree", &a.MemAddressFree, &a.why) &&
           resolve("cuMemMap", &a.MemMap, &a.why) && resolve("cuMemUnmap", &a.MemUnmap, &a.why) &&
           resolve("cuMemSetAccess", &a.MemSetAccess, &a.why) &&
           resolve("cuMemExportToShareableHandle", &a.MemExportToShareableHandle, &a.why) &&
           resolve("cuMemImportFromShareableHandle", &a.MemImportFromShareableHandle, &a.why);
    return a;
  }();
  return api;
}

int cu_fail(const char* what, CUresult r, std::string* err) {
  const char* s = nullptr;
  if (driver().GetErrorString) driver().GetErrorString(r, &s);
  *err = std::string(what) + ": " + (s ? s : "CUDA driver error");
  return -1;
}

CUmemAllocationProp device_prop(int device) {
  CUmemAllocationProp p;
  memset(&p, 0, sizeof(p));
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

// reserve + map + grant `device` access to an allocation handle
int map_handle(ShardMem* m, CUmemGenericAllocationHandle h, size_t padded, size_t gran, int device, bool writable, std::string* err) {
  const DriverApi& d = driver();
  CUdeviceptr va = 0;
  CUresult r = d.MemAddressReserve(&va, padded, gran, 0, 0);
  if (r != CUDA_SUCCESS) return cu_fail("cuMemAddressReserve", r, err);
  r = d.MemMap(va, padded, 0, h, 0);
  if (r != CUDA_SUCCESS) { d.MemAddressFree(va, padded); return cu_fail("cuMemMap", r, err); }
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = writable ? CU_MEM_ACCESS_FLAGS_PROT_READWRITE : CU_MEM_ACCESS_FLAGS_PROT_READ;
  r = d.MemSetAccess(va, padded, &acc, 1);
  if (r != CUDA_SUCCESS) { d.MemUnmap(va, padded); d.MemAddressFree(va, padded); return cu_fail("cuMemSetAccess", r, err); }
  m->ptr = reinterpret_cast<void*>(va);
  m->mapped = padded;
  m->handle = h;
  m->vmm = true;
  return 0;
}

}  // namespace

bool shard_vmm_requested() {  // default for sharded indices; BANG_B200_SHARD_VMM=0 selects the CUDA-IPC scheme
  const char* e = getenv("BANG_B200_SHARD_VMM");
  return !(e && strcmp(e, "0") == 0);
}

int shard_alloc(ShardMem* m, size_t bytes, int device, bool vmm, std::string* err) {
  *m = ShardMem();
  if (bytes < 256) bytes = 256;
  if (!vmm) {
    const cudaError_t e = cudaMalloc(&m->ptr, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); *err = cudaGetErrorString(e); return e == cudaErrorMemoryAllocation ? -2 : -1; }
    m->bytes = bytes;
    return 0;
  }
  const DriverApi& d = driver();
  if (!d.ok) { *err = d.why; return -1; }
  const CUmemAllocationProp prop = device_prop(device);
  size_t gran = 0;
  CUresult r = d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
  if (r != CUDA_SUCCESS || gran == 0) return cu_fail("cuMemGetAllocationGranularity", r, err);
  const size_t padded = (bytes + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h = 0;
  r = d.MemCreate(&h, padded, &prop, 0);
  if (r != CUDA_SUCCESS) { cu_fail("cuMemCreate", r, err); return r == CUDA_ERROR_OUT_OF_MEMORY ? -2 : -1; }
  if (map_handle(m, h, padded, gran, device, true, err) != 0) { d.MemRelease(h); return -1; }
  m->bytes = bytes;
  return 0;
}

void shard_release(ShardMem* m) {
  if (!m->ptr) return;
  if (m->vmm) {
    const DriverApi& d = driver();
    const CUdeviceptr va = reinterpret_cast<CUdeviceptr>(m->ptr);
    d.MemUnmap(va, m->mapped);
    d.MemAddressFree(va, m->mapped);
    d.MemRelease(m->handle);
  } else if (m->imported) {
    cudaIpcCloseMemHandle(m->ptr);
  } else {
    cudaFree(m->ptr);
  }
  *m = ShardMem();
}

int shard_export_fd(const ShardMem* m, int* fd_out, std::string* err) {
  if (!m->vmm || !m->ptr) { *err = "the shard was not allocated with BANG_B200_SHARD_VMM=1"; return -1; }
  int fd = -1;
  const CUresult r = driver().MemExportToShareableHandle(&fd, m->handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) return cu_fail("cuMemExportToShareableHandle", r, err);
  *fd_out = fd;
  return 0;
}

int shard_import_fd(ShardMem* m, int fd, size_t bytes, int device, std::string* err) {
  *m = ShardMem();
  const DriverApi& d = driver();
  if (!d.ok) { *err = d.why; return -1; }
  // the granularity of the importing device bounds the alignment of the mapping; the owner padded to its own
  // (identical on one box)
  const CUmemAllocationProp prop = device_prop(device);
  size_t gran = 0;
  CUresult r = d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
  if (r != CUDA_SUCCESS || gran == 0) return cu_fail("cuMemGetAllocationGranularity", r, err);
  const size_t padded = (std::max<size_t>(bytes, 256) + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h = 0;
  r = d.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  if (r != CUDA_SUCCESS) return cu_fail("cuMemImportFromShareableHandle", r, err);
  if (map_handle(m, h, padded, gran, device, false, err) != 0) { d.MemRelease(h); return -1; }
  m->bytes = bytes;
  m->imported = true;
  return 0;
}

}  // namespace bang
