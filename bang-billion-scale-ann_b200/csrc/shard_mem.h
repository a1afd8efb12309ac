// shard_mem — device memory for the graph rows that other processes (one per GPU) map and read over NVLink.
//
// Two schemes:
//   legacy  cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle (64-byte handles, any transport)
//   vmm     cuMemCreate (pinned device memory, 2 MiB-or-larger granularity, shareable as a POSIX file descriptor)
//           + cuMemAddressReserve / cuMemMap / cuMemSetAccess on both sides; the descriptor travels between the
//           processes over a Unix-domain socket (SCM_RIGHTS; bang_b200/sharding.py).  The owner controls the
//           physical granularity and both sides the alignment of the mapping, which the legacy scheme leaves to the
//           driver.  Measured (profiles/r2_c5.md): the traversal over peer rows mapped this way runs within 4 % of
//           the single-GPU time, while the legacy mappings have a slow mode at some sizes (1.7x at 9 M points on 2 GPUs).
// Sharded indices use vmm (BANG_B200_SHARD_VMM=0 selects legacy); unsharded ones plain cudaMalloc.  The driver API is reached
// through cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace bang {

struct ShardMem {
  void* ptr = nullptr;      // device pointer (owner) or peer mapping (importer)
  size_t bytes = 0;         // usable bytes
  size_t mapped = 0;        // vmm: bytes reserved and mapped (multiple of the granularity)
  unsigned long long handle = 0;  // vmm: CUmemGenericAllocationHandle
  bool vmm = false;
  bool imported = false;    // legacy: opened with cudaIpcOpenMemHandle
};

bool shard_vmm_requested();  // true unless BANG_B200_SHARD_VMM=0

// All return 0 on success; otherwise a negative value and *err describes the failure.
int shard_alloc(ShardMem* m, size_t bytes, int device, bool vmm, std::string* err);
void shard_release(ShardMem* m);  // owner's allocation or an imported mapping
int shard_export_fd(const ShardMem* m, int* fd_out, std::string* err);                       // vmm only; caller closes fd
int shard_import_fd(ShardMem* m, int fd, size_t bytes, int device, std::string* err);       // vmm only; fd stays open

}  // namespace bang
