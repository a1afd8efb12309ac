// search_inst.cuh — the fused search kernel is instantiated once per element type, each in its own translation unit
// (search_inst_u8.cu / _i8.cu / _f32.cu: 25 kernels apiece, compiled in parallel by build.py); the host code picks
// an instantiation through these lookups.
#pragma once
#include "search_kernel.cuh"

namespace bang {
typedef void (*search_fn_t)(const SearchArgs);
typedef void (*table_fn_t)(const SearchArgs, float*);

// mode: 0 Base / 1 Inmemory / 2 Exactdistance; cs: 4 (32 uniform chunks, padded to 4 dimensions) or 0 (general); wpc: 16, 24 or 32 (wpc_variant);
// ph: rows carry the precomputed visited-filter slots of their neighbours (PQ modes only, search_kernel.cuh kSlotBytes)
search_fn_t search_kernel_u8(int mode, uint32_t cs, int wpc, bool ph);
search_fn_t search_kernel_i8(int mode, uint32_t cs, int wpc, bool ph);
search_fn_t search_kernel_f32(int mode, uint32_t cs, int wpc, bool ph);
table_fn_t table_kernel_u8();
table_fn_t table_kernel_i8();
table_fn_t table_kernel_f32();

#ifdef BANG_INST_T
template <typename T, int CS, int WPC, bool PH>
static search_fn_t inst_mode(int mode) {
  return mode == kBase ? bang_search_kernel<T, kBase, CS, WPC, PH> : bang_search_kernel<T, kInmemory, CS, WPC, PH>;
}
template <typename T, int CS, bool PH>
static search_fn_t inst_wpc(int mode, int wpc) {
  return wpc <= 16 ? inst_mode<T, CS, 16, PH>(mode) : (wpc <= 24 ? inst_mode<T, CS, 24, PH>(mode) : inst_mode<T, CS, 32, PH>(mode));
}
template <typename T, bool PH>
static search_fn_t inst_cs(int mode, uint32_t cs, int wpc) {
  return cs == 4 ? inst_wpc<T, 4, PH>(mode, wpc) : inst_wpc<T, 0, PH>(mode, wpc);
}
template <typename T>
static search_fn_t inst_lookup(int mode, uint32_t cs, int wpc, bool ph) {
  if (mode == kExact) return bang_search_kernel<T, kExact, 0, 16, false>;
  return ph ? inst_cs<T, true>(mode, cs, wpc) : inst_cs<T, false>(mode, cs, wpc);
}
#endif
}  // namespace bang
