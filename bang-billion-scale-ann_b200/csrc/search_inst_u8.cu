// The uint8_t instantiations of the fused search kernel and of the stand-alone PQ table kernel (see search_inst.cuh).
#define BANG_INST_T uint8_t
#include "search_inst.cuh"

namespace bang {
search_fn_t search_kernel_u8(int mode, uint32_t cs, int wpc, bool ph) { return inst_lookup<uint8_t>(mode, cs, wpc, ph); }
table_fn_t table_kernel_u8() { return pq_table_kernel<uint8_t>; }
}  // namespace bang
