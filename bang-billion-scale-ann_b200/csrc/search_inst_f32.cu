// The float instantiations of the fused search kernel and of the stand-alone PQ table kernel (see search_inst.cuh).
#define BANG_INST_T float
#include "search_inst.cuh"

namespace bang {
search_fn_t search_kernel_f32(int mode, uint32_t cs, int wpc, bool ph) { return inst_lookup<float>(mode, cs, wpc, ph); }
table_fn_t table_kernel_f32() { return pq_table_kernel<float>; }
}  // namespace bang
