/* bang_b200_builder.h — GPU Vamana index builder (tooling; NOT part of the drop-in search boundary).
 *
 * The reference consumes DiskANN's `build_disk_index -R 64 -L 200` output (README.md:46-58) converted by
 * BANG_Base/bang_preprocess.py; neither DiskANN nor the datasets are available offline, so the indices of
 * the BASELINE.json shapes are built by this entry point and written in the reference's `_disk.bin` format
 * by bang_b200/builder.py.  See csrc/builder.cu.
 */
#ifndef BANG_B200_BUILDER_H_
#define BANG_B200_BUILDER_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* d_vectors: DEVICE T[N][D] (dtype as bang_dtype_t); d_order: DEVICE u32[n_order] insertion order (ids may
 * repeat for a second pass); the first n_first insertions prune with alpha_first, the rest with alpha_rest;
 * medoid: entry point of every search; max_batch 0 = min(max(N/50,1024),65536).
 * Outputs (HOST): h_deg u32[N] in [1,64]; h_nbrs u32[N][64], first h_deg[i] ids ascending, rest 0.
 * With bit 31 of max_batch set the outputs are DEVICE arrays instead (h_deg may be NULL) and unused neighbour
 * slots hold 0xFFFFFFFF — the form bang_b200_load_device_rows consumes. */
int bang_b200_build_vamana(int dtype, const void* d_vectors, uint64_t N, uint32_t D, uint32_t L_build, float alpha_first,
                           uint64_t n_first, float alpha_rest, const uint32_t* d_order, uint64_t n_order, uint64_t medoid,
                           uint32_t max_batch, uint32_t* h_deg, uint32_t* h_nbrs, float* stats_out);
const char* bang_b200_builder_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
