// bang_b200_filter_slots: the host side of the visited-filter slot arithmetic (hash1 / hash2 / vis_slot_word of
// search_kernel.cuh, which the load kernels and the search kernels use on the device), exported so that it can be
// checked against the oracle's restatement of hashFn1_d / hashFn2_d (bang_search.cu:1168-1189) without a GPU.
#include "bang_b200.h"
#include "search_kernel.cuh"

extern "C" void bang_b200_filter_slots(uint32_t id, uint32_t pos[2], uint32_t words[2]) {
  pos[0] = bang::hash1(id);
  pos[1] = bang::hash2(id);
  words[0] = bang::vis_slot_word(pos[0]);
  words[1] = bang::vis_slot_word(pos[1]);
}
