// Lookahead attention, tcgen05 + TMA path (impl=2).  Placeholder until the kernel lands: fails loudly.
#include "common.cuh"
namespace lade {
int attn_fwd_tc_launch(cudaStream_t, const void*, const void*, const void*, void*, const int32_t*, const int32_t*,
                       void*, int, int, int, int, int, int, int) {
  return LADE_EUNSUPPORTED;
}
}  // namespace lade
