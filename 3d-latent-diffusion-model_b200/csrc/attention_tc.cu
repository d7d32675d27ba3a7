// attention_tc.cu - tcgen05 flash attention (placeholder until the kernel lands; AUTO routes to SIMT).
#include "common.cuh"
namespace b200ldm {
bool attention_tcgen05_supported(int d) { (void)d; return false; }
int attention_tcgen05(const void*, const void*, const void*, void*, int, int, int, int, long long, float,
                      cudaStream_t) {
  return fail(B200LDM_BAD_ARG, "tcgen05 attention not built");
}
}  // namespace b200ldm
