// vqb200 K1 (tensor-core variant) -- placeholder until the tcgen05 kernel lands.
#include "common.cuh"
namespace vqb200 {
bool assign_tc_eligible(const ZView&, int, int) { return false; }
size_t assign_tc_workspace_bytes(long long N) { return 256 + (size_t)(N > 0 ? N : 0) * sizeof(int32_t); }
int launch_assign_tc(const ZView&, const float*, const float*, const void*, const float*, int, int, int32_t*, float*,
                     void*, size_t, cudaStream_t) {
  return fail(VQB200_EUNSUPPORTED, "vq_assign(TC): not built");
}
}  // namespace vqb200
