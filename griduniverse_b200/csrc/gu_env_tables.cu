// gu_env_tables.cu -- table-driven rollout fast paths (transition tables staged in shared
// memory).  Filled in after the layout-agnostic kernels are parity-green; until then no
// shape has a table format and gu_rollout uses rollout_generic_kernel.
#include "gu_common.cuh"

namespace gu {

int rollout_tables(const gu_levels*, int64_t, int64_t, const int32_t*, int32_t*, int32_t*, int32_t*,
                   uint8_t*, const int32_t*, int32_t*, int32_t*, int64_t*, const uint32_t*, uint32_t,
                   cudaStream_t) {
  return GU_ERR_UNSUPPORTED;
}

}  // namespace gu

extern "C" __attribute__((visibility("default"))) int64_t gu_tables_bytes(const gu_levels*, int64_t) { return 0; }

extern "C" __attribute__((visibility("default"))) int gu_pack_tables(const gu_levels*, int64_t, uint32_t*, uint32_t, void*) {
  return GU_ERR_UNSUPPORTED;
}
