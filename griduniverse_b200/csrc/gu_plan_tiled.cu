// gu_plan_tiled.cu -- tiled fast paths for large aligned grids (filled in after the generic
// kernels are parity-green).  Until then every entry reports GU_ERR_UNSUPPORTED and the
// callers in gu_plan.cu use the generic kernels.
#include "gu_cell.cuh"

namespace gu {

int sweep_tiled_f32(const gu_grid*, const float*, float*, int, const void*, float, float*, const float*,
                    float, cudaStream_t) {
  return GU_ERR_UNSUPPORTED;
}
int sweep_tiled_f64(const gu_grid*, const double*, double*, int, const void*, double, double*, const double*,
                    double, cudaStream_t) {
  return GU_ERR_UNSUPPORTED;
}
int greedy_tiled_f32(const gu_grid*, const float*, uint8_t*, float, cudaStream_t) { return GU_ERR_UNSUPPORTED; }
int greedy_tiled_f64(const gu_grid*, const double*, uint8_t*, double, cudaStream_t) { return GU_ERR_UNSUPPORTED; }

}  // namespace gu
