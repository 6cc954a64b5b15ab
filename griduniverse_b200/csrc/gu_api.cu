// gu_api.cu -- version / arch / error-string queries of the C ABI.
#include <cuda_runtime.h>
#include "gu_b200.h"

extern "C" __attribute__((visibility("default"))) int gu_version(void) { return 100; }   // 0.1.0

extern "C" __attribute__((visibility("default"))) const char* gu_arch(void) { return "sm_100a"; }

extern "C" __attribute__((visibility("default"))) const char* gu_error_string(int code) {
  switch (code) {
    case GU_OK: return "ok";
    case GU_ERR_NULL: return "required pointer is NULL";
    case GU_ERR_SHAPE: return "shape argument out of the supported range";
    case GU_ERR_ALIGN: return "pointer or pitch violates the documented alignment";
    case GU_ERR_MODE: return "unknown policy kind / flag / table format";
    case GU_ERR_UNSUPPORTED: return "configuration not supported by this kernel";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown gu_b200 error";
}
