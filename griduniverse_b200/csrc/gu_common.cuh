// gu_common.cuh -- shared device helpers for the GridUniverse sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <math_constants.h>

#include "gu_b200.h"

#define GU_CHECK_LAUNCH()                              \
  do {                                                 \
    cudaError_t e__ = cudaGetLastError();              \
    if (e__ != cudaSuccess) return (int)e__;           \
  } while (0)

namespace gu {

constexpr int kRewardStep = -1;   // griduniverse_env.py:80
constexpr int kRewardGoal = 10;    // :83
constexpr int kRewardLava = -10;   // :88 (written last, wins)

__host__ __device__ __forceinline__ int reward_of(bool goal, bool lava) {
  return lava ? kRewardLava : (goal ? kRewardGoal : kRewardStep);
}

// ---- exact (non-contracted) arithmetic in the reference's evaluation order ----
template <typename T> struct Num;
template <> struct Num<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double rnd(double a) { return rint(a); }
  static __device__ __forceinline__ double abs(double a) { return fabs(a); }
  static __device__ __forceinline__ double scale() { return 1e8; }
  static __device__ __forceinline__ double neg_inf() { return -CUDART_INF; }
  static __device__ __forceinline__ double inv(int n) {
    return n == 1 ? 1.0 : n == 2 ? 0.5 : n == 3 ? (1.0 / 3.0) : 0.25;
  }
  // rint() is only needed when |t| < 2^52; below 2^51 the magic-number add is exact.
  static __device__ __forceinline__ double magic() { return 6755399441055744.0; }   // 1.5 * 2^52
  static __device__ __forceinline__ double magic_limit() { return 2251799813685248.0; }  // 2^51
};
template <> struct Num<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float rnd(float a) { return rintf(a); }
  static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
  static __device__ __forceinline__ float scale() { return 1e8f; }
  static __device__ __forceinline__ float neg_inf() { return -CUDART_INF_F; }
  static __device__ __forceinline__ float inv(int n) {
    return n == 1 ? 1.0f : n == 2 ? 0.5f : n == 3 ? (1.0f / 3.0f) : 0.25f;
  }
  static __device__ __forceinline__ float magic() { return 12582912.0f; }        // 1.5 * 2^23
  static __device__ __forceinline__ float magic_limit() { return 4194304.0f; }   // 2^22
};

// ---- signed max reductions ----------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

// atomic max on a float / double scalar initialised to -inf (ordered-bits trick).
__device__ __forceinline__ void atomic_max_signed(float* addr, float v) {
  if (v >= 0.0f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_signed(double* addr, double v) {
  if (v >= 0.0)
    atomicMax(reinterpret_cast<long long*>(addr), __double_as_longlong(v));
  else
    atomicMin(reinterpret_cast<unsigned long long*>(addr),
              static_cast<unsigned long long>(__double_as_longlong(v)));
}

// Block-wide signed max of `v`, result combined into *out by one atomic per block.
// `scratch` holds one T per warp.
template <typename T>
__device__ __forceinline__ void block_max_to_global(T v, T* scratch, T* out) {
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nthreads = blockDim.x * blockDim.y;
  v = warp_max(v);
  if ((tid & 31) == 0) scratch[tid >> 5] = v;
  __syncthreads();
  if (tid < 32) {
    const int nwarps = (nthreads + 31) >> 5;
    T w = tid < nwarps ? scratch[tid] : Num<T>::neg_inf();
    w = warp_max(w);
    if (tid == 0 && out != nullptr) atomic_max_signed(out, w);
  }
}

// ---- mbarrier / TMA bulk-copy helpers (cp.async.bulk -> SASS UBLKCP, SYNCS) -------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "GU_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra GU_DONE;\n\t"
      "bra GU_WAIT;\n\t"
      "GU_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

}  // namespace gu
