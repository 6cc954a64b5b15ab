// gu_env.cuh -- level view, transition function and counters shared by the env kernels.
//
// Reference: core/envs/griduniverse_env.py:51-54 (edge-clamped moves), :136-155
// (look_step_ahead), :157-174 (wall / terminal predicates).
#pragma once
#include "gu_common.cuh"

namespace gu {

struct LevelsView {
  int X, Y, per_env, words;
  const uint32_t* wall;
  const uint32_t* goal;
  const uint32_t* lava;
  const int32_t* start;
  int64_t N;
};

inline LevelsView view_of(const gu_levels* lv, int64_t n) {
  LevelsView v;
  v.X = lv->X; v.Y = lv->Y; v.per_env = lv->per_env; v.words = lv->words;
  v.wall = lv->wall; v.goal = lv->goal; v.lava = lv->lava; v.start = lv->start; v.N = n;
  return v;
}

inline int check_levels(const gu_levels* lv, int64_t n) {
  if (!lv || !lv->wall || !lv->goal || !lv->lava) return GU_ERR_NULL;
  if (lv->X <= 0 || lv->Y <= 0 || n < 0) return GU_ERR_SHAPE;
  const int64_t cells = static_cast<int64_t>(lv->X) * lv->Y;
  if (cells > (1ll << 30) || lv->words != static_cast<int32_t>((cells + 31) / 32)) return GU_ERR_SHAPE;
  return GU_OK;
}

// SMEM: the planes of a shared level have been staged in shared memory by the block (plain loads;
// the read-only global path does not apply to shared addresses)
template <bool SMEM = false>
__device__ __forceinline__ bool plane_bit(const uint32_t* __restrict__ plane, const LevelsView& lv,
                                          int64_t env, int s) {
  if (SMEM) return (plane[s >> 5] >> (s & 31)) & 1u;
  const int64_t idx = lv.per_env ? static_cast<int64_t>(s >> 5) * lv.N + env : (s >> 5);
  return (__ldg(plane + idx) >> (s & 31)) & 1u;
}

// The four lambdas at griduniverse_env.py:51-54.  Only the two low bits of the action are
// used, so -1..-4 behave like the reference's negative list index (LEFT..UP).
__device__ __forceinline__ int clamp_move(int s, int a, int X, int Y) {
  const int y = s / X, x = s - y * X;
  switch (a & 3) {
    case 0: return y > 0 ? s - X : s;
    case 1: return x < X - 1 ? s + 1 : s;
    case 2: return y < Y - 1 ? s + X : s;
    default: return x > 0 ? s - 1 : s;
  }
}

// look_step_ahead (griduniverse_env.py:136-155)
template <bool SMEM = false>
__device__ __forceinline__ void transition(const LevelsView& lv, int64_t env, int s, int a, bool care,
                                           int& n, int& r, bool& term) {
  n = s;
  const bool stay = care && (plane_bit<SMEM>(lv.goal, lv, env, s) || plane_bit<SMEM>(lv.lava, lv, env, s));
  if (!stay) {
    const int c = clamp_move(s, a, lv.X, lv.Y);
    if (!plane_bit<SMEM>(lv.wall, lv, env, c)) n = c;
  }
  const bool g = plane_bit<SMEM>(lv.goal, lv, env, n), l = plane_bit<SMEM>(lv.lava, lv, env, n);
  r = reward_of(g, l);
  term = g | l;
}

__device__ __forceinline__ int start_of(const LevelsView& lv, int64_t env) {
  return __ldg(lv.start + (lv.per_env ? env : 0));
}

// Warp-reduce the per-thread counters and publish them with one atomic pair per warp.
__device__ __forceinline__ void publish_stats(long long rsum, long long dcnt, int64_t* stats) {
  if (stats == nullptr) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
    dcnt += __shfl_xor_sync(0xffffffffu, dcnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (rsum != 0) atomicAdd(reinterpret_cast<unsigned long long*>(stats), static_cast<unsigned long long>(rsum));
    if (dcnt != 0) atomicAdd(reinterpret_cast<unsigned long long*>(stats + 1), static_cast<unsigned long long>(dcnt));
  }
}

// Block-wide variant for the one-step kernels: with a single step per launch the counters are
// the only same-address traffic, and one atomic pair per warp (N/128 pairs) serialises in L2 for
// longer than the step itself takes.  Every thread of the block must call this.
__device__ __forceinline__ void publish_stats_block(long long rsum, long long dcnt, int64_t* stats) {
  if (stats == nullptr) return;
  __shared__ long long part[2][32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
    dcnt += __shfl_xor_sync(0xffffffffu, dcnt, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = (blockDim.x + 31) >> 5;
  if (lane == 0) { part[0][warp] = rsum; part[1][warp] = dcnt; }
  __syncthreads();
  if (warp == 0) {
    rsum = lane < nwarps ? part[0][lane] : 0;
    dcnt = lane < nwarps ? part[1][lane] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
      dcnt += __shfl_xor_sync(0xffffffffu, dcnt, o);
    }
    if (lane == 0) {
      if (rsum != 0) atomicAdd(reinterpret_cast<unsigned long long*>(stats), static_cast<unsigned long long>(rsum));
      if (dcnt != 0) atomicAdd(reinterpret_cast<unsigned long long*>(stats + 1), static_cast<unsigned long long>(dcnt));
    }
  }
}

}  // namespace gu
