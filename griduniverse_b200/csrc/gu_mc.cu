// gu_mc.cu -- Monte-Carlo evaluation: per-episode truncated discounted returns, first /
// every-visit accumulation and the value update (core/algorithms/monte_carlo.py:53-97).
//
// The reference sums  G(idx) = sum_i gamma^i * r[idx+i]  over the i with gamma^i > threshold,
// left to right, for every counted visit index, then folds the per-state sums into V
// sequentially over episodes.  To stay bit-exact the device keeps exactly that order:
// one thread per visit index for G (sequential in i), one thread per state for the
// accumulation over visit indices (sequential in idx).  fp64, no fused multiply-add.
#include "gu_env.cuh"

namespace gu {

__device__ __forceinline__ int mc_state(int start, const int32_t* __restrict__ obs, int64_t stride, int idx) {
  return idx == 0 ? start : __ldg(obs + static_cast<int64_t>(idx - 1) * stride);
}

// G[idx], idx in [0, L]  (monte_carlo.py:69-70; an empty tail gives 0)
__global__ void __launch_bounds__(128)
mc_returns_kernel(int L, const int32_t* __restrict__ rewards, int64_t stride,
                  const double* __restrict__ weights, const uint8_t* __restrict__ keep,
                  double* __restrict__ G) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx > L) return;
  double acc = 0.0;
  const int n = L - idx;
  for (int i = 0; i < n; ++i) {
    if (__ldg(keep + i)) {
      const double r = static_cast<double>(__ldg(rewards + static_cast<int64_t>(idx + i) * stride));
      acc = __dadd_rn(acc, __dmul_rn(__ldg(weights + i), r));
    }
  }
  G[idx] = acc;
}

// per state: visit counting (:56-68), return accumulation (:71) and the update (:74-91)
__global__ void __launch_bounds__(128)
mc_update_kernel(int cells, int L, const int32_t* __restrict__ start_p, const int32_t* __restrict__ obs,
                 int64_t stride, const double* __restrict__ G, int every_visit, int mode, double alpha,
                 double* __restrict__ total_visits, double* __restrict__ total_return,
                 double* __restrict__ value) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cells) return;
  const int start = __ldg(start_p);
  double visits = 0.0, ret = 0.0;
  for (int idx = 0; idx <= L; ++idx) {
    if (mc_state(start, obs, stride, idx) != s) continue;
    if (visits != 0.0 && !every_visit) continue;
    visits = __dadd_rn(visits, 1.0);
    ret = __dadd_rn(ret, G[idx]);
  }
  const double tv = __dadd_rn(total_visits[s], visits);
  total_visits[s] = tv;
  if (mode == 2) {                       // not incremental_mean: S(s) += G (:78-80)
    total_return[s] = __dadd_rn(total_return[s], ret);
  } else if (mode == 0) {                // V += (1/N) * (G - V) for every state with N > 0 (:83-87)
    if (tv > 0.0) {
      const double v = value[s];
      value[s] = __dadd_rn(v, __dmul_rn(__ddiv_rn(1.0, tv), __dadd_rn(ret, -v)));
    }
  } else {                               // non-stationary: V += alpha * (G - V) (:88-91)
    const double v = value[s];
    value[s] = __dadd_rn(v, __dmul_rn(alpha, __dadd_rn(ret, -v)));
  }
}

__global__ void __launch_bounds__(128)
mc_finalize_kernel(int cells, const double* __restrict__ total_visits,
                   const double* __restrict__ total_return, double* __restrict__ value) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cells) return;
  if (total_visits[s] > 0.0) value[s] = __ddiv_rn(total_return[s], total_visits[s]);   // :93-97
}

// ---- a whole batch of episodes in one launch ------------------------------------------------------
// monte_carlo_evaluation's loop (monte_carlo.py:49-91) for E episodes: the episodes are sequential by
// nature -- episode e consumes the uniform draws that follow episode e-1's, and V is folded episode
// by episode -- so one thread block walks them in order: thread 0 plays the episode
// (run_episode, :7-26), then all threads compute the truncated returns (one visit index each) and
// the per-state accumulation / update, in exactly the order of mc_returns_kernel / mc_update_kernel.
// No host round trip per episode; the host reads the lengths, done flags and the number of draws
// consumed once per launch.
constexpr int kMcThreads = 1024;      // the return and per-state passes are one visit index / one state per thread
constexpr int kMcMaxStagedT = 4096;   // episodes up to this cap keep masked weights + rewards in shared memory

__global__ void __launch_bounds__(kMcThreads)
mc_evaluate_kernel(LevelsView lv, const double* __restrict__ cdf, const double* __restrict__ uniforms,
                   long long n_uniforms, const int32_t* __restrict__ starts, int E, int T,
                   const double* __restrict__ weights, const uint8_t* __restrict__ keep, int every_visit, int mode,
                   double alpha, int cells, int32_t* __restrict__ obs, int32_t* __restrict__ rew,
                   double* __restrict__ G, double* __restrict__ total_visits, double* __restrict__ total_return,
                   double* __restrict__ value, int32_t* __restrict__ lengths, uint8_t* __restrict__ done,
                   long long* __restrict__ meta, int use_smem, int use_wr) {
  __shared__ int L_sh, stop_sh;
  __shared__ long long off_sh;
  // Level small enough (dynamic shared memory was granted): the CDF rows and a (cell, action) ->
  // landing cell / reward / done table are staged once, so a step of the sequential walk costs two
  // shared-memory round trips instead of eight dependent global loads.
  extern __shared__ __align__(16) double mc_smem[];
  const bool staged = use_smem != 0;
  // use_wr: the discount weights, zeroed where the reference drops the term (a zero weight adds +-0 to the
  // running sum, which leaves it bit-identical), and the episode's rewards as doubles live in shared
  // memory, so the truncated-return pass is a branch-free multiply-add loop over shared operands
  double* wk_s = mc_smem;                                              // [T]
  double* rs_s = mc_smem + (use_wr ? T : 0);                           // [T]
  double* cdf_s = mc_smem + (use_wr ? 2 * T : 0);                      // [cells][4]
  uint16_t* nt_s = reinterpret_cast<uint16_t*>(cdf_s + 4 * cells);     // [cells][4]: landing | goal << 14 | lava << 15
  const int tid = threadIdx.x;
  if (use_wr)
    for (int i = tid; i < T; i += kMcThreads) wk_s[i] = __ldg(keep + i) ? __ldg(weights + i) : 0.0;
  if (staged) {
    for (int i = tid; i < cells * 4; i += kMcThreads) {
      cdf_s[i] = __ldg(cdf + i);
      int n, r;
      bool d;
      transition(lv, 0, i >> 2, i & 3, true, n, r, d);
      nt_s[i] = static_cast<uint16_t>(n | (r == kRewardGoal ? 0x4000 : 0) | (r == kRewardLava ? 0x8000 : 0));
    }
  }
  if (tid == 0) { off_sh = 0; stop_sh = 0; }
  __syncthreads();
  int e = 0;
  for (; e < E; ++e) {
    const int start = __ldg(starts + e);
    if (tid < 32) {
      // Warp 0 walks the episode in lockstep -- every lane holds the same state, lane 0 writes -- so that
      // the uniform draws can be fetched 32 at a time, one per lane and one window ahead, and handed to the
      // step by warp shuffle: no global-memory round trip on the sequential chain state -> action -> state.
      int s = start, t = 0, stop = 0;
      bool d = false;
      const long long off = off_sh;
      double ucur = off + tid < n_uniforms ? __ldg(uniforms + off + tid) : 0.0;
      while (t < T && !d && stop == 0) {
        const long long nx = off + t + 32 + tid;
        const double unext = nx < n_uniforms ? __ldg(uniforms + nx) : 0.0;
        const int wend = min(t + 32, T);
        for (; t < wend && !d; ++t) {
          if (off + t >= n_uniforms) { stop = 1; break; }              // out of draws: the host continues
          const double u = __shfl_sync(0xffffffffu, ucur, t & 31);
          int n, r;
          if (staged) {
            const double2 c01 = *reinterpret_cast<const double2*>(cdf_s + s * 4);
            const double2 c23 = *reinterpret_cast<const double2*>(cdf_s + s * 4 + 2);
            const uint2 ntr = *reinterpret_cast<const uint2*>(nt_s + s * 4);   // the four landing entries of s
            if (c23.y != c23.y) { stop = 2 + s; break; }               // np.random.choice would raise here
            const int a = (c01.x <= u) + (c01.y <= u) + (c23.x <= u);
            const uint32_t e = (((a & 2) ? ntr.y : ntr.x) >> ((a & 1) * 16)) & 0xffffu;
            n = static_cast<int>(e & 0x3fffu);
            r = (e & 0x8000u) ? kRewardLava : ((e & 0x4000u) ? kRewardGoal : kRewardStep);
            d = (e & 0xc000u) != 0;
          } else {
            const double* row = cdf + static_cast<int64_t>(s) * 4;
            if (__ldg(row + 3) != __ldg(row + 3)) { stop = 2 + s; break; }
            const int a = (__ldg(row) <= u) + (__ldg(row + 1) <= u) + (__ldg(row + 2) <= u);
            transition(lv, 0, s, a, true, n, r, d);
          }
          if (tid == 0) {
            obs[t] = n;
            if (use_wr) rs_s[t] = static_cast<double>(r);
            else rew[t] = r;
          }
          s = n;
        }
        ucur = unext;
      }
      if (tid == 0) {
        stop_sh = stop;
        if (stop == 0) {
          L_sh = t;
          off_sh = off + t;
          lengths[e] = t;
          done[e] = d;
        } else if (stop >= 2) {
          off_sh = off + t;                     // draws consumed before the failing step
        }
      }
    }
    __syncthreads();
    if (stop_sh != 0) break;
    const int L = L_sh;
    for (int idx = tid; idx <= L; idx += kMcThreads) {                 // G[idx] (:69-70)
      double acc = 0.0;
      const int n = L - idx;
      if (use_wr) {
        const double* r = rs_s + idx;
#pragma unroll 8
        for (int i = 0; i < n; ++i) acc = __dadd_rn(acc, __dmul_rn(wk_s[i], r[i]));
      } else {
        for (int i = 0; i < n; ++i)
          if (__ldg(keep + i)) acc = __dadd_rn(acc, __dmul_rn(__ldg(weights + i), static_cast<double>(rew[idx + i])));
      }
      G[idx] = acc;
    }
    __syncthreads();
    for (int s = tid; s < cells; s += kMcThreads) {                    // per state (:56-91)
      double visits = 0.0, ret = 0.0;
      for (int idx = 0; idx <= L; ++idx) {
        if ((idx == 0 ? start : obs[idx - 1]) != s) continue;
        if (visits != 0.0 && !every_visit) continue;
        visits = __dadd_rn(visits, 1.0);
        ret = __dadd_rn(ret, G[idx]);
      }
      const double tv = __dadd_rn(total_visits[s], visits);
      total_visits[s] = tv;
      if (mode == 2) {
        total_return[s] = __dadd_rn(total_return[s], ret);
      } else if (mode == 0) {
        if (tv > 0.0) {
          const double v = value[s];
          value[s] = __dadd_rn(v, __dmul_rn(__ddiv_rn(1.0, tv), __dadd_rn(ret, -v)));
        }
      } else {
        const double v = value[s];
        value[s] = __dadd_rn(v, __dmul_rn(alpha, __dadd_rn(ret, -v)));
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    meta[0] = e;                 // episodes completed
    meta[1] = off_sh;            // uniform draws consumed
    meta[2] = stop_sh;           // 0 ok, 1 out of draws, 2 + s: state s has an unnormalisable policy row
    meta[3] = e > 0 ? (L_sh > 0 ? obs[L_sh - 1] : __ldg(starts + e - 1)) : -1;   // where the last episode ended
  }
}

}  // namespace gu

using namespace gu;

extern "C" __attribute__((visibility("default"))) int gu_mc_evaluate_f64(
    const gu_levels* lv, const double* cdf, const double* uniforms, int64_t n_uniforms, const int32_t* starts,
    int32_t n_episodes, int32_t max_steps, const double* weights, const uint8_t* keep, int32_t every_visit,
    int32_t mode, double alpha, int32_t* obs_scratch, int32_t* rew_scratch, double* g_scratch, double* total_visits,
    double* total_return, double* value, int32_t* lengths, uint8_t* done, int64_t* meta, void* stream) {
  int rc = check_levels(lv, 1);
  if (rc) return rc;
  if (lv->per_env) return GU_ERR_UNSUPPORTED;
  if (!cdf || !uniforms || !starts || !weights || !keep || !obs_scratch || !rew_scratch || !g_scratch || !total_visits ||
      !total_return || !value || !lengths || !done || !meta)
    return GU_ERR_NULL;
  if (n_episodes < 0 || max_steps < 0 || n_uniforms < 0 || mode < 0 || mode > 2) return GU_ERR_SHAPE;
  const int cells = lv->X * lv->Y;
  const int use_wr = max_steps <= kMcMaxStagedT;
  const size_t wr_bytes = use_wr ? static_cast<size_t>(max_steps) * 16 : 0;   // masked weights + rewards, f64
  const size_t tab_bytes = static_cast<size_t>(cells) * 40;                   // 4 f64 + 4 u16 per cell
  const int use_smem = cells <= 16383 && wr_bytes + tab_bytes <= 200 * 1024;
  const size_t smem = wr_bytes + (use_smem ? tab_bytes : 0);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mc_evaluate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  mc_evaluate_kernel<<<1, kMcThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      view_of(lv, 1), cdf, uniforms, n_uniforms, starts, n_episodes, max_steps, weights, keep, every_visit, mode, alpha,
      cells, obs_scratch, rew_scratch, g_scratch, total_visits, total_return, value, lengths, done,
      reinterpret_cast<long long*>(meta), use_smem, use_wr);
  GU_CHECK_LAUNCH();
  return GU_OK;
}

extern "C" __attribute__((visibility("default"))) int gu_mc_episode_f64(
    int32_t cells, int32_t episode_len, const int32_t* start, const int32_t* obs,
    const int32_t* rewards, int64_t stride, const double* weights, const uint8_t* keep, int32_t every_visit,
    int32_t mode, double alpha, double* g_scratch, double* total_visits, double* total_return, double* value,
    void* stream) {
  if (!start || !obs || !rewards || !weights || !keep || !g_scratch || !total_visits || !total_return || !value)
    return GU_ERR_NULL;
  if (cells <= 0 || episode_len < 0 || stride <= 0 || mode < 0 || mode > 2) return GU_ERR_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int L = episode_len;
  mc_returns_kernel<<<(L + 1 + 127) / 128, 128, 0, st>>>(L, rewards, stride, weights, keep, g_scratch);
  GU_CHECK_LAUNCH();
  mc_update_kernel<<<(cells + 127) / 128, 128, 0, st>>>(cells, L, start, obs, stride, g_scratch, every_visit,
                                                         mode, alpha, total_visits, total_return, value);
  GU_CHECK_LAUNCH();
  return GU_OK;
}

extern "C" __attribute__((visibility("default"))) int gu_mc_finalize_f64(
    int32_t cells, const double* total_visits, const double* total_return, double* value, void* stream) {
  if (!total_visits || !total_return || !value) return GU_ERR_NULL;
  if (cells <= 0) return GU_ERR_SHAPE;
  mc_finalize_kernel<<<(cells + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      cells, total_visits, total_return, value);
  GU_CHECK_LAUNCH();
  return GU_OK;
}
