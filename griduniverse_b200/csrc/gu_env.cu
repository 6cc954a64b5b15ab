// gu_env.cu -- batched GridUniverseEnv.step / look_step_ahead / rollouts (sm_100a).
//
// Reference: core/envs/griduniverse_env.py:51-54 (edge-clamped moves), :136-155
// (look_step_ahead), :157-174 (wall / terminal predicates), :176-193 (_step / _reset).
// One thread owns one env (four consecutive envs in the vectorised variants, so that
// positions, actions and outputs move as coalesced 16-byte requests).
#include "gu_env.cuh"

namespace gu {

// ---- one step per launch ("gym mode") ---------------------------------------------------
// VEC = 4: each thread moves four consecutive envs with int4 / uchar4 requests.
// SMEM: one shared level for the whole batch -- its three bit planes are staged in shared memory once
// per block, so the (up to five) plane look-ups of a step never leave the SM.
template <int VEC, bool SMEM = false>
__global__ void __launch_bounds__(256)
step_kernel(LevelsView lv, const int32_t* __restrict__ actions, int32_t* __restrict__ pos,
            int32_t* __restrict__ obs, int32_t* __restrict__ reward, uint8_t* __restrict__ done,
            const int32_t* __restrict__ start_choice, int64_t* stats, uint32_t flags) {
  extern __shared__ uint32_t planes_s[];
  if (SMEM) {
    for (int w = threadIdx.x; w < lv.words; w += blockDim.x) {
      planes_s[w] = __ldg(lv.wall + w);
      planes_s[lv.words + w] = __ldg(lv.goal + w);
      planes_s[2 * lv.words + w] = __ldg(lv.lava + w);
    }
    __syncthreads();
    lv.wall = planes_s;
    lv.goal = planes_s + lv.words;
    lv.lava = planes_s + 2 * lv.words;
  }
  const int64_t i0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * VEC;
  const bool care = !(flags & GU_FLAG_NO_CARE_TERMINAL);
  const bool auto_reset = flags & GU_FLAG_AUTO_RESET;
  long long rsum = 0, dcnt = 0;
  if (i0 < lv.N) {
    int a[VEC], s[VEC], n[VEC], r[VEC], nxt[VEC];
    bool d[VEC];
    if (VEC == 4) {
      const int4 av = *reinterpret_cast<const int4*>(actions + i0);
      const int4 sv = *reinterpret_cast<const int4*>(pos + i0);
      a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
      s[0] = sv.x; s[1] = sv.y; s[2] = sv.z; s[3] = sv.w;
    } else {
      a[0] = actions[i0];
      s[0] = pos[i0];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      transition<SMEM>(lv, i0 + k, s[k], a[k], care, n[k], r[k], d[k]);
      nxt[k] = n[k];
      if (auto_reset && d[k]) nxt[k] = start_choice ? __ldg(start_choice + i0 + k) : start_of(lv, i0 + k);
      rsum += r[k];
      dcnt += d[k] ? 1 : 0;
    }
    if (VEC == 4) {
      *reinterpret_cast<int4*>(pos + i0) = make_int4(nxt[0], nxt[1], nxt[2], nxt[3]);
      if (obs) *reinterpret_cast<int4*>(obs + i0) = make_int4(n[0], n[1], n[2], n[3]);
      if (reward) *reinterpret_cast<int4*>(reward + i0) = make_int4(r[0], r[1], r[2], r[3]);
      if (done) *reinterpret_cast<uchar4*>(done + i0) = make_uchar4(d[0], d[1], d[2], d[3]);
    } else {
      pos[i0] = nxt[0];
      if (obs) obs[i0] = n[0];
      if (reward) reward[i0] = r[0];
      if (done) done[i0] = d[0];
    }
  }
  publish_stats_block(rsum, dcnt, stats);
}

// Per-env levels of at most 64 cells (cfg 4: 8x8): the whole level is WORDS words per plane, so
// instead of chasing pos -> wall word -> landing cell -> goal / lava word (three dependent
// loads), a thread fetches every plane word of its four envs up front as independent 16-byte
// requests and does the step on 64-bit masks in registers.  Same bytes, no dependent chain.
template <int WORDS>
__global__ void __launch_bounds__(256)
step_small_kernel(LevelsView lv, const int32_t* __restrict__ actions, const int32_t* pos_in, int32_t* pos,
                  int32_t* __restrict__ obs, int32_t* __restrict__ reward, uint8_t* __restrict__ done,
                  const int32_t* __restrict__ start_choice, int64_t* stats, uint32_t flags) {
  const int64_t i0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  const bool care = !(flags & GU_FLAG_NO_CARE_TERMINAL);
  const bool auto_reset = flags & GU_FLAG_AUTO_RESET;
  long long rsum = 0, dcnt = 0;
  if (i0 < lv.N) {
    const int4 av = *reinterpret_cast<const int4*>(actions + i0);
    const int4 sv = *reinterpret_cast<const int4*>(pos_in + i0);
    uint4 pw[WORDS], pg[WORDS], pl[WORDS];
#pragma unroll
    for (int k = 0; k < WORDS; ++k) {
      pw[k] = __ldg(reinterpret_cast<const uint4*>(lv.wall + k * lv.N + i0));
      pg[k] = __ldg(reinterpret_cast<const uint4*>(lv.goal + k * lv.N + i0));
      pl[k] = __ldg(reinterpret_cast<const uint4*>(lv.lava + k * lv.N + i0));
    }
    const int a[4] = {av.x, av.y, av.z, av.w};
    const int s[4] = {sv.x, sv.y, sv.z, sv.w};
    int n[4], r[4], nxt[4];
    bool d[4];
    auto lane_of = [](const uint4& v, int e) -> uint32_t { return e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w; };
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      unsigned long long W = lane_of(pw[0], e), G = lane_of(pg[0], e), L = lane_of(pl[0], e);
      if (WORDS == 2) {
        W |= static_cast<unsigned long long>(lane_of(pw[1], e)) << 32;
        G |= static_cast<unsigned long long>(lane_of(pg[1], e)) << 32;
        L |= static_cast<unsigned long long>(lane_of(pl[1], e)) << 32;
      }
      const unsigned long long Tm = G | L;
      const bool stay = care && ((Tm >> s[e]) & 1ull);
      const int c = clamp_move(s[e], a[e], lv.X, lv.Y);
      n[e] = (stay || ((W >> c) & 1ull)) ? s[e] : c;
      const bool g = (G >> n[e]) & 1ull, l = (L >> n[e]) & 1ull;
      r[e] = reward_of(g, l);
      d[e] = g | l;
      nxt[e] = n[e];
      if (auto_reset && d[e]) nxt[e] = start_choice ? __ldg(start_choice + i0 + e) : __ldg(lv.start + i0 + e);
      rsum += r[e];
      dcnt += d[e] ? 1 : 0;
    }
    if (pos) *reinterpret_cast<int4*>(pos + i0) = make_int4(nxt[0], nxt[1], nxt[2], nxt[3]);   // NULL: look_step_ahead
    if (obs) *reinterpret_cast<int4*>(obs + i0) = make_int4(n[0], n[1], n[2], n[3]);
    if (reward) *reinterpret_cast<int4*>(reward + i0) = make_int4(r[0], r[1], r[2], r[3]);
    if (done) *reinterpret_cast<uchar4*>(done + i0) = make_uchar4(d[0], d[1], d[2], d[3]);
  }
  publish_stats_block(rsum, dcnt, stats);
}

// ---- T steps per launch, layout-agnostic version ------------------------------------------
// One thread per env, position in a register, actions[t][n] read coalesced across the warp.
__global__ void __launch_bounds__(256)
rollout_generic_kernel(LevelsView lv, int64_t T, const int32_t* __restrict__ actions,
                       int32_t* __restrict__ pos, int32_t* __restrict__ obs,
                       int32_t* __restrict__ reward, uint8_t* __restrict__ done,
                       const int32_t* __restrict__ start_choice, int32_t* __restrict__ env_return,
                       int32_t* __restrict__ env_done, int64_t* stats, uint32_t flags) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool care = !(flags & GU_FLAG_NO_CARE_TERMINAL);
  const bool auto_reset = flags & GU_FLAG_AUTO_RESET;
  long long rsum = 0, dcnt = 0;
  if (i < lv.N) {
    int s = pos[i];
    const int st = auto_reset && !start_choice ? start_of(lv, i) : 0;    // lv.start may be NULL otherwise
    const bool packed = flags & GU_FLAG_PACKED_ACTIONS;
    uint32_t word = 0;
    for (int64_t t = 0; t < T; ++t) {
      const int64_t o = t * lv.N + i;
      int n, r;
      bool d;
      int a;
      if (packed) {                                   // 16 steps per word: uint32[ceil(T/16)][N]
        if ((t & 15) == 0) word = static_cast<uint32_t>(__ldg(actions + (t >> 4) * lv.N + i));
        a = static_cast<int>((word >> (2 * (t & 15))) & 3u);
      } else {
        a = __ldg(actions + o);
      }
      transition(lv, i, s, a, care, n, r, d);
      if (obs) obs[o] = n;
      if (reward) reward[o] = r;
      if (done) done[o] = d;
      s = n;
      if (auto_reset && d) s = start_choice ? __ldg(start_choice + o) : st;
      rsum += r;
      dcnt += d ? 1 : 0;
    }
    pos[i] = s;
    const bool acc = flags & GU_FLAG_ACCUMULATE;
    if (env_return) env_return[i] = static_cast<int32_t>(rsum) + (acc ? env_return[i] : 0);
    if (env_done) env_done[i] = static_cast<int32_t>(dcnt) + (acc ? env_done[i] : 0);
  }
  publish_stats(rsum, dcnt, stats);
}

__global__ void __launch_bounds__(256)
look_kernel(LevelsView lv, int64_t m, const int32_t* __restrict__ states,
            const int32_t* __restrict__ actions, int32_t* __restrict__ next,
            int32_t* __restrict__ reward, uint8_t* __restrict__ terminal, uint32_t flags) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= m) return;
  int n, r;
  bool d;
  transition(lv, i, states[i], actions[i], !(flags & GU_FLAG_NO_CARE_TERMINAL), n, r, d);
  if (next) next[i] = n;
  if (reward) reward[i] = r;
  if (terminal) terminal[i] = d;
}

// ---- resident look_step_ahead service for the one-env step loop --------------------------------------
// The reference's `env.step(a)` is one (state, action) -> (next, reward, done) lookup
// (griduniverse_env.py:176-185).  One launch + one stream synchronise per lookup costs ~20 us; this kernel
// instead stays resident while requests keep coming and answers each one over PCIe: the host writes an
// 8-byte request word into a pinned mailbox, thread 0 polls it (system-scope loads), evaluates the same
// transition() every other kernel uses, and stores the 16-byte answer back into the mailbox; the host
// spins on the answer's sequence number.  The kernel leaves by itself after `idle_cycles` without a
// request (and in any case after `max_cycles`), clearing the mailbox's `alive` word, so device-wide
// synchronisation points are never held up for long; the host relaunches it on demand.
//   mailbox (32 x uint32, 128-byte aligned, pinned host memory):
//     [0:2]   request, one uint64: bits 0-31 sequence number, 32-33 action, 34 "do not care about
//             terminals" (care_about_terminal=False), 35-63 state
//     [16:20] answer, one 16-byte store: {sequence number, next state, reward, terminal}
//     [20]    alive: set by the host before the launch, cleared by the kernel when it leaves
__global__ void __launch_bounds__(32)
look_server_kernel(LevelsView lv, const uint64_t* req, uint32_t* ack, uint32_t* alive, uint32_t seq0,
                   long long idle_cycles, long long max_cycles) {
  if (threadIdx.x != 0) return;
  uint32_t last = seq0;
  const long long born = clock64();
  long long idle_since = born;
  for (;;) {
    uint64_t r;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(r) : "l"(req) : "memory");
    const uint32_t seq = static_cast<uint32_t>(r);
    if (seq != last) {
      const uint32_t hi = static_cast<uint32_t>(r >> 32);
      int n, rew;
      bool term;
      transition(lv, 0, static_cast<int>(hi >> 3), static_cast<int>(hi & 3u), (hi & 4u) == 0, n, rew, term);
      asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(ack), "r"(seq), "r"(n), "r"(rew),
                   "r"(term ? 1u : 0u) : "memory");
      last = seq;
      idle_since = clock64();
    } else {
      const long long now = clock64();
      if (now - idle_since > idle_cycles || now - born > max_cycles) break;
    }
  }
  __threadfence_system();
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(alive), "r"(0u) : "memory");
}

// Policy-driven episodes (monte_carlo.py:7-26): one thread per episode, shared level.
__global__ void __launch_bounds__(256)
rollout_policy_kernel(LevelsView lv, int64_t T, const double* __restrict__ cdf,
                      const double* __restrict__ uniforms, int32_t* __restrict__ pos,
                      int32_t* __restrict__ obs, int32_t* __restrict__ reward,
                      int32_t* __restrict__ length, uint8_t* __restrict__ done) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= lv.N) return;
  int s = pos[i];
  bool d = false;
  int64_t t = 0;
  for (; t < T && !d; ++t) {
    const int64_t o = t * lv.N + i;
    const double u = __ldg(uniforms + o);
    const double* row = cdf + static_cast<int64_t>(s) * 4;
    if (__ldg(row + 3) != __ldg(row + 3)) {          // NaN row: np.random.choice would raise in this state
      pos[i] = s;
      if (length) length[i] = static_cast<int32_t>(-1 - t);
      if (done) done[i] = 0;
      return;
    }
    // searchsorted(cdf, u, side='right'): number of entries <= u (cdf[3] == 1 > u)
    int a = (__ldg(row) <= u) + (__ldg(row + 1) <= u) + (__ldg(row + 2) <= u);
    int n, r;
    transition(lv, i, s, a, true, n, r, d);
    if (obs) obs[o] = n;
    if (reward) reward[o] = r;
    s = n;
  }
  pos[i] = s;
  if (length) length[i] = static_cast<int32_t>(t);
  if (done) done[i] = d;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

// Table-driven fast paths (gu_env_tables.cu).
int rollout_tables(const gu_levels* lv, int64_t n, int64_t T, const int32_t* actions, int32_t* pos,
                   int32_t* obs, int32_t* reward, uint8_t* done, const int32_t* start_choice,
                   int32_t* env_return, int32_t* env_done, int64_t* stats, const uint32_t* tables,
                   uint32_t flags, cudaStream_t st);

// int32 actions [T][N] -> packed 2-bit actions uint32[ceil(T/16)][N] (GU_FLAG_PACKED_ACTIONS)
__global__ void __launch_bounds__(256)
pack_actions_kernel(const int32_t* __restrict__ actions, uint32_t* __restrict__ packed, int64_t T, int64_t N) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t w = blockIdx.y;
  if (i >= N) return;
  uint32_t word = 0;
#pragma unroll
  for (int s = 0; s < 16; ++s) {
    const int64_t t = w * 16 + s;
    if (t < T) word |= (static_cast<uint32_t>(__ldg(actions + t * N + i)) & 3u) << (2 * s);
  }
  packed[w * N + i] = word;
}

}  // namespace gu

#include <cstdlib>
#include <thread>
#include <vector>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define GU_HOST_SIMD 1
#else
#define GU_HOST_SIMD 0
#endif

namespace gu {
// Host packer.  One packed row = 16 input rows read as 16 concurrent streams; the scalar form costs ~50
// instructions per output word (3 GB/s of input per core), so a full 16-step word row is packed with
// vector instructions -- 16 envs (one 64-byte line of every stream) per iteration: AVX2 where the CPU
// has it (checked at run time), SSE2 otherwise -- which leaves the cores waiting on memory instead.
static void pack_row_scalar(const int32_t* in, uint32_t* out, int64_t N, int steps, int64_t i0, int64_t i1) {
  for (int64_t i = i0; i < i1; ++i) {
    uint32_t word = 0;
    for (int s = 0; s < steps; ++s) word |= (static_cast<uint32_t>(in[s * N + i]) & 3u) << (2 * s);
    out[i] = word;
  }
}

#if GU_HOST_SIMD
// Horner form, last step first: word = (...((a15 & 3) << 2 | (a14 & 3)) << 2 ...) | (a0 & 3) -- the shift count is
// the immediate 2 throughout.
__attribute__((target("avx2"))) static void pack_row16_avx2(const int32_t* in, uint32_t* out, int64_t N, int64_t i0,
                                                             int64_t i1) {
  const __m256i three = _mm256_set1_epi32(3);
  int64_t i = i0;
  for (; i + 16 <= i1; i += 16) {
    __m256i a = _mm256_setzero_si256(), b = _mm256_setzero_si256();
#pragma GCC unroll 16
    for (int s = 15; s >= 0; --s) {
      const int32_t* p = in + s * N + i;
      a = _mm256_or_si256(_mm256_slli_epi32(a, 2), _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(p)), three));
      b = _mm256_or_si256(_mm256_slli_epi32(b, 2), _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 8)), three));
    }
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + i), a);
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + i + 8), b);
  }
  pack_row_scalar(in, out, N, 16, i, i1);
}

static void pack_row16_sse2(const int32_t* in, uint32_t* out, int64_t N, int64_t i0, int64_t i1) {
  const __m128i three = _mm_set1_epi32(3);
  int64_t i = i0;
  for (; i + 8 <= i1; i += 8) {
    __m128i a = _mm_setzero_si128(), b = _mm_setzero_si128();
#pragma GCC unroll 16
    for (int s = 15; s >= 0; --s) {
      const int32_t* p = in + s * N + i;
      a = _mm_or_si128(_mm_slli_epi32(a, 2), _mm_and_si128(_mm_loadu_si128(reinterpret_cast<const __m128i*>(p)), three));
      b = _mm_or_si128(_mm_slli_epi32(b, 2), _mm_and_si128(_mm_loadu_si128(reinterpret_cast<const __m128i*>(p + 4)), three));
    }
    _mm_storeu_si128(reinterpret_cast<__m128i*>(out + i), a);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(out + i + 4), b);
  }
  pack_row_scalar(in, out, N, 16, i, i1);
}
#endif

// columns [i0, i1) of every packed row; mode 0 = scalar, 1 = SSE2, 2 = AVX2 (host_pack_mode)
static void pack_cols_host(const int32_t* actions, uint32_t* packed, int64_t T, int64_t N, int64_t i0, int64_t i1,
                           int mode) {
  const int64_t words = (T + 15) / 16;
  for (int64_t w = 0; w < words; ++w) {
    uint32_t* out = packed + w * N;
    const int32_t* in = actions + w * 16 * N;
    const int steps = static_cast<int>(T - w * 16 < 16 ? T - w * 16 : 16);
#if GU_HOST_SIMD
    if (steps == 16 && mode != 0) {
      if (mode == 2) pack_row16_avx2(in, out, N, i0, i1);
      else pack_row16_sse2(in, out, N, i0, i1);
      continue;
    }
#endif
    pack_row_scalar(in, out, N, steps, i0, i1);
  }
}

// decided once per call, before the worker threads start
static int host_pack_mode() {
#if GU_HOST_SIMD
  const char* force = getenv("GU_HOST_PACK");               // developer / test switch: "scalar", "sse2"
  if (force && force[0] == 's' && force[1] == 'c') return 0;
  if (force && force[0] == 's' && force[1] == 's') return 1;
  return __builtin_cpu_supports("avx2") ? 2 : 1;
#else
  return 0;
#endif
}
}  // namespace gu

using namespace gu;

extern "C" __attribute__((visibility("default"))) int gu_pack_actions(const int32_t* actions, int64_t n_steps,
                                                                        int64_t n_envs, uint32_t* packed, void* stream) {
  if (n_steps < 0 || n_envs < 0) return GU_ERR_SHAPE;
  if (n_steps == 0 || n_envs == 0) return GU_OK;
  if (!actions || !packed) return GU_ERR_NULL;
  const int64_t words = (n_steps + 15) / 16;
  if (words > 65535) return GU_ERR_SHAPE;
  dim3 grid(static_cast<unsigned>((n_envs + 255) / 256), static_cast<unsigned>(words));
  pack_actions_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(actions, packed, n_steps, n_envs);
  GU_CHECK_LAUNCH();
  return GU_OK;
}

extern "C" __attribute__((visibility("default"))) int gu_pack_actions_host(const int32_t* actions, int64_t n_steps,
                                                                             int64_t n_envs, uint32_t* packed,
                                                                             int32_t n_threads) {
  if (n_steps < 0 || n_envs < 0) return GU_ERR_SHAPE;
  if (n_steps == 0 || n_envs == 0) return GU_OK;
  if (!actions || !packed) return GU_ERR_NULL;
  int nt = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  if (nt < 1) nt = 1;
  const int64_t chunks = (n_envs + 4095) / 4096;          // at least 16 KB of output per thread
  if (nt > chunks) nt = static_cast<int>(chunks);
  const int mode = host_pack_mode();
  if (nt == 1) {
    pack_cols_host(actions, packed, n_steps, n_envs, 0, n_envs, mode);
    return GU_OK;
  }
  std::vector<std::thread> pool;
  for (int k = 0; k < nt; ++k) {
    const int64_t i0 = (n_envs * k / nt) & ~static_cast<int64_t>(15), i1 = k + 1 == nt ? n_envs : (n_envs * (k + 1) / nt) & ~static_cast<int64_t>(15);
    pool.emplace_back(pack_cols_host, actions, packed, n_steps, n_envs, i0, i1, mode);
  }
  for (auto& t : pool) t.join();
  return GU_OK;
}

extern "C" __attribute__((visibility("default"))) int gu_step(const gu_levels* lv, int64_t n, const int32_t* actions, int32_t* pos, int32_t* obs,
                       int32_t* reward, uint8_t* done, const int32_t* start_choice, int64_t* stats,
                       uint32_t flags, void* stream) {
  int rc = check_levels(lv, n);
  if (rc) return rc;
  if (n == 0) return GU_OK;
  if (!actions || !pos) return GU_ERR_NULL;
  if ((flags & GU_FLAG_AUTO_RESET) && !start_choice && !lv->start) return GU_ERR_NULL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const LevelsView v = view_of(lv, n);
  const bool vec = (n % 4 == 0) && aligned16(actions) && aligned16(pos) && (!obs || aligned16(obs)) &&
                   (!reward || aligned16(reward)) && (!done || aligned4(done));
  const bool small = vec && lv->per_env && lv->words <= 2 && aligned16(lv->wall) && aligned16(lv->goal) &&
                     aligned16(lv->lava);
  if (small) {
    const unsigned blocks = static_cast<unsigned>((n / 4 + 255) / 256);
    if (lv->words == 1)
      step_small_kernel<1><<<blocks, 256, 0, st>>>(v, actions, pos, pos, obs, reward, done, start_choice, stats, flags);
    else
      step_small_kernel<2><<<blocks, 256, 0, st>>>(v, actions, pos, pos, obs, reward, done, start_choice, stats, flags);
  } else if (vec) {
    const int64_t threads = n / 4;
    const unsigned blocks = static_cast<unsigned>((threads + 255) / 256);
    if (!lv->per_env && lv->words <= 4096)      // shared level: planes staged in shared memory (<= 48 KB)
      step_kernel<4, true><<<blocks, 256, 3 * lv->words * sizeof(uint32_t), st>>>(v, actions, pos, obs, reward, done,
                                                                                  start_choice, stats, flags);
    else
      step_kernel<4><<<blocks, 256, 0, st>>>(v, actions, pos, obs, reward, done, start_choice, stats, flags);
  } else {
    const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
    if (!lv->per_env && lv->words <= 4096 && n >= 256)
      step_kernel<1, true><<<blocks, 256, 3 * lv->words * sizeof(uint32_t), st>>>(v, actions, pos, obs, reward, done,
                                                                                  start_choice, stats, flags);
    else
      step_kernel<1><<<blocks, 256, 0, st>>>(v, actions, pos, obs, reward, done, start_choice, stats, flags);
  }
  GU_CHECK_LAUNCH();
  return GU_OK;
}

extern "C" __attribute__((visibility("default"))) int gu_rollout(const gu_levels* lv, int64_t n, int64_t T, const int32_t* actions, int32_t* pos,
                          int32_t* obs, int32_t* reward, uint8_t* done, const int32_t* start_choice,
                          int32_t* env_return, int32_t* env_done, int64_t* stats, const uint32_t* tables,
                          uint32_t flags, void* stream) {
  int rc = check_levels(lv, n);
  if (rc) return rc;
  if (T < 0) return GU_ERR_SHAPE;
  if (n == 0 || T == 0) return GU_OK;      // nothing to do: empty action matrices may be NULL
  if (!actions || !pos) return GU_ERR_NULL;
  if ((flags & GU_FLAG_AUTO_RESET) && !start_choice && !lv->start) return GU_ERR_NULL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (tables) {
    rc = rollout_tables(lv, n, T, actions, pos, obs, reward, done, start_choice, env_return, env_done,
                        stats, tables, flags, st);
    if (rc != GU_ERR_UNSUPPORTED) return rc;
  }
  rollout_generic_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      view_of(lv, n), T, actions, pos, obs, reward, done, start_choice, env_return, env_done, stats, flags);
  GU_CHECK_LAUNCH();
  return GU_OK;
}

extern "C" __attribute__((visibility("default"))) int gu_rollout_policy(
    const gu_levels* lv, int64_t n, int64_t T, const double* cdf, const double* uniforms, int32_t* pos,
    int32_t* obs, int32_t* reward, int32_t* length, uint8_t* done, void* stream) {
  int rc = check_levels(lv, n);
  if (rc) return rc;
  if (!cdf || !uniforms || !pos) return GU_ERR_NULL;
  if (lv->per_env) return GU_ERR_UNSUPPORTED;
  if (T < 0) return GU_ERR_SHAPE;
  if (n == 0) return GU_OK;
  rollout_policy_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      view_of(lv, n), T, cdf, uniforms, pos, obs, reward, length, done);
  GU_CHECK_LAUNCH();
  return GU_OK;
}

extern "C" __attribute__((visibility("default"))) int gu_look_step_ahead(const gu_levels* lv, int64_t m, const int32_t* states,
                                  const int32_t* actions, int32_t* next, int32_t* reward,
                                  uint8_t* terminal, uint32_t flags, void* stream) {
  int rc = check_levels(lv, m);
  if (rc) return rc;
  if (!states || !actions) return GU_ERR_NULL;
  if (m == 0) return GU_OK;
  if (flags & ~static_cast<uint32_t>(GU_FLAG_NO_CARE_TERMINAL)) return GU_ERR_MODE;
  const bool small = lv->per_env && lv->words <= 2 && m % 4 == 0 && aligned16(lv->wall) && aligned16(lv->goal) &&
                     aligned16(lv->lava) && aligned16(states) && aligned16(actions) && (!next || aligned16(next)) &&
                     (!reward || aligned16(reward)) && (!terminal || aligned4(terminal));
  if (small) {   // pair i is looked up in env i's level: the one-step kernel without the position update
    const unsigned blocks = static_cast<unsigned>((m / 4 + 255) / 256);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const LevelsView v = view_of(lv, m);
    if (lv->words == 1)
      step_small_kernel<1><<<blocks, 256, 0, st>>>(v, actions, states, nullptr, next, reward, terminal, nullptr, nullptr, flags);
    else
      step_small_kernel<2><<<blocks, 256, 0, st>>>(v, actions, states, nullptr, next, reward, terminal, nullptr, nullptr, flags);
    GU_CHECK_LAUNCH();
    return GU_OK;
  }
  look_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      view_of(lv, m), m, states, actions, next, reward, terminal, flags);
  GU_CHECK_LAUNCH();
  return GU_OK;
}

extern "C" __attribute__((visibility("default"))) int gu_look_server_start(const gu_levels* lv, void* mailbox,
                                                                           uint32_t seq0, int64_t idle_cycles,
                                                                           int64_t max_cycles, void* stream) {
  int rc = check_levels(lv, 1);
  if (rc) return rc;
  if (lv->per_env) return GU_ERR_UNSUPPORTED;
  if (!mailbox) return GU_ERR_NULL;
  if (reinterpret_cast<uintptr_t>(mailbox) & 127u) return GU_ERR_ALIGN;
  if (static_cast<int64_t>(lv->X) * lv->Y > (1ll << 29) || idle_cycles <= 0 || max_cycles <= 0) return GU_ERR_SHAPE;
  uint32_t* mb = static_cast<uint32_t*>(mailbox);
  look_server_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      view_of(lv, 1), reinterpret_cast<const uint64_t*>(mb), mb + 16, mb + 20, seq0, idle_cycles, max_cycles);
  GU_CHECK_LAUNCH();
  return GU_OK;
}
