// gu_plan_tiled.cu -- register-tiled Bellman sweep / greedy extraction for pitched grids.
//
// Memory-bound 2-D stencil: every value of V is read once from HBM and every cell written
// once.  A thread owns CPT consecutive columns (one 16-byte vector: 4 x f32 or 2 x f64) and
// marches down a block of rows keeping a three-row sliding window in registers, so the up /
// down neighbours never touch memory again; left / right neighbours come from the adjacent
// lanes by warp shuffle (the two edge lanes of a warp load one halo column each).
//
// What travels through the window is not V but the per-cell quantities every consumer
// needs:  G[c] = gamma * V[c]  and, for the greedy tie rule,  RT[c] = rint((R[c] + G[c]) * 1e8).
// q[s,a] = R[n] + gamma*V[n] with n = next(s,a) is a function of the landing cell only, so
// computing G / RT once per cell and selecting by the "blocked" bit gives exactly the values
// (same operations, same order) the reference computes per (s, a) pair
// (core/algorithms/utils.py:23-26,65-67); gu_cell.cuh's generic path is the cross-check.
//
// Per-cell static data comes from the derived `info` plane (gu_pack_info): bits 0-3 = action a
// is blocked (grid edge | wall at the target | s terminal), bit 4 goal, bit 5 lava.
#include "gu_cell.cuh"

namespace gu {

template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int CPT = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int CPT = 2; };

template <typename T> __device__ __forceinline__ void unpack(const float4& v, T (&o)[4]) {
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void unpack(const double2& v, double (&o)[2]) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ float4 pack(const float (&o)[4]) { return make_float4(o[0], o[1], o[2], o[3]); }
__device__ __forceinline__ double2 pack(const double (&o)[2]) { return make_double2(o[0], o[1]); }

template <typename T> __device__ __forceinline__ T shfl_up1(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <typename T> __device__ __forceinline__ T shfl_down1(T v) { return __shfl_down_sync(0xffffffffu, v, 1); }

__device__ __forceinline__ float reward_f(uint32_t info, float) {
  return (info & 0x20u) ? -10.0f : ((info & 0x10u) ? 10.0f : -1.0f);
}
__device__ __forceinline__ double reward_f(uint32_t info, double) {
  return (info & 0x20u) ? -10.0 : ((info & 0x10u) ? 10.0 : -1.0);
}

constexpr int kTiledWarps = 4;

// One row of the sliding window: discounted values and rounded q-values of the thread's own
// columns plus the left / right neighbour columns, the raw V (for the residual) and info bytes.
template <typename T, int CPT, bool TIES>
struct WinRow {
  T g[CPT], gl, gr;
  T rt[TIES ? CPT : 1], rtl, rtr;
  T v[CPT];
  uint32_t info;   // CPT info bytes, little-endian
};

template <typename T, bool TIES>
__device__ __forceinline__ void cell_terms(T v, uint32_t info, T gamma, T& g, T& rt) {
  using N = Num<T>;
  g = N::mul(gamma, v);
  if (TIES) {
    const T t = N::mul(N::add(reward_f(info, T(0)), g), N::scale());
    // rint(): the magic-number add is exact below 2^22 (f32) / 2^51 (f64); rare slow path above
    rt = N::add(N::add(t, N::magic()), -N::magic());
    if (!(N::abs(t) < N::magic_limit())) rt = N::rnd(t);
  }
}

template <typename T, int KIND, bool WRITE_TIE>
__global__ void __launch_bounds__(kTiledWarps * 32)
sweep_tiled_kernel(GridView g, const uint8_t* __restrict__ info, const T* __restrict__ vin,
                   T* __restrict__ vout, uint8_t* __restrict__ tie_out, const void* __restrict__ policy,
                   T gamma, T* residual, const T* gate, T gate_thr, int rows_per_block) {
  using N = Num<T>;
  using V = typename Vec<T>::type;
  constexpr int CPT = Vec<T>::CPT;
  constexpr bool TIES = (KIND == GU_POLICY_GREEDY) || WRITE_TIE;
  __shared__ T scratch[kTiledWarps];
  __shared__ T inv_cnt[8];                 // 1/len(ties): exact 1, 1/2, 1/3 (correctly rounded), 1/4
  if (gate != nullptr && *gate < gate_thr) return;
  if (threadIdx.x < 8) inv_cnt[threadIdx.x] = N::inv(threadIdx.x);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int x0 = ((blockIdx.x * kTiledWarps + (threadIdx.x >> 5)) * 32 + lane) * CPT;
  const int rows = g.row_end - g.row_begin;
  const int ry0 = blockIdx.y * rows_per_block;
  const int ry1 = min(ry0 + rows_per_block, rows);
  const bool active = x0 < g.X;                       // lanes past the grid still take part in shuffles
  const bool has_l = active && lane == 0 && x0 > 0;   // edge lanes fetch one halo column each
  const bool has_r = active && lane == 31 && x0 + CPT < g.X;
  const size_t pitch = g.pitch;

  WinRow<T, CPT, TIES> w[3];

  auto load_row = [&](int ar, WinRow<T, CPT, TIES>& r) {
    const size_t o = static_cast<size_t>(ar) * pitch + x0;
    T hv = T(0);
    uint32_t hinfo = 0;
    if (active) {
      unpack(*reinterpret_cast<const V*>(vin + o), r.v);
      r.info = CPT == 4 ? *reinterpret_cast<const uint32_t*>(info + o)
                        : static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(info + o));
      if (has_l) { hv = vin[o - 1]; hinfo = info[o - 1]; }
      if (has_r) { hv = vin[o + CPT]; hinfo = info[o + CPT]; }
    } else {
#pragma unroll
      for (int j = 0; j < CPT; ++j) r.v[j] = T(0);
      r.info = 0;
    }
    T rtj = T(0);
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      cell_terms<T, TIES>(r.v[j], (r.info >> (8 * j)) & 0xffu, gamma, r.g[j], rtj);
      if constexpr (TIES) r.rt[j] = rtj;
    }
    T hg, hrt = T(0);
    cell_terms<T, TIES>(hv, hinfo, gamma, hg, hrt);
    r.gl = shfl_up1(r.g[CPT - 1]);
    r.gr = shfl_down1(r.g[0]);
    if (lane == 0) r.gl = hg;
    if (lane == 31) r.gr = hg;
    if constexpr (TIES) {
      r.rtl = shfl_up1(r.rt[CPT - 1]);
      r.rtr = shfl_down1(r.rt[0]);
      if (lane == 0) r.rtl = hrt;
      if (lane == 31) r.rtr = hrt;
    }
  };

  T dmax = N::neg_inf();
  load_row(ry0, w[0]);        // array row ry0     = row above the first owned row
  load_row(ry0 + 1, w[1]);    // array row ry0 + 1 = first owned row

  for (int base = ry0; base < ry1; base += 3) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int ry = base + j;
      if (ry < ry1) {
        WinRow<T, CPT, TIES>& up = w[j % 3];
        WinRow<T, CPT, TIES>& cur = w[(j + 1) % 3];
        WinRow<T, CPT, TIES>& dn = w[(j + 2) % 3];
        load_row(ry + 2, dn);
        if (active) {
          const size_t o = static_cast<size_t>(ry + 1) * pitch + x0;
          T out[CPT];
          uint32_t ties = 0;
          uint32_t pm = 0;
          if (KIND == GU_POLICY_MASK)
            pm = CPT == 4 ? *reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(policy) + o)
                          : static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(static_cast<const uint8_t*>(policy) + o));
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const uint32_t inf = (cur.info >> (8 * c)) & 0xffu;
            const T gs = cur.g[c];
            // discounted value of the landing cell per action: UP, RIGHT, DOWN, LEFT
            T ga[4];
            ga[0] = (inf & 1u) ? gs : up.g[c];
            ga[1] = (inf & 2u) ? gs : (c == CPT - 1 ? cur.gr : cur.g[c + 1 < CPT ? c + 1 : c]);
            ga[2] = (inf & 4u) ? gs : dn.g[c];
            ga[3] = (inf & 8u) ? gs : (c == 0 ? cur.gl : cur.g[c > 0 ? c - 1 : c]);
            const T rs = reward_f(inf, T(0));
            bool tie[4] = {true, true, true, true};
            if constexpr (TIES) {
              const T rts = cur.rt[c];
              T ra[4];
              ra[0] = (inf & 1u) ? rts : up.rt[c];
              ra[1] = (inf & 2u) ? rts : (c == CPT - 1 ? cur.rtr : cur.rt[c + 1 < CPT ? c + 1 : c]);
              ra[2] = (inf & 4u) ? rts : dn.rt[c];
              ra[3] = (inf & 8u) ? rts : (c == 0 ? cur.rtl : cur.rt[c > 0 ? c - 1 : c]);
              const T m = fmax(fmax(ra[0], ra[1]), fmax(ra[2], ra[3]));
              const bool live = !(inf & 0x30u);       // terminal rows are all zero (utils.py:70)
#pragma unroll
              for (int a = 0; a < 4; ++a) tie[a] = (ra[a] == m) && live;
              if (WRITE_TIE)
                ties |= ((tie[0] ? 1u : 0u) | (tie[1] ? 2u : 0u) | (tie[2] ? 4u : 0u) | (tie[3] ? 8u : 0u)) << (8 * c);
            }
            if (!WRITE_TIE) {
              T acc = rs;
              if (KIND == GU_POLICY_PROBS) {
                const T* pp = static_cast<const T*>(policy) + (o + c) * 4;
#pragma unroll
                for (int a = 0; a < 4; ++a) acc = N::add(acc, N::mul(pp[a], ga[a]));
              } else {
                if (KIND == GU_POLICY_MASK) {
#pragma unroll
                  for (int a = 0; a < 4; ++a) tie[a] = (pm >> (8 * c + a)) & 1u;
                }
                T p = T(0.25);
                if (KIND != GU_POLICY_UNIFORM)
                  p = inv_cnt[(tie[0] ? 1 : 0) + (tie[1] ? 1 : 0) + (tie[2] ? 1 : 0) + (tie[3] ? 1 : 0)];
#pragma unroll
                for (int a = 0; a < 4; ++a)
                  if (tie[a]) acc = N::add(acc, N::mul(p, ga[a]));
              }
              out[c] = acc;
            }
          }
          if (!WRITE_TIE) {
            if (x0 + CPT <= g.X) {
#pragma unroll
              for (int c = 0; c < CPT; ++c) dmax = fmax(dmax, N::add(cur.v[c], -out[c]));
            } else {
#pragma unroll
              for (int c = 0; c < CPT; ++c)
                if (x0 + c < g.X) dmax = fmax(dmax, N::add(cur.v[c], -out[c]));
            }
          }
          if (WRITE_TIE) {
            if (CPT == 4) *reinterpret_cast<uint32_t*>(tie_out + o) = ties;
            else *reinterpret_cast<uint16_t*>(tie_out + o) = static_cast<uint16_t>(ties);
          } else {
            *reinterpret_cast<V*>(vout + o) = pack(out);
          }
        }
      }
    }
  }
  if (!WRITE_TIE) {
    dmax = warp_max(dmax);
    if (lane == 0) scratch[threadIdx.x >> 5] = dmax;
    __syncthreads();
    if (threadIdx.x == 0 && residual != nullptr) {
      T m = scratch[0];
#pragma unroll
      for (int k = 1; k < kTiledWarps; ++k) m = scratch[k] > m ? scratch[k] : m;
      atomic_max_signed(residual, m);
    }
  }
}

// info plane from the three bit planes (one thread per cell; run once per level / shard)
__global__ void __launch_bounds__(256)
pack_info_kernel(GridView g, uint8_t* __restrict__ info) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int ar = blockIdx.y;                         // array row, ghost rows included
  const int y = g.row_begin - 1 + ar;
  if (x >= g.pitch) return;
  uint32_t v = 0;
  if (x < g.X && y >= 0 && y < g.Y) {
    const size_t wb = static_cast<size_t>(ar) * g.pitch_words;
    auto bit = [&](const uint32_t* plane, int drow, int xx) -> uint32_t {
      return (plane[wb + static_cast<ptrdiff_t>(drow) * g.pitch_words + (xx >> 5)] >> (xx & 31)) & 1u;
    };
    const uint32_t goal = bit(g.goal, 0, x), lava = bit(g.lava, 0, x);
    const bool term = goal | lava;
    const bool inner_up = ar > 0, inner_dn = ar < (g.row_end - g.row_begin + 1);   // neighbour row is in the arrays
    const uint32_t bu = (term || y == 0 || (inner_up && bit(g.wall, -1, x))) ? 1u : 0u;
    const uint32_t br = (term || x == g.X - 1 || bit(g.wall, 0, x + 1)) ? 2u : 0u;
    const uint32_t bd = (term || y == g.Y - 1 || (inner_dn && bit(g.wall, 1, x))) ? 4u : 0u;
    const uint32_t bl = (term || x == 0 || bit(g.wall, 0, x - 1)) ? 8u : 0u;
    v = bu | br | bd | bl | (goal << 4) | (lava << 5);
  }
  info[static_cast<size_t>(ar) * g.pitch + x] = static_cast<uint8_t>(v);
}

static inline GridView tview(const gu_grid* g) {
  GridView v;
  v.X = g->X; v.Y = g->Y; v.row_begin = g->row_begin; v.row_end = g->row_end;
  v.pitch = g->pitch; v.pitch_words = g->pitch_words;
  v.wall = g->wall; v.goal = g->goal; v.lava = g->lava;
  return v;
}

static inline bool tiled_ok(const gu_grid* g, size_t elem, const void* a, const void* b) {
  if (g->info == nullptr) return false;
  if ((static_cast<size_t>(g->pitch) * elem) % 16 != 0 || g->pitch % 4 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(a) & 15u) || (reinterpret_cast<uintptr_t>(b) & 15u)) return false;
  if (reinterpret_cast<uintptr_t>(g->info) & 3u) return false;
  return true;
}

template <typename T, bool WRITE_TIE>
static int launch_tiled(const gu_grid* g, const T* vin, T* vout, uint8_t* tie, int kind, const void* policy,
                        T gamma, T* residual, const T* gate, T gate_thr, cudaStream_t st) {
  constexpr int CPT = Vec<T>::CPT;
  const int rows = g->row_end - g->row_begin;
  const int cols_per_block = kTiledWarps * 32 * CPT;
  int rpb = 48;                                     // rows per block (halo re-read: 2/48 = 4 %)
  dim3 grid((g->X + cols_per_block - 1) / cols_per_block, (rows + rpb - 1) / rpb);
  if (grid.y > 65535u) return GU_ERR_SHAPE;
  const GridView v = tview(g);
  const uint8_t* info = g->info;
#define GU_LAUNCH(KIND)                                                                                   \
  sweep_tiled_kernel<T, KIND, WRITE_TIE><<<grid, kTiledWarps * 32, 0, st>>>(v, info, vin, vout, tie, policy, \
                                                                           gamma, residual, gate, gate_thr, rpb)
  if (WRITE_TIE) {
    GU_LAUNCH(GU_POLICY_GREEDY);
  } else {
    switch (kind) {
      case GU_POLICY_PROBS: GU_LAUNCH(GU_POLICY_PROBS); break;
      case GU_POLICY_MASK: GU_LAUNCH(GU_POLICY_MASK); break;
      case GU_POLICY_UNIFORM: GU_LAUNCH(GU_POLICY_UNIFORM); break;
      case GU_POLICY_GREEDY: GU_LAUNCH(GU_POLICY_GREEDY); break;
      default: return GU_ERR_MODE;
    }
  }
#undef GU_LAUNCH
  GU_CHECK_LAUNCH();
  return GU_OK;
}

int sweep_tiled_f32(const gu_grid* g, const float* vin, float* vout, int kind, const void* policy, float gamma,
                    float* residual, const float* gate, float gate_thr, cudaStream_t st) {
  if (!tiled_ok(g, sizeof(float), vin, vout)) return GU_ERR_UNSUPPORTED;
  if (kind == GU_POLICY_PROBS && (reinterpret_cast<uintptr_t>(policy) & 15u)) return GU_ERR_UNSUPPORTED;
  return launch_tiled<float, false>(g, vin, vout, nullptr, kind, policy, gamma, residual, gate, gate_thr, st);
}
int sweep_tiled_f64(const gu_grid* g, const double* vin, double* vout, int kind, const void* policy, double gamma,
                    double* residual, const double* gate, double gate_thr, cudaStream_t st) {
  if (!tiled_ok(g, sizeof(double), vin, vout)) return GU_ERR_UNSUPPORTED;
  return launch_tiled<double, false>(g, vin, vout, nullptr, kind, policy, gamma, residual, gate, gate_thr, st);
}
int greedy_tiled_f32(const gu_grid* g, const float* v, uint8_t* tie, float gamma, cudaStream_t st) {
  if (!tiled_ok(g, sizeof(float), v, tie)) return GU_ERR_UNSUPPORTED;
  return launch_tiled<float, true>(g, v, nullptr, tie, GU_POLICY_GREEDY, nullptr, gamma, nullptr, nullptr, 0.f, st);
}
int greedy_tiled_f64(const gu_grid* g, const double* v, uint8_t* tie, double gamma, cudaStream_t st) {
  if (!tiled_ok(g, sizeof(double), v, tie)) return GU_ERR_UNSUPPORTED;
  return launch_tiled<double, true>(g, v, nullptr, tie, GU_POLICY_GREEDY, nullptr, gamma, nullptr, nullptr, 0.0, st);
}

}  // namespace gu

using namespace gu;

extern "C" __attribute__((visibility("default"))) int gu_pack_info(const gu_grid* g, uint8_t* info, void* stream) {
  if (!g || !g->wall || !g->goal || !g->lava || !info) return GU_ERR_NULL;
  if (g->X <= 0 || g->Y <= 0 || g->row_begin < 0 || g->row_end > g->Y || g->row_begin >= g->row_end ||
      g->pitch < g->X || g->pitch_words * 32 < g->X)
    return GU_ERR_SHAPE;
  const int arows = g->row_end - g->row_begin + 2;
  if (arows > 65535) return GU_ERR_SHAPE;
  dim3 grid((g->pitch + 255) / 256, arows);
  pack_info_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(tview(g), info);
  GU_CHECK_LAUNCH();
  return GU_OK;
}
