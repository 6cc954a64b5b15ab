// gu_cell.cuh -- the per-cell Bellman update shared by every sweep / greedy kernel.
//
// One definition of the arithmetic (evaluation order, rounding, tie rule) so the
// generic, tiled and single-block kernels are bit-identical by construction.
// Reference: core/algorithms/utils.py:15-27 (sweep) and :55-72 (greedy tie set).
#pragma once
#include "gu_common.cuh"

namespace gu {

struct GridView {
  int X, Y, row_begin, row_end, pitch, pitch_words;
  const uint32_t* wall;
  const uint32_t* goal;
  const uint32_t* lava;
};

// Everything one cell needs from its 5-point neighbourhood.
template <typename T>
struct CellIn {
  T vs;          // v[s]
  T vn[4];       // v of the neighbour in direction a (unused where blk bit a is set)
  int rs;        // R[s]
  int rn[4];     // R[neighbour a] (unused where blocked)
  uint32_t blk;  // bit a set: action a leaves the agent in s
                 // (grid edge | wall at the target | s terminal; griduniverse_env.py:51-54,145-149)
  bool term;     // s is terminal (goal or lava)
};

// gamma * v[next(s,a)] for the four actions.
template <typename T>
__device__ __forceinline__ void discounted_next(const CellIn<T>& c, T gamma, T (&g)[4]) {
  using N = Num<T>;
  const T gs = N::mul(gamma, c.vs);
#pragma unroll
  for (int a = 0; a < 4; ++a) g[a] = ((c.blk >> a) & 1u) ? gs : N::mul(gamma, c.vn[a]);
}

// Tie set of utils.py:67: around(q,8) == around(max q,8)  <=>  rint(q*1e8) == max rint(q*1e8).
template <typename T>
__device__ __forceinline__ uint32_t tie_mask_of(const CellIn<T>& c, const T (&g)[4]) {
  using N = Num<T>;
  if (c.term) return 0u;  // utils.py:70: terminal rows are all zero
  T t[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = ((c.blk >> a) & 1u) ? c.rs : c.rn[a];   // reward of the landing state (:65-66)
    t[a] = N::mul(N::add(static_cast<T>(r), g[a]), N::scale());
  }
  // round-half-even to an integer.  Below magic_limit the magic-number add is exact and
  // avoids the slow conversion pipe; otherwise fall back to rint().
  const T lim = N::magic_limit();
  T r[4];
  if (N::abs(t[0]) < lim && N::abs(t[1]) < lim && N::abs(t[2]) < lim && N::abs(t[3]) < lim) {
#pragma unroll
    for (int a = 0; a < 4; ++a) r[a] = N::add(N::add(t[a], N::magic()), -N::magic());
  } else {
#pragma unroll
    for (int a = 0; a < 4; ++a) r[a] = N::rnd(t[a]);
  }
  T m = r[0];
#pragma unroll
  for (int a = 1; a < 4; ++a) m = r[a] > m ? r[a] : m;
  uint32_t mask = 0;
#pragma unroll
  for (int a = 0; a < 4; ++a) mask |= (r[a] == m ? 1u : 0u) << a;
  return mask;
}

// v_new[s] for a policy that is uniform on the action subset `mask` (prob 1/popc on set bits).
// Skipping the zero-probability terms is exact: utils.py:26 would add p*x = +-0 to a sum that
// starts at R[s] != 0, and x + (+-0) == x.
template <typename T>
__device__ __forceinline__ T backup_mask(const CellIn<T>& c, const T (&g)[4], uint32_t mask) {
  using N = Num<T>;
  const T p = N::inv(__popc(mask & 15u));
  T acc = static_cast<T>(c.rs);           // 0.0 + R[s]  (utils.py:23)
#pragma unroll
  for (int a = 0; a < 4; ++a)
    if ((mask >> a) & 1u) acc = N::add(acc, N::mul(p, g[a]));   // utils.py:26, left to right
  return acc;
}

// v_new[s] for arbitrary probabilities.
template <typename T>
__device__ __forceinline__ T backup_probs(const CellIn<T>& c, const T (&g)[4], const T (&p)[4]) {
  using N = Num<T>;
  T acc = static_cast<T>(c.rs);
#pragma unroll
  for (int a = 0; a < 4; ++a) acc = N::add(acc, N::mul(p[a], g[a]));
  return acc;
}

// Dispatch on the policy kind (GU_POLICY_*).  `cell` indexes the padded per-cell arrays.
template <typename T, int KIND>
__device__ __forceinline__ T cell_update(const CellIn<T>& c, T gamma, const void* __restrict__ policy,
                                         size_t cell) {
  T g[4];
  discounted_next(c, gamma, g);
  if (KIND == GU_POLICY_PROBS) {
    T p[4];
    const T* pp = static_cast<const T*>(policy) + cell * 4;
#pragma unroll
    for (int a = 0; a < 4; ++a) p[a] = pp[a];
    return backup_probs(c, g, p);
  } else if (KIND == GU_POLICY_MASK) {
    return backup_mask(c, g, static_cast<uint32_t>(static_cast<const uint8_t*>(policy)[cell]));
  } else if (KIND == GU_POLICY_UNIFORM) {
    return backup_mask(c, g, 15u);
  } else {
    return backup_mask(c, g, tie_mask_of(c, g));
  }
}

// Gather a cell's neighbourhood straight from global memory (any X / pitch).
template <typename T>
__device__ __forceinline__ void gather_cell(const GridView& g, const T* __restrict__ vin, int x, int y,
                                            CellIn<T>& c) {
  const int ar = y - g.row_begin + 1;
  const size_t base = static_cast<size_t>(ar) * g.pitch + x;
  const size_t wbase = static_cast<size_t>(ar) * g.pitch_words;
  auto bit = [&](const uint32_t* plane, int drow, int xx) -> bool {
    return (plane[wbase + static_cast<ptrdiff_t>(drow) * g.pitch_words + (xx >> 5)] >> (xx & 31)) & 1u;
  };
  const bool goal_s = bit(g.goal, 0, x), lava_s = bit(g.lava, 0, x);
  c.term = goal_s | lava_s;
  c.rs = reward_of(goal_s, lava_s);
  c.vs = vin != nullptr ? vin[base] : T(0);      // vin == nullptr: value function of zeros
  c.blk = 0;
  const int dx[4] = {0, 1, 0, -1}, dy[4] = {-1, 0, 1, 0};
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int nx = x + dx[a], ny = y + dy[a];
    bool b = (nx < 0) | (nx >= g.X) | (ny < 0) | (ny >= g.Y) | c.term;
    if (!b) b = bit(g.wall, dy[a], nx);
    if (!b) {
      c.vn[a] = vin != nullptr ? vin[base + static_cast<ptrdiff_t>(dy[a]) * g.pitch + dx[a]] : T(0);
      c.rn[a] = reward_of(bit(g.goal, dy[a], nx), bit(g.lava, dy[a], nx));
    } else {
      c.vn[a] = c.vs;
      c.rn[a] = c.rs;
    }
    c.blk |= (b ? 1u : 0u) << a;
  }
}

}  // namespace gu
