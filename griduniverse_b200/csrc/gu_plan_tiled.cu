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
// Per-cell static data comes from the derived `info` plane (gu_pack_info): one bit per action
// "a is blocked" (grid edge | wall at the target | s terminal) plus the goal and lava bits.
#include <cuda.h>

#include <cstdlib>

#include "gu_cell.cuh"

namespace gu {

template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int W = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int W = 2; };

__device__ __forceinline__ void unpack(const float4& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void unpack(const double2& v, double* o) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ float4 pack(const float* o) { return make_float4(o[0], o[1], o[2], o[3]); }
__device__ __forceinline__ double2 pack(const double* o) { return make_double2(o[0], o[1]); }

template <typename T> __device__ __forceinline__ T shfl_up1(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <typename T> __device__ __forceinline__ T shfl_down1(T v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// info byte layout (gu_pack_info): bit0 UP blocked, bit1 RIGHT blocked, bit2 DOWN blocked,
// bit3 goal, bit4 lava, bit5 LEFT blocked.  (info & 0x18) is the byte offset of the cell's entry
// in the 4-entry {reward, poison} table below (8-byte entries for f32, << 1 for f64).
constexpr uint32_t kBlkU = 1u, kBlkR = 2u, kBlkD = 4u, kGoal = 8u, kLava = 16u, kBlkL = 32u;

// Per-block lookup tables in shared memory.
template <typename T>
struct Luts {
  T reward[4][2];   // [goal | lava<<1] -> {R (griduniverse_env.py:80-90), poison: 0 or NaN for terminals}
  T inv_cnt[8];     // 1/len(ties) at index len: exact 1, 1/2, 1/3 (correctly rounded), 1/4
  // [len(ties) << 2 | goal | lava<<1] -> {1/len(ties), or 0 in a terminal cell (utils.py:70); R[s]}:
  // one index serves the probability and the reward, and terminals need no NaN "poison"
  T prs[32][2];
};

template <typename T>
__device__ __forceinline__ void init_luts(Luts<T>& l) {
  const int t = threadIdx.x;
  if (t < 4) {
    l.reward[t][0] = (t & 2) ? T(-10) : ((t & 1) ? T(10) : T(-1));
    l.reward[t][1] = t ? T(CUDART_NAN) : T(0);
  }
  if (t < 8) l.inv_cnt[t] = Num<T>::inv(t);
  if (t < 32) {
    l.prs[t][0] = (t & 3) ? T(0) : Num<T>::inv(t >> 2);
    l.prs[t][1] = (t & 2) ? T(-10) : ((t & 1) ? T(10) : T(-1));
  }
}

template <typename T>
__device__ __forceinline__ T reward_lut(const Luts<T>& l, uint32_t inf) {
  return *reinterpret_cast<const T*>(reinterpret_cast<const char*>(&l.reward[0][0]) +
                                     (sizeof(T) == 4 ? (inf & 0x18u) : ((inf & 0x18u) << 1)));
}
// {reward, poison} with one 8 / 16-byte shared load
__device__ __forceinline__ float2 reward_poison_lut(const Luts<float>& l, uint32_t inf) {
  return *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(&l.reward[0][0]) + (inf & 0x18u));
}
__device__ __forceinline__ double2 reward_poison_lut(const Luts<double>& l, uint32_t inf) {
  return *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(&l.reward[0][0]) + ((inf & 0x18u) << 1));
}

// Bellman backup of one cell for the policy that is uniform on the tie set (utils.py:23-26,67-71):
//   acc = R[s];  for a in 0..3 (left to right): if r_a == m: acc = acc + p * g_a,   p = 1/#ties.
// m carries the terminal "poison" (NaN) so a terminal cell has no ties and keeps acc = R[s].
//
// f32: predicate-free.  w_a = (r_a == m) as 1.0f / 0.0f (FSET), the count goes through the
// mantissa of a magic-number FMA chain into the inv_cnt table, and non-tied actions add
// (0 * p) * g_a = +-0, which leaves the non-zero running sum bit-identical.
__device__ __forceinline__ float backup_ties(float rs, const float (&ra)[4], float m, const float (&ga)[4],
                                             const Luts<float>& l) {
  float w[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) asm("set.eq.f32.f32 %0, %1, %2;" : "=f"(w[a]) : "f"(ra[a]), "f"(m));
  float c = 8388608.0f;                                   // 2^23: integers land in the low mantissa bits
#pragma unroll
  for (int a = 0; a < 4; ++a) c = __fmaf_rn(w[a], 4.0f, c);
  // (a table-free reciprocal -- exponent flip for 1, 2, 4 and a select for 3 -- has a shorter
  // dependency chain but more ALU-pipe work; it measured 2 % slower on B200.)
  const float p = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(&l.inv_cnt[0]) +
                                                  (__float_as_uint(c) & 0x1cu));
  // non-tied actions add (0 * (p*g)) = +-0, which leaves the non-zero running sum bit-identical.
  // (Folding w into a fused multiply-add, acc = fma(w, p*g, acc), is also exact and saves four
  // instructions per cell, but measured 2 % slower on B200: the dependent FFMA chain is longer.)
  float acc = rs;
#pragma unroll
  for (int a = 0; a < 4; ++a) acc = __fadd_rn(acc, __fmul_rn(w[a], __fmul_rn(p, ga[a])));
  return acc;
}
// f64: the four compares feed predicated mul / add pairs (PTX).  A/B on B200 at 16384^2:
// 1.38 ms this way, 1.42 ms with C++ selects, 1.54 ms with 0/1 weight multiplies.
// GU_F64_FMA_W: the compare selects a 0.0 / 1.0 weight (one SEL on the high word) and the term goes in as
// acc = fma(w, p * g, acc) -- exact for w in {0, 1}: fma(1, x, acc) rounds acc + x once like the add, and
// fma(0, x, acc) is acc -- instead of a predicated add, which costs a DADD and two 32-bit selects.
#ifndef GU_F64_FMA_W
#define GU_F64_FMA_W 1
#endif
__device__ __forceinline__ double backup_ties(double rs, const double (&ra)[4], double m, const double (&ga)[4],
                                              const Luts<double>& l) {
  double acc;
  const uint32_t lut = static_cast<uint32_t>(__cvta_generic_to_shared(&l.inv_cnt[0]));
#if GU_F64_FMA_W
  asm("{\n\t"
      ".reg .pred t0, t1, t2, t3;\n\t"
      ".reg .u32 c, a;\n\t"
      ".reg .f64 p, x, w;\n\t"
      "setp.eq.f64 t0, %2, %6;\n\t"
      "setp.eq.f64 t1, %3, %6;\n\t"
      "setp.eq.f64 t2, %4, %6;\n\t"
      "setp.eq.f64 t3, %5, %6;\n\t"
      "mov.u32 c, 0;\n\t"
      "@t0 add.u32 c, c, 8;\n\t"
      "@t1 add.u32 c, c, 8;\n\t"
      "@t2 add.u32 c, c, 8;\n\t"
      "@t3 add.u32 c, c, 8;\n\t"
      "add.u32 a, c, %11;\n\t"
      "ld.shared.f64 p, [a];\n\t"
      "mul.rn.f64 x, p, %7;\n\t"
      "selp.f64 w, 0d3FF0000000000000, 0d0000000000000000, t0;\n\t"
      "fma.rn.f64 %0, w, x, %1;\n\t"
      "mul.rn.f64 x, p, %8;\n\t"
      "selp.f64 w, 0d3FF0000000000000, 0d0000000000000000, t1;\n\t"
      "fma.rn.f64 %0, w, x, %0;\n\t"
      "mul.rn.f64 x, p, %9;\n\t"
      "selp.f64 w, 0d3FF0000000000000, 0d0000000000000000, t2;\n\t"
      "fma.rn.f64 %0, w, x, %0;\n\t"
      "mul.rn.f64 x, p, %10;\n\t"
      "selp.f64 w, 0d3FF0000000000000, 0d0000000000000000, t3;\n\t"
      "fma.rn.f64 %0, w, x, %0;\n\t"
      "}"
      : "=&d"(acc)
      : "d"(rs), "d"(ra[0]), "d"(ra[1]), "d"(ra[2]), "d"(ra[3]), "d"(m), "d"(ga[0]), "d"(ga[1]), "d"(ga[2]),
        "d"(ga[3]), "r"(lut));
#else
  asm("{\n\t"
      ".reg .pred t0, t1, t2, t3;\n\t"
      ".reg .u32 c, a;\n\t"
      ".reg .f64 p, x;\n\t"
      "setp.eq.f64 t0, %2, %6;\n\t"
      "setp.eq.f64 t1, %3, %6;\n\t"
      "setp.eq.f64 t2, %4, %6;\n\t"
      "setp.eq.f64 t3, %5, %6;\n\t"
      "mov.u32 c, 0;\n\t"
      "@t0 add.u32 c, c, 8;\n\t"
      "@t1 add.u32 c, c, 8;\n\t"
      "@t2 add.u32 c, c, 8;\n\t"
      "@t3 add.u32 c, c, 8;\n\t"
      "add.u32 a, c, %11;\n\t"
      "ld.shared.f64 p, [a];\n\t"
      "mov.f64 %0, %1;\n\t"
      "mul.rn.f64 x, p, %7;\n\t"
      "@t0 add.rn.f64 %0, %0, x;\n\t"
      "mul.rn.f64 x, p, %8;\n\t"
      "@t1 add.rn.f64 %0, %0, x;\n\t"
      "mul.rn.f64 x, p, %9;\n\t"
      "@t2 add.rn.f64 %0, %0, x;\n\t"
      "mul.rn.f64 x, p, %10;\n\t"
      "@t3 add.rn.f64 %0, %0, x;\n\t"
      "}"
      : "=&d"(acc)
      : "d"(rs), "d"(ra[0]), "d"(ra[1]), "d"(ra[2]), "d"(ra[3]), "d"(m), "d"(ga[0]), "d"(ga[1]), "d"(ga[2]),
        "d"(ga[3]), "r"(lut));
#endif
  return acc;
}

// ---- f32x2 formulation (sm_100a FMUL2 / FADD2 / FFMA2) ------------------------------------------------
// The fused-greedy sweep is bound by instruction issue and by the half-rate ALU pipe (selects,
// compares, min/max), not by HBM.  Blackwell's packed f32x2 instructions do two IEEE fp32 operations
// per issue slot (measured on B200, tools/ubench/pipes.cu: 2 cycles per warp instruction, and they
// co-issue with ALU-pipe instructions), so every multiply / add of the backup is done for two
// neighbouring cells at once -- same operations, same order, same roundings per lane as the scalar
// form; the freed issue slots go to the ALU pipe.
#ifndef GU_F32_PACK
#define GU_F32_PACK 1
#endif
#ifndef GU_F32_PACK_FOLD
#define GU_F32_PACK_FOLD 1      // acc = fma(w, p*g, acc): exact for w in {0, 1}, one packed op instead of two
#endif
#ifndef GU_F32_DENORM_IDX
#define GU_F32_DENORM_IDX 1     // tie count + table offset accumulated as a denormal (its bits ARE the byte offset)
#endif
#ifndef GU_F32_PACK_W
#define GU_F32_PACK_W 0         // 1: tie weights on the FMA pipe (max(ra - m + 1, 0)) instead of FSET
#endif

__device__ __forceinline__ float lds_f32(const void* base, uint32_t off) {
  return *reinterpret_cast<const float*>(static_cast<const char*>(base) + off);
}

// Backup of two neighbouring cells.  ra*/ga*: rounded scaled q-value / discounted value of the
// landing cell per action; m*: max of ra*; off*: (goal | lava << 1) << 3 of the cell.
// Terminal cells: all four actions are blocked, so all tie and the count is 4; the table entry then
// holds p = 0 and the sum stays R[s] + (+-0) = R[s] (utils.py:70).
__device__ __forceinline__ void backup_ties_pair(const float (&ra0)[4], const float (&ra1)[4], float m0, float m1,
                                                 const float (&ga0)[4], const float (&ga1)[4], uint32_t off0,
                                                 uint32_t off1, const Luts<float>& l, float& out0, float& out1) {
  float2 w[4];
#if GU_F32_PACK_W
  // ra, m are integer-valued (rint), ra <= m: ra - m is exact 0 for a tie and <= -1 otherwise
  const float2 nm = make_float2(-m0, -m1), one = make_float2(1.0f, 1.0f);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float2 u = __fadd2_rn(__fadd2_rn(make_float2(ra0[a], ra1[a]), nm), one);
    w[a] = make_float2(fmaxf(u.x, 0.0f), fmaxf(u.y, 0.0f));
  }
#else
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    asm("set.eq.f32.f32 %0, %1, %2;" : "=f"(w[a].x) : "f"(ra0[a]), "f"(m0));
    asm("set.eq.f32.f32 %0, %1, %2;" : "=f"(w[a].y) : "f"(ra1[a]), "f"(m1));
  }
#endif
#if GU_F32_DENORM_IDX
  // The table offset is accumulated as a DENORMAL: a denormal's bit pattern is its value in units of
  // 2^-149, so starting from the cell's goal / lava offset and adding 32 * 2^-149 per tie leaves the byte
  // offset of the {1/len(ties), R} entry in the register as it is -- no mask, no OR (sums of denormals
  // below 2^-126 are exact, and FFMA2 handles denormals at full rate without .ftz).
  float2 c = make_float2(__uint_as_float(off0), __uint_as_float(off1));
  const float2 k32 = make_float2(__uint_as_float(32u), __uint_as_float(32u));
#pragma unroll
  for (int a = 0; a < 4; ++a) c = __ffma2_rn(w[a], k32, c);
  const uint32_t i0 = __float_as_uint(c.x), i1 = __float_as_uint(c.y);
#else
  // count of ties in mantissa bits 5-7 of a magic-number sum
  float2 c = make_float2(8388608.0f, 8388608.0f);
  const float2 k32 = make_float2(32.0f, 32.0f);
#pragma unroll
  for (int a = 0; a < 4; ++a) c = __ffma2_rn(w[a], k32, c);
  const uint32_t i0 = (__float_as_uint(c.x) & 0xe0u) | off0, i1 = (__float_as_uint(c.y) & 0xe0u) | off1;
#endif
  const float2 p = make_float2(lds_f32(&l.prs[0][0], i0), lds_f32(&l.prs[0][0], i1));
  float2 acc = make_float2(lds_f32(&l.prs[0][1], i0), lds_f32(&l.prs[0][1], i1));      // R[s]
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float2 pg = __fmul2_rn(p, make_float2(ga0[a], ga1[a]));
#if GU_F32_PACK_FOLD
    acc = __ffma2_rn(w[a], pg, acc);
#else
    acc = __fadd2_rn(acc, __fmul2_rn(w[a], pg));
#endif
  }
  out0 = acc.x;
  out1 = acc.y;
}

// Greedy extraction: one tie of action a in cell j of a cell pair adds k = 2^(a + 8 j) to a magic-number
// accumulator.  f32: FSET (1.0 / 0.0) feeding an FFMA; f64: the compare predicates an FADD.
#ifndef GU_TIE_FMA
#define GU_TIE_FMA 1
#endif
__device__ __forceinline__ void tie_accumulate(float ra, float m, float k, float& acc) {
  float w;
  asm("set.eq.f32.f32 %0, %1, %2;" : "=f"(w) : "f"(ra), "f"(m));
  acc = __fmaf_rn(w, k, acc);
}
__device__ __forceinline__ void tie_accumulate(double ra, double m, float k, float& acc) {
  asm("{\n\t"
      ".reg .pred t;\n\t"
      "setp.eq.f64 t, %1, %2;\n\t"
      "@t add.rn.f32 %0, %0, %3;\n\t"
      "}"
      : "+f"(acc)
      : "d"(ra), "d"(m), "f"(k));
}

#ifndef GU_TILED_WARPS
#define GU_TILED_WARPS 4
#endif
#ifndef GU_TILED_PREFETCH_ROWS
#define GU_TILED_PREFETCH_ROWS 3
#endif
constexpr int kTiledWarps = GU_TILED_WARPS;

// max without NaN semantics: one compare + select (fmax(double) costs ~7 instructions on sm_100)
__device__ __forceinline__ float max_nn(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double max_nn(double a, double b) { return a > b ? a : b; }

// One row of the sliding window.  v / info / hv / hinfo are loaded one row of compute ahead
// ("raw" part); convert() derives the discounted values g and the rounded scaled q-values rt
// of the thread's own columns plus the left / right neighbour columns.
template <typename T, int CPT, bool TIES>
struct WinRow {
  T v[CPT], hv;
  uint32_t info[CPT / 4 ? CPT / 4 : 1], hinfo;   // CPT info bytes, little-endian
  T g[CPT], gl, gr;
  T rt[TIES ? CPT : 1], rtl, rtr;
};

__device__ __forceinline__ uint32_t info_of(const uint32_t* info, int j) {
  return (info[j >> 2] >> (8 * (j & 3))) & 0xffu;
}

// Byte offset of cell j's {reward, poison} entry: the goal / lava bits of all four cells of an
// info word are masked at once and one PRMT per cell moves the byte into place (instead of a
// shift and a mask per cell).  f32 entries are 8 bytes (bits 3-4 as they are), f64 entries 16.
#ifndef GU_LUT_PRMT
#define GU_LUT_PRMT 1
#endif
template <typename T>
__device__ __forceinline__ uint32_t lut_offset(const uint32_t* info, int j) {
#if GU_LUT_PRMT
  const uint32_t m = sizeof(T) == 4 ? (info[j >> 2] & 0x18181818u) : ((info[j >> 2] << 1) & 0x30303030u);
  return (j & 3) == 0 ? (m & 0xffu) : __byte_perm(m, 0u, 0x4440u + (j & 3));
#else
  const uint32_t inf = (info[j >> 2] >> (8 * (j & 3))) & 0xffu;
  return sizeof(T) == 4 ? (inf & 0x18u) : ((inf & 0x18u) << 1);
#endif
}
template <typename T>
__device__ __forceinline__ T reward_at(const Luts<T>& l, uint32_t off) {
  return *reinterpret_cast<const T*>(reinterpret_cast<const char*>(&l.reward[0][0]) + off);
}
__device__ __forceinline__ float2 reward_poison_at(const Luts<float>& l, uint32_t off) {
  return *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(&l.reward[0][0]) + off);
}
__device__ __forceinline__ double2 reward_poison_at(const Luts<double>& l, uint32_t off) {
  return *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(&l.reward[0][0]) + off);
}

// rint(t) for t = q * 1e8.  f32: |t| >= 2^23 is already integral (the common case: |q| > 0.084).
// f64: the magic-number add is exact below 2^51.  `need_slow` collects the rare other cases.
#ifndef GU_F32_FRND
#define GU_F32_FRND 1
#endif
__device__ __forceinline__ float round_fast(float t, float& worst) {
#if GU_F32_FRND
  return rintf(t);                          // FRND on the otherwise idle XU pipe; no slow path needed
#else
  worst = fminf(worst, fabsf(t));           // slow path if any |t| < 2^23
  return t;
#endif
}
#ifndef GU_F64_FRND
#define GU_F64_FRND 1
#endif
__device__ __forceinline__ double round_fast(double t, double& worst) {
#if GU_F64_FRND
  return rint(t);                           // FRND.F64
#else
  worst = max_nn(worst, fabs(t));           // slow path if any |t| >= 2^51
  return __dadd_rn(__dadd_rn(t, 6755399441055744.0), -6755399441055744.0);
#endif
}
__device__ __forceinline__ bool round_needs_slow(float worst) { return GU_F32_FRND ? false : worst < 8388608.0f; }
__device__ __forceinline__ bool round_needs_slow(double worst) { return GU_F64_FRND ? false : !(worst < 2251799813685248.0); }
__device__ __forceinline__ float round_worst_init(float) { return CUDART_INF_F; }
__device__ __forceinline__ double round_worst_init(double) { return 0.0; }

// ---- NVLink peer-memory links of a row shard (fused halo exchange + residual all-reduce) --------
// With PEER the sweep kernel replaces both collectives of the row-sharded value iteration
// (protocol: include/gu_b200.h, gu_peer_links):
//  * the blocks that own the shard's first / last row also store their output vectors straight
//    into the ghost row of the neighbour's v_out (peer memory over NVLink); the last of them to
//    finish raises the neighbour's halo flag to slot + 1.  Only the edge blocks of the next sweep
//    wait, and only for their own neighbour's flag; they are scheduled first.
//  * the last block of the shard publishes the shard's residual into slot `slot` of every rank's
//    residual table; sweep `slot` gates on the complete row slot - lag (lag 2: the row has normally
//    arrived long ago, so no rank stalls on the slowest one) and on the sticky stop word.
//  * a wait that times out raises the abort word of every rank; aborted ranks stop writing.
template <typename T>
struct PeerArgs {
  int rank, world, slot, lag, first_slot;
  T* up_ghost;                 // neighbour above: bottom ghost row of its v_out, or nullptr
  T* down_ghost;               // neighbour below: top ghost row of its v_out, or nullptr
  T* tables[GU_MAX_PEERS];     // rank r's residual table T[slots][world]; tables[rank] is local
  int* aborts[GU_MAX_PEERS];   // rank r's abort word; aborts[rank] is local
  int* done;                   // local block counter, zero between launches
  int* err;                    // local error flag (a wait of this rank timed out)
  int* halo;                   // local [2]: sweeps delivered from above / below
  int* up_flag;                // neighbour above: its halo[1]
  int* down_flag;              // neighbour below: its halo[0]
  int* edge;                   // local [2]: finished blocks of the first / last block row
  int* stop;                   // local sticky "converged" word
  const int* slot_base;
  long long timeout;
  T thr;
};

constexpr long long kPeerTimeoutCycles = 6000000000ll;   // ~3 s: give up instead of hanging the GPU

template <typename T>
__device__ __noinline__ void peer_abort(const PeerArgs<T>& p) {
  *reinterpret_cast<volatile int*>(p.err) = 1;
  for (int r = 0; r < p.world; ++r) *reinterpret_cast<volatile int*>(p.aborts[r]) = 1;
  __threadfence_system();
}

template <typename T>
__device__ __forceinline__ bool peer_aborted(const PeerArgs<T>& p) {
  return *reinterpret_cast<const volatile int*>(p.aborts[p.rank]) != 0;
}

template <typename T>
__device__ __forceinline__ void peer_publish(const PeerArgs<T>& p, int slot, T val) {
  __threadfence_system();
  for (int r = 0; r < p.world; ++r)
    *reinterpret_cast<volatile T*>(p.tables[r] + static_cast<size_t>(slot) * p.world + p.rank) = val;
  __threadfence_system();
}

// Outcome of a wait: the value / flag arrived, the solve was aborted (this rank's wait timed out or
// another rank raised the abort word), or the sticky stop word is set.  After convergence the rows
// and flags of the gated-off sweeps never arrive, so every spin also polls the stop word.
enum { kWaitOk = 0, kWaitAborted = 1, kWaitStopped = 2 };

template <bool POLL_STOP, typename T, typename Cond>
__device__ __forceinline__ int peer_spin(const PeerArgs<T>& p, Cond arrived) {
  if (arrived()) return kWaitOk;
  const long long t0 = clock64();
  do {
    if (POLL_STOP && *reinterpret_cast<const volatile int*>(p.stop) != 0) return kWaitStopped;
    if (peer_aborted(p)) return kWaitAborted;
    if (clock64() - t0 > p.timeout) { peer_abort(p); return kWaitAborted; }
  } while (!arrived());
  return kWaitOk;
}

// one entry of slot `slot` of the local table (spins until rank r has reported; NaN = not yet)
template <bool POLL_STOP = true, typename T>
__device__ __forceinline__ int peer_wait_entry(const PeerArgs<T>& p, int slot, int r, T& out) {
  const volatile T* e = p.tables[p.rank] + static_cast<size_t>(slot) * p.world + r;
  T v = Num<T>::neg_inf();
  const int rc = peer_spin<POLL_STOP>(p, [&]() { v = *e; return v == v; });
  out = rc == kWaitOk ? v : Num<T>::neg_inf();
  return rc;
}

// spin until *word >= target (a neighbour's halo flag)
template <bool POLL_STOP = true, typename T>
__device__ __forceinline__ int peer_wait_flag(const PeerArgs<T>& p, const int* word, int target) {
  const volatile int* w = word;
  return peer_spin<POLL_STOP>(p, [&]() { return *w >= target; });
}

// all entries of a slot (the host-side wait kernel, which waits whatever the stop word says): max
// over ranks, or false if aborted
template <typename T>
__device__ __forceinline__ bool peer_wait_slot(const PeerArgs<T>& p, int slot, T& out) {
  T m = Num<T>::neg_inf();
  for (int r = 0; r < p.world; ++r) {
    T v;
    if (peer_wait_entry<false>(p, slot, r, v) != kWaitOk) return false;
    m = v > m ? v : m;
  }
  out = m;
  return true;
}

template <typename T>
__global__ void peer_wait_kernel(const __grid_constant__ PeerArgs<T> p) {
  const int slot = p.slot + (p.slot_base ? *p.slot_base : 0);
  T m;
  if (!peer_wait_slot(p, slot, m)) return;
  if (p.up_flag && peer_wait_flag<false>(p, p.halo + 0, slot + 1) != kWaitOk) return;
  if (p.down_flag && peer_wait_flag<false>(p, p.halo + 1, slot + 1) != kWaitOk) return;
  __threadfence_system();
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// Resident blocks per SM the register allocation is held to.  The fp32 fused-greedy sweep needs ~160
// registers (three window rows of v / g / rt for 8 cells + the packed backup of a cell pair): 3 blocks
// x 4 warps without spills measured 0.480 ms at 16384^2 against 0.550 ms for 4 blocks with spills; the
// other fp32 kinds fit 128 registers (4 blocks); fp64 is left to the compiler (2 blocks).
#ifndef GU_TILED_MIN_BLOCKS
#define GU_TILED_MIN_BLOCKS 3
#endif
#ifndef GU_TILED_MIN_BLOCKS_F64
#define GU_TILED_MIN_BLOCKS_F64 3
#endif
template <typename T, int KIND, bool WRITE_TIE>
constexpr int tiled_min_blocks() {
  return sizeof(T) != 4 ? GU_TILED_MIN_BLOCKS_F64 : ((KIND == GU_POLICY_GREEDY && !WRITE_TIE) ? GU_TILED_MIN_BLOCKS : 4);
}
// ---- TMA-staged window (TMA = true) ------------------------------------------------------------------
// Each warp's strip of the value grid and of the info plane arrives as 2-D TMA tiles
// (cp.async.bulk.tensor.2d, SASS UTMALDG) of kTmaRows rows -- the strip plus one 16-byte unit of halo on
// either side, out-of-range columns zero-filled by the hardware -- through a kTmaStages-deep ring in
// shared memory, completed on per-warp mbarriers; no block-level barrier in the row loop.  The window
// rows are then read with shared loads, so there are no global loads, no L2 prefetches and no halo-column
// special case in the loop, and several KB per warp are always in flight.
// Tensor maps: the value array is described as 8-byte elements (two f32 / one f64) and the info plane as
// 4-byte elements, which keeps the 264-column boxes under the 256-element box limit.
#ifndef GU_TMA_ROWS
#define GU_TMA_ROWS 4
#endif
#ifndef GU_TMA_STAGES
#define GU_TMA_STAGES 3
#endif
constexpr int kTmaRows = GU_TMA_ROWS, kTmaStages = GU_TMA_STAGES;
template <typename T>
struct TmaGeom {
  static constexpr int CPT = Vec<T>::W * 2;                                  // cells per thread (NV = 2)
  static constexpr int kPadV = 16 / static_cast<int>(sizeof(T));             // halo unit of V, elements
  static constexpr int kVRowBytes = (32 * CPT + 2 * kPadV) * static_cast<int>(sizeof(T));
  static constexpr int kIRowBytes = 32 * CPT + 32;                           // info: 16 bytes of halo each side
  static constexpr int kStageBytes = kTmaRows * (kVRowBytes + kIRowBytes);
  static constexpr int kWarpBytes = kTmaStages * kStageBytes;
  static_assert((kTmaRows * kVRowBytes) % 128 == 0 && kStageBytes % 128 == 0, "TMA destinations are 128-byte aligned");
};

__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}

template <typename T, int KIND, bool WRITE_TIE, int NV, bool PEER = false, bool TMA = false>
__global__ void __launch_bounds__(kTiledWarps * 32, (tiled_min_blocks<T, KIND, WRITE_TIE>()))
sweep_tiled_kernel(GridView g, const uint8_t* __restrict__ info, const T* __restrict__ vin,
                   T* __restrict__ vout, uint8_t* __restrict__ tie_out, const void* __restrict__ policy,
                   T gamma, T* residual, const T* gate, T gate_thr, int rows_per_block,
                   const __grid_constant__ PeerArgs<T> peer,     // read in place from the constant bank: the
                                                                 // helpers take it by reference, and a by-value
                                                                 // copy would be spilled to local memory by every thread
                   const __grid_constant__ CUtensorMap vmap, const __grid_constant__ CUtensorMap imap) {
  using N = Num<T>;
  using V = typename Vec<T>::type;
  constexpr int W = Vec<T>::W;
  constexpr int CPT = W * NV;                 // cells per thread: NV 16-byte vectors
  constexpr int IW = CPT / 4 ? CPT / 4 : 1;   // 32-bit words of info per thread-row
  constexpr bool TIES = (KIND == GU_POLICY_GREEDY) || WRITE_TIE;
  // f32x2 arithmetic for pairs of neighbouring cells (fp32 only; rint via FRND, no slow path)
  constexpr bool kPack = GU_F32_PACK && GU_F32_FRND && sizeof(T) == 4 && CPT % 2 == 0;
  __shared__ T scratch[kTiledWarps];
  __shared__ __align__(16) Luts<T> luts;
  // GU_POLICY_PROBS: a thread's CPT cells own CPT*4 probabilities = UPT 16-byte units that sit 128
  // bytes apart from the next lane's, so a direct load touches 32 lines per request.  Instead the
  // warp copies its strip of the [cells][4] row with coalesced 16-byte cp.async requests (unit
  // k*32+lane per request) into a double-buffered shared stage, padded by one unit per thread chunk
  // so the per-thread reads are bank-conflict free, one row ahead of the compute.
  constexpr int UPT = CPT * static_cast<int>(sizeof(T)) / 4;
  constexpr int kStage = 32 * UPT + 32;
  __shared__ __align__(16) uint4 pstage[KIND == GU_POLICY_PROBS ? kTiledWarps * 2 * kStage : 1];
  // Programmatic dependent launch: this grid may have been started while the previous kernel of the
  // stream (the sweep before) was still draining -- its blocks take the SM slots that free up, set up
  // their tables and wait here; nothing the previous kernel wrote is read above this line, and every
  // exit path is below it.  In turn the next launch is released as soon as all blocks of this grid
  // are resident.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  init_luts(luts);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int by = blockIdx.y;
  int slot = 0;
  bool top_edge = false, bot_edge = false;
  if (!PEER) {
    if (gate != nullptr && *gate < gate_thr) return;
  } else {
    const int nby = gridDim.y;
    // (Scheduling the last block row second, so that its rows travel while the interior computes, was
    // measured at +23 % kernel time on one GPU -- 0.320 against 0.260 ms at 8192 rows -- for reasons the
    // instruction mix does not explain; with the natural order the only exposed latency is one NVLink
    // store + flag at the start of the neighbour's next sweep, in its first 16 blocks.)
#ifdef GU_PEER_REMAP
    if (nby >= 3) by = blockIdx.y == 0 ? 0 : (blockIdx.y == 1 ? nby - 1 : static_cast<int>(blockIdx.y) - 1);
#endif
    top_edge = by == 0 && peer.up_flag != nullptr;
    bot_edge = by == nby - 1 && peer.down_flag != nullptr;
    slot = peer.slot + (peer.slot_base != nullptr ? *peer.slot_base : 0);
    __shared__ int go_sh;
    if (threadIdx.x < 32) {
      // one warp, one round trip in the common case: lane r < world reads rank r's entry of the gate
      // row, lanes 31 / 30 the sticky stop and the abort word, lanes 29 / 28 the halo flags
      const int lane0 = threadIdx.x;
      const int gslot = slot - peer.lag;
      int bad = 0;                                   // non-zero: stop (converged earlier / aborted)
      T v = N::neg_inf();
      if (lane0 == 31) bad = *reinterpret_cast<const volatile int*>(peer.stop) != 0;
      if (lane0 == 30) bad = peer_aborted(peer);
      if (gslot >= peer.first_slot && lane0 < peer.world) bad = peer_wait_entry(peer, gslot, lane0, v);
      if (slot > peer.first_slot) {
        if (lane0 == 29 && top_edge) bad = peer_wait_flag(peer, peer.halo + 0, slot);
        if (lane0 == 28 && bot_edge) bad = peer_wait_flag(peer, peer.halo + 1, slot);
      }
      const bool any_bad = __any_sync(0xffffffffu, bad != 0);
      const T m = warp_max(v);
      int go = 1;
      if (any_bad) go = 0;
      else if (gslot >= peer.first_slot && m < peer.thr) {     // converged earlier: keep V
        if (lane0 == 0) *reinterpret_cast<volatile int*>(peer.stop) = 1;
        go = 0;
      }
      if (go && (top_edge || bot_edge)) __threadfence_system();     // acquire the neighbour's rows
      if (lane0 == 0) go_sh = go;
    }
    __syncthreads();
    if (!go_sh) return;
  }
  __syncthreads();                                        // luts

  const int lane = threadIdx.x & 31;
  const int x0 = ((blockIdx.x * kTiledWarps + (threadIdx.x >> 5)) * 32 + lane) * CPT;
  const int rows = g.row_end - g.row_begin;
  const int ry0 = by * rows_per_block;
  const int ry1 = min(ry0 + rows_per_block, rows);
  const bool active = x0 < g.X;                       // lanes past the grid still take part in shuffles
  const bool has_l = active && lane == 0 && x0 > 0;   // edge lanes fetch one halo column each
  const bool has_r = active && lane == 31 && x0 + CPT < g.X;
  const bool full = x0 + CPT <= g.X;
  const int pitch = g.pitch;                         // element offsets fit in 31 bits (tiled_ok)
  const int hoff = has_l ? -1 : CPT;                  // halo column relative to x0

  WinRow<T, CPT, TIES> w[3];

  // L2 prefetch of the warp's strip kPrefetchRows rows ahead of the demand loads: one 128-byte
  // line per lane (V lines first, then the info lines), so the demand loads hit L2.
  constexpr int kPrefetchRows = GU_TILED_PREFETCH_ROWS;
  constexpr int kVLines = 32 * CPT * static_cast<int>(sizeof(T)) / 128;
  constexpr int kILines = (32 * CPT + 127) / 128;
  const int wx0 = x0 - lane * CPT;                    // first column of the warp's strip
  const char* pf_base = nullptr;
  if (lane < kVLines) pf_base = reinterpret_cast<const char*>(vin + wx0) + lane * 128;
  else if (lane < kVLines + kILines) pf_base = reinterpret_cast<const char*>(info + wx0) + (lane - kVLines) * 128;
  else if (KIND == GU_POLICY_MASK && !WRITE_TIE && lane < kVLines + 2 * kILines)      // the tie-mask plane, same geometry as info
    pf_base = static_cast<const char*>(policy) + wx0 + (lane - kVLines - kILines) * 128;
  const int pf_pitch = lane < kVLines ? pitch * static_cast<int>(sizeof(T)) : pitch;   // bytes; prefetch only
  const int pf_col = lane < kVLines ? lane * (128 / static_cast<int>(sizeof(T)))
                                    : (lane < kVLines + kILines ? lane - kVLines : lane - kVLines - kILines) * 128;
  const bool pf_on = pf_base != nullptr && wx0 + pf_col < g.pitch;

  // ---- TMA ring (TMA = true): per-warp state -------------------------------------------------------
  extern __shared__ __align__(1024) uint8_t ring_raw[];
  __shared__ __align__(8) uint64_t ring_bars[TMA ? kTiledWarps : 1][kTmaStages];
  using TG = TmaGeom<T>;
  const int warp_u = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);            // warp-uniform for the TMA operands
  const uint32_t ring_s = TMA ? ((smem_u32(ring_raw) + 127u) & ~127u) + warp_u * TG::kWarpBytes : 0u;
  const uint32_t bar_s = smem_u32(&ring_bars[TMA ? warp_u : 0][0]);
  int ld_slot = 0, ld_row = 0, ld_box = 0;      // next row to read: ring slot, row inside its box, box number
  int cv_row = 0, cv_box = 0;                   // next row to be converted (its box is released after its last row)
  const int nboxes = TMA ? (ry1 - ry0 + 3 + kTmaRows - 1) / kTmaRows : 0;      // array rows ry0 .. ry1 + 2
  auto ring_issue = [&](int box) {              // one elected lane: V tile + info tile of box `box`
    const int slot = box % kTmaStages;
    const uint32_t dst = ring_s + slot * TG::kStageBytes, bar = bar_s + 8 * slot;
    mbar_expect_tx(bar, TG::kStageBytes);
    // x in map elements: V is mapped as 8-byte elements, the info plane as 4-byte elements
    tma_tile_2d(dst, &vmap, (wx0 - TG::kPadV) * static_cast<int>(sizeof(T)) / 8, ry0 + box * kTmaRows, bar);
    tma_tile_2d(dst + kTmaRows * TG::kVRowBytes, &imap, (wx0 - 16) / 4, ry0 + box * kTmaRows, bar);
  };
  if constexpr (TMA) {
    if (elect_one()) {
      for (int i = 0; i < kTmaStages; ++i) mbar_init(bar_s + 8 * i, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (elect_one())
      for (int b = 0; b < kTmaStages && b < nboxes; ++b) ring_issue(b);
  }
  // the row just converted was read from the ring a row earlier: when it was the last row of its box,
  // every lane is done with that box and its slot takes the box kTmaStages further on
  auto ring_release = [&]() {
    if constexpr (TMA) {
      if (++cv_row == kTmaRows) {
        cv_row = 0;
        __syncwarp();
        if (cv_box + kTmaStages < nboxes && elect_one()) ring_issue(cv_box + kTmaStages);
        ++cv_box;
      }
    }
  };

  auto issue_loads = [&](int ar, WinRow<T, CPT, TIES>& r) {
    if constexpr (TMA) {
      // rows are read strictly in order; `ar` is only what the LDG path needs
      if (ld_row == 0) mbar_wait(bar_s + 8 * ld_slot, static_cast<uint32_t>(ld_box / kTmaStages) & 1u);
      const uint32_t vrow = ring_s + ld_slot * TG::kStageBytes + ld_row * TG::kVRowBytes;
      const uint32_t irow = ring_s + ld_slot * TG::kStageBytes + kTmaRows * TG::kVRowBytes + ld_row * TG::kIRowBytes;
      const uint32_t vme = vrow + (TG::kPadV + lane * CPT) * static_cast<int>(sizeof(T));
      const uint32_t ime = irow + 16 + lane * CPT;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        V vec;
        if constexpr (sizeof(T) == 4)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vec.x), "=f"(vec.y), "=f"(vec.z), "=f"(vec.w) : "r"(vme + 16 * k));
        else
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(vec.x), "=d"(vec.y) : "r"(vme + 16 * k));
        unpack(vec, r.v + k * W);
      }
      if constexpr (CPT == 8) {
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.info[0]), "=r"(r.info[1]) : "r"(ime));
      } else {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r.info[0]) : "r"(ime));
      }
      // halo columns: the unit left of the strip ends with column wx0 - 1, the unit right of it starts
      // with column wx0 + 32 * CPT (zero-filled outside the array; outside the GRID they are not used)
      r.hv = T(0);
      r.hinfo = 0;
      if (has_l || has_r) {
        const uint32_t hva = has_l ? vrow + (TG::kPadV - 1) * static_cast<int>(sizeof(T)) : vrow + (TG::kPadV + 32 * CPT) * static_cast<int>(sizeof(T));
        const uint32_t hia = has_l ? irow + 15 : irow + 16 + 32 * CPT;
        if constexpr (sizeof(T) == 4) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r.hv) : "r"(hva));
        else asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r.hv) : "r"(hva));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r.hinfo) : "r"(hia));
      }
      if (!active) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) r.v[j] = T(0);
#pragma unroll
        for (int k = 0; k < IW; ++k) r.info[k] = 0;
      }
      if (++ld_row == kTmaRows) {
        ld_row = 0;
        ++ld_box;
        if (++ld_slot == kTmaStages) ld_slot = 0;
      }
      return;
    }
    const int o = ar * pitch + x0;
    if (pf_on) {
      const int par = min(ar + kPrefetchRows, rows + 1);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_base + static_cast<size_t>(static_cast<unsigned>(par) * static_cast<unsigned>(pf_pitch))));
    }
    r.hv = T(0);
    r.hinfo = 0;
    if (active) {
#pragma unroll
      for (int k = 0; k < NV; ++k) unpack(*reinterpret_cast<const V*>(vin + o + k * W), r.v + k * W);
      if (CPT == 2) r.info[0] = *reinterpret_cast<const uint16_t*>(info + o);
      else {
#pragma unroll
        for (int k = 0; k < IW; ++k) r.info[k] = *reinterpret_cast<const uint32_t*>(info + o + 4 * k);
      }
      if (has_l || has_r) { r.hv = vin[o + hoff]; r.hinfo = info[o + hoff]; }
    } else {
#pragma unroll
      for (int j = 0; j < CPT; ++j) r.v[j] = T(0);
#pragma unroll
      for (int k = 0; k < IW; ++k) r.info[k] = 0;
    }
  };

  auto convert = [&](WinRow<T, CPT, TIES>& r) {
    T worst = round_worst_init(T(0));
    T ht = T(0);
    if constexpr (kPack) {
#pragma unroll
      for (int j = 0; j < CPT; j += 2) {
        // ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 when the product has no other use, even
        // with --fmad=false (and it rewrites fma(g, 1, R) to that add first).  In the sweep kernels g
        // lives on in the window, so the packed pair stays unfused (checked in SASS: FMUL2 + FADD2); the
        // greedy-extraction kernel has no other use for g and takes the scalar .rn forms, which are
        // never contracted.
        float2 g2, s2;
        const float2 rw = make_float2(TIES ? reward_at(luts, lut_offset<T>(r.info, j)) : 0.0f,
                                      TIES ? reward_at(luts, lut_offset<T>(r.info, j + 1)) : 0.0f);
        if constexpr (WRITE_TIE) {
          g2 = make_float2(__fmul_rn(gamma, r.v[j]), __fmul_rn(gamma, r.v[j + 1]));
          s2 = make_float2(__fadd_rn(rw.x, g2.x), __fadd_rn(rw.y, g2.y));
        } else {
          g2 = __fmul2_rn(make_float2(gamma, gamma), make_float2(r.v[j], r.v[j + 1]));
          if constexpr (TIES) s2 = __fadd2_rn(rw, g2);
        }
        r.g[j] = g2.x;
        r.g[j + 1] = g2.y;
        if constexpr (TIES) {
          const float2 t2 = __fmul2_rn(s2, make_float2(1e8f, 1e8f));
          r.rt[j] = rintf(t2.x);
          r.rt[j + 1] = rintf(t2.y);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        r.g[j] = N::mul(gamma, r.v[j]);
        if constexpr (TIES)
          r.rt[j] = round_fast(N::mul(N::add(reward_at(luts, lut_offset<T>(r.info, j)), r.g[j]), N::scale()), worst);
      }
    }
    const T hg = N::mul(gamma, r.hv);
    if constexpr (TIES) {
      ht = round_fast(N::mul(N::add(reward_lut(luts, r.hinfo), hg), N::scale()), worst);
      if (round_needs_slow(worst)) {   // rare: re-round everything with rint() (idempotent on integers)
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          r.rt[j] = N::rnd(N::mul(N::add(reward_at(luts, lut_offset<T>(r.info, j)), r.g[j]), N::scale()));
        ht = N::rnd(N::mul(N::add(reward_lut(luts, r.hinfo), hg), N::scale()));
      }
    }
    r.gl = shfl_up1(r.g[CPT - 1]);
    r.gr = shfl_down1(r.g[0]);
    if (lane == 0) r.gl = hg;
    if (lane == 31) r.gr = hg;
    if constexpr (TIES) {
      r.rtl = shfl_up1(r.rt[CPT - 1]);
      r.rtr = shfl_down1(r.rt[0]);
      if (lane == 0) r.rtl = ht;
      if (lane == 31) r.rtr = ht;
    }
  };

  auto issue_policy = [&](int ry, int buf) {
    if constexpr (KIND == GU_POLICY_PROBS) {
      if (ry < ry1) {
        const char* src = static_cast<const char*>(policy) +
                          (static_cast<size_t>(ry + 1) * pitch + wx0) * (4 * sizeof(T));
        const uint32_t dst = smem_u32(&pstage[((threadIdx.x >> 5) * 2 + buf) * kStage]);
        const int strip_units = min(g.pitch - wx0, 32 * CPT) * static_cast<int>(sizeof(T)) / 4;
#pragma unroll
        for (int k = 0; k < UPT; ++k) {
          const int u = k * 32 + lane;
          if (u < strip_units) cp_async16(dst + (u + u / UPT) * 16, src + static_cast<size_t>(u) * 16);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };
  issue_policy(ry0, 0);

  T dmax = N::neg_inf();
  const int last_ar = rows + 1;                 // bottom ghost row of the shard's arrays
  // GU_POLICY_MASK: the row's tie masks are loaded one row of compute ahead, like the window (loaded at
  // the point of use they were the kernel's exposed latency: long-scoreboard stalls of 5 per issue)
  uint32_t pm_next[IW];
#pragma unroll
  for (int k = 0; k < IW; ++k) pm_next[k] = 0;
  auto load_masks = [&](int ar) {               // masks of array row `ar` (an owned row)
    if constexpr (KIND == GU_POLICY_MASK && !WRITE_TIE) {
      if (active) {
        const uint8_t* pp = static_cast<const uint8_t*>(policy) + ar * pitch + x0;
        if (CPT == 2) pm_next[0] = *reinterpret_cast<const uint16_t*>(pp);
        else {
#pragma unroll
          for (int k = 0; k < IW; ++k) pm_next[k] = *reinterpret_cast<const uint32_t*>(pp + 4 * k);
        }
      }
    }
  };
  load_masks(ry0 + 1);
  issue_loads(ry0, w[0]);                       // array row ry0     = row above the first owned row
  issue_loads(ry0 + 1, w[1]);                   // array row ry0 + 1 = first owned row
  issue_loads(ry0 + 2, w[2]);
  convert(w[0]);
  ring_release();
  convert(w[1]);
  ring_release();

  for (int base = ry0; base < ry1; base += 3) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int ry = base + j;
      if (ry < ry1) {
        WinRow<T, CPT, TIES>& up = w[j % 3];
        WinRow<T, CPT, TIES>& cur = w[(j + 1) % 3];
        WinRow<T, CPT, TIES>& dn = w[(j + 2) % 3];
        convert(dn);                               // row ry + 2, loaded during the previous row
        ring_release();
        // the row after that goes into the raw fields of `up` (its v / info are dead, its g / rt
        // stay valid for this row's compute); the loads are in flight while this row is computed
        issue_loads(min(ry + 3, last_ar), up);
        if constexpr (KIND == GU_POLICY_PROBS) {
          __syncwarp();                                  // everyone is done reading the other buffer
          issue_policy(ry + 1, (ry + 1 - ry0) & 1);
          asm volatile("cp.async.wait_group 1;" ::: "memory");
          __syncwarp();                                  // this row's units from all lanes have landed
        }
        if (active) {
          const int o = (ry + 1) * pitch + x0;
          T out[CPT];
          uint32_t ties[IW];
          uint32_t pm[IW];
#pragma unroll
          for (int k = 0; k < IW; ++k) { ties[k] = 0; pm[k] = pm_next[k]; }
          if (ry + 1 < ry1) load_masks(ry + 2);          // next row's masks: in flight while this row is computed
          if constexpr (WRITE_TIE && GU_TIE_FMA) {
            // Greedy extraction: the kernel is bound by the half-rate ALU pipe (selects, compares, bit
            // assembly), so the tie mask is assembled on the FMA pipe instead: every tie adds 2^(a + 8 j)
            // to a magic-number accumulator (2^23: the integer lands in the low mantissa bits), two cells
            // per accumulator, and one PRMT puts the four masks of an info word together.  Terminal
            // cells (all four actions blocked, so all four tie) are cleared with word-wide bit
            // arithmetic on the goal / lava bits instead of a NaN "poison" added to every max.
#pragma unroll
            for (int k = 0; k < IW; ++k) {
              float acc2[2] = {8388608.0f, 8388608.0f};
#pragma unroll
              for (int j = 0; j < 4 && 4 * k + j < CPT; ++j) {
                const int c = 4 * k + j;
                const uint32_t inf = info_of(cur.info, c);
                const T rts = cur.rt[c];
                T ra[4];
                ra[0] = (inf & kBlkU) ? rts : up.rt[c];
                ra[1] = (inf & kBlkR) ? rts : (c == CPT - 1 ? cur.rtr : cur.rt[c + 1 < CPT ? c + 1 : c]);
                ra[2] = (inf & kBlkD) ? rts : dn.rt[c];
                ra[3] = (inf & kBlkL) ? rts : (c == 0 ? cur.rtl : cur.rt[c > 0 ? c - 1 : c]);
                const T m = max_nn(max_nn(ra[0], ra[1]), max_nn(ra[2], ra[3]));
#pragma unroll
                for (int a = 0; a < 4; ++a) tie_accumulate(ra[a], m, static_cast<float>(1u << (a + 8 * (j & 1))), acc2[j >> 1]);
              }
              const uint32_t word = __byte_perm(__float_as_uint(acc2[0]), __float_as_uint(acc2[1]), 0x5410u);
              const uint32_t term = (cur.info[k] | (cur.info[k] << 1)) & 0x10101010u;   // goal (bit 3) | lava (bit 4)
              ties[k] = word & ~((term >> 4) * 15u);                                    // utils.py:70: all-zero rows
            }
          } else if constexpr (kPack && KIND == GU_POLICY_GREEDY && !WRITE_TIE) {
#pragma unroll
            for (int c = 0; c < CPT; c += 2) {
              float ga[2][4], ra[2][4], m[2];
              uint32_t off[2];
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const int cc = c + k;
                const uint32_t inf = info_of(cur.info, cc);
                const float gs = cur.g[cc], rts = cur.rt[cc];
                ga[k][0] = (inf & kBlkU) ? gs : up.g[cc];
                ga[k][1] = (inf & kBlkR) ? gs : (cc == CPT - 1 ? cur.gr : cur.g[cc + 1 < CPT ? cc + 1 : cc]);
                ga[k][2] = (inf & kBlkD) ? gs : dn.g[cc];
                ga[k][3] = (inf & kBlkL) ? gs : (cc == 0 ? cur.gl : cur.g[cc > 0 ? cc - 1 : cc]);
                ra[k][0] = (inf & kBlkU) ? rts : up.rt[cc];
                ra[k][1] = (inf & kBlkR) ? rts : (cc == CPT - 1 ? cur.rtr : cur.rt[cc + 1 < CPT ? cc + 1 : cc]);
                ra[k][2] = (inf & kBlkD) ? rts : dn.rt[cc];
                ra[k][3] = (inf & kBlkL) ? rts : (cc == 0 ? cur.rtl : cur.rt[cc > 0 ? cc - 1 : cc]);
                m[k] = max_nn(max_nn(ra[k][0], ra[k][1]), max_nn(ra[k][2], ra[k][3]));
                off[k] = lut_offset<T>(cur.info, cc);
              }
              float o0, o1;
              backup_ties_pair(ra[0], ra[1], m[0], m[1], ga[0], ga[1], off[0], off[1], luts, o0, o1);
              out[c] = o0;
              out[c + 1] = o1;
            }
          } else {
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              const uint32_t inf = info_of(cur.info, c);
              const T gs = cur.g[c];
              // discounted value of the landing cell per action: UP, RIGHT, DOWN, LEFT
              T ga[4];
              ga[0] = (inf & kBlkU) ? gs : up.g[c];
              ga[1] = (inf & kBlkR) ? gs : (c == CPT - 1 ? cur.gr : cur.g[c + 1 < CPT ? c + 1 : c]);
              ga[2] = (inf & kBlkD) ? gs : dn.g[c];
              ga[3] = (inf & kBlkL) ? gs : (c == 0 ? cur.gl : cur.g[c > 0 ? c - 1 : c]);
              const auto rp = reward_poison_at(luts, lut_offset<T>(cur.info, c));   // {R[s], NaN if s terminal else 0}
              const T rs = rp.x;
              T ra[4], m = T(0);
              if constexpr (TIES) {
                const T rts = cur.rt[c];
                ra[0] = (inf & kBlkU) ? rts : up.rt[c];
                ra[1] = (inf & kBlkR) ? rts : (c == CPT - 1 ? cur.rtr : cur.rt[c + 1 < CPT ? c + 1 : c]);
                ra[2] = (inf & kBlkD) ? rts : dn.rt[c];
                ra[3] = (inf & kBlkL) ? rts : (c == 0 ? cur.rtl : cur.rt[c > 0 ? c - 1 : c]);
                // terminal rows are all zero (utils.py:70): NaN never compares equal
                m = N::add(max_nn(max_nn(ra[0], ra[1]), max_nn(ra[2], ra[3])), rp.y);
              }
              if constexpr (WRITE_TIE) {
                const uint32_t mk = (ra[0] == m ? 1u : 0u) | (ra[1] == m ? 2u : 0u) | (ra[2] == m ? 4u : 0u) |
                                    (ra[3] == m ? 8u : 0u);
                ties[c >> 2] |= mk << (8 * (c & 3));
              } else if constexpr (KIND == GU_POLICY_GREEDY) {
                out[c] = backup_ties(rs, ra, m, ga, luts);
              } else if constexpr (KIND == GU_POLICY_PROBS) {
                const uint4* mine = &pstage[((threadIdx.x >> 5) * 2 + ((ry - ry0) & 1)) * kStage + lane * (UPT + 1)];
                T pp[4];
                if constexpr (sizeof(T) == 4) {
                  const float4 f = *reinterpret_cast<const float4*>(mine + c);
                  pp[0] = f.x; pp[1] = f.y; pp[2] = f.z; pp[3] = f.w;
                } else {
                  const double2 d0 = *reinterpret_cast<const double2*>(mine + 2 * c);
                  const double2 d1 = *reinterpret_cast<const double2*>(mine + 2 * c + 1);
                  pp[0] = d0.x; pp[1] = d0.y; pp[2] = d1.x; pp[3] = d1.y;
                }
                T acc = rs;
  #pragma unroll
                for (int a = 0; a < 4; ++a) acc = N::add(acc, N::mul(pp[a], ga[a]));
                out[c] = acc;
              } else {
                const uint32_t mk = KIND == GU_POLICY_UNIFORM ? 15u : (pm[c >> 2] >> (8 * (c & 3))) & 15u;
                const T p = KIND == GU_POLICY_UNIFORM ? T(0.25) : luts.inv_cnt[__popc(mk)];
                T acc = rs;
  #pragma unroll
                for (int a = 0; a < 4; ++a)
                  if ((mk >> a) & 1u) acc = N::add(acc, N::mul(p, ga[a]));
                out[c] = acc;
              }
            }
          }
          if constexpr (WRITE_TIE) {
            if (CPT == 2) *reinterpret_cast<uint16_t*>(tie_out + o) = static_cast<uint16_t>(ties[0]);
            else {
#pragma unroll
              for (int k = 0; k < IW; ++k) *reinterpret_cast<uint32_t*>(tie_out + o + 4 * k) = ties[k];
            }
          } else {
            if constexpr (kPack) {
              if (full) {
#pragma unroll
                for (int c = 0; c < CPT; c += 2) {      // v - out = fma(out, -1, v): one rounding, like the scalar subtract
                  const float2 d = __ffma2_rn(make_float2(out[c], out[c + 1]), make_float2(-1.0f, -1.0f),
                                              make_float2(cur.v[c], cur.v[c + 1]));
                  dmax = max_nn(max_nn(dmax, d.x), d.y);
                }
              }
            } else if (full) {
#pragma unroll
              for (int c = 0; c < CPT; ++c) dmax = max_nn(dmax, N::add(cur.v[c], -out[c]));
            }
            if (!full) {
#pragma unroll
              for (int c = 0; c < CPT; ++c)
                if (x0 + c < g.X) dmax = max_nn(dmax, N::add(cur.v[c], -out[c]));
            }
#pragma unroll
            for (int k = 0; k < NV; ++k) *reinterpret_cast<V*>(vout + o + k * W) = pack(out + k * W);
          }
        }
      }
    }
  }
  if constexpr (PEER && !WRITE_TIE) {
    // The shard's first / last output row also goes into the neighbour's ghost row (peer memory over
    // NVLink).  Done after the row loop -- each thread reads back the vectors it has just stored -- so
    // the loop itself carries no peer code (inside it the two per-row tests cost 27 % at 8192 rows).
    if (active && (top_edge || bot_edge)) {
      if (top_edge) {
#pragma unroll
        for (int k = 0; k < NV; ++k)
          *reinterpret_cast<V*>(peer.up_ghost + x0 + k * W) = *reinterpret_cast<const V*>(vout + pitch + x0 + k * W);
      }
      if (bot_edge) {
#pragma unroll
        for (int k = 0; k < NV; ++k)
          *reinterpret_cast<V*>(peer.down_ghost + x0 + k * W) =
              *reinterpret_cast<const V*>(vout + rows * pitch + x0 + k * W);
      }
    }
  }
  if (!WRITE_TIE) {
    dmax = warp_max(dmax);
    if (lane == 0) scratch[threadIdx.x >> 5] = dmax;
    // ghost-row stores of the first / last block row are out (system scope) before the block reports
    if (PEER && (top_edge || bot_edge)) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (PEER) {
        // the last block of an edge block row raises the neighbour's halo flag
        if (top_edge && atomicAdd(peer.edge + 0, 1) == static_cast<int>(gridDim.x) - 1) {
          peer.edge[0] = 0;
          __threadfence_system();
          *reinterpret_cast<volatile int*>(peer.up_flag) = slot + 1;
        }
        if (bot_edge && atomicAdd(peer.edge + 1, 1) == static_cast<int>(gridDim.x) - 1) {
          peer.edge[1] = 0;
          __threadfence_system();
          *reinterpret_cast<volatile int*>(peer.down_flag) = slot + 1;
        }
      }
      if (residual != nullptr) {
        T m = scratch[0];
#pragma unroll
        for (int k = 1; k < kTiledWarps; ++k) m = scratch[k] > m ? scratch[k] : m;
        atomic_max_signed(residual, m);
#ifndef GU_PEER_NO_DONE
        if (PEER) {
          __threadfence();
          const int nblocks = gridDim.x * gridDim.y;
          if (atomicAdd(peer.done, 1) == nblocks - 1) {     // last block of the sweep
            __threadfence();
            const T fin = *reinterpret_cast<volatile T*>(residual);
            *peer.done = 0;
            peer_publish(peer, slot, fin);
          }
        }
#endif
      }
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
    const uint32_t bu = (term || y == 0 || (inner_up && bit(g.wall, -1, x))) ? kBlkU : 0u;
    const uint32_t br = (term || x == g.X - 1 || bit(g.wall, 0, x + 1)) ? kBlkR : 0u;
    const uint32_t bd = (term || y == g.Y - 1 || (inner_dn && bit(g.wall, 1, x))) ? kBlkD : 0u;
    const uint32_t bl = (term || x == 0 || bit(g.wall, 0, x - 1)) ? kBlkL : 0u;
    v = bu | br | bd | bl | (goal ? kGoal : 0u) | (lava ? kLava : 0u);
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
  if ((static_cast<size_t>(g->pitch) * elem) % 16 != 0 || g->pitch % 32 != 0) return false;
  // the kernels index with 32-bit element offsets (and a 32-bit byte offset for the V prefetch)
  if (static_cast<int64_t>(g->row_end - g->row_begin + 2) * g->pitch * static_cast<int64_t>(elem) >= (1ll << 32) ||
      static_cast<int64_t>(g->row_end - g->row_begin + 2) * g->pitch >= (1ll << 31))
    return false;
  if ((reinterpret_cast<uintptr_t>(a) & 15u) || (reinterpret_cast<uintptr_t>(b) & 15u)) return false;
  if (reinterpret_cast<uintptr_t>(g->info) & 3u) return false;
  return true;
}

#ifndef GU_TILED_NV_F32
#define GU_TILED_NV_F32 2
#endif
#ifndef GU_TILED_ROWS_PER_BLOCK
#define GU_TILED_ROWS_PER_BLOCK 48
#endif
#ifndef GU_TILED_NV_F64
#define GU_TILED_NV_F64 2
#endif
template <typename T> struct TiledNV;
template <> struct TiledNV<float> { static constexpr int value = GU_TILED_NV_F32; };
template <> struct TiledNV<double> { static constexpr int value = GU_TILED_NV_F64; };

// Rows per block: the default amortises the two halo rows of a block over 48 rows.  When a shard is
// small enough that the grid is only a few waves of resident blocks, pick the row-block count that
// fills whole waves instead (e.g. 2048 rows x 16 column blocks on 148 SMs x 4 blocks: 56 rows per
// block = exactly one wave instead of 1.16).
template <typename K>
static int choose_rows_per_block(K kernel, int rows, int blocks_x) {
  const int def = GU_TILED_ROWS_PER_BLOCK;
  static const char* fixed = getenv("GU_TILED_FIXED_RPB");   // developer switches for A/B timing
  if (fixed) return def;
  static const char* force = getenv("GU_TILED_RPB");
  if (force && atoi(force) > 0) return atoi(force);
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTiledWarps * 32, 0) != cudaSuccess || per_sm <= 0)
    return def;
  const long long slots = static_cast<long long>(sms) * per_sm;
  const long long blocks_def = static_cast<long long>(blocks_x) * ((rows + def - 1) / def);
  if (blocks_def >= 6 * slots) return def;            // many waves: quantisation is negligible
  for (int w = 1; w <= 8; ++w) {
    const long long nby = (w * slots) / blocks_x;
    if (nby <= 0) continue;
    const int rpb = static_cast<int>((rows + nby - 1) / nby);
    if (rpb >= 16 && rpb <= 72) return rpb;
  }
  return def;
}

// ---- tensor maps of the TMA-staged window ---------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn sweep_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// 2-D map of a row-major array of `elem`-byte map elements: cols x rows, rows row_bytes apart
static bool sweep_map(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t cols, uint64_t rows,
                      uint64_t row_bytes, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn fn = sweep_encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {row_bytes};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// Which window feeds a sweep: TMA tiles through the shared-memory ring, or 16-byte global loads one row
// ahead + L2 prefetch three rows ahead.  Measured on B200 at 16384^2 (ms per sweep, LDG / TMA):
//   f32  fused-greedy 0.482 / 0.499   tie-mask 0.528 / 0.547   uniform 0.393 / 0.387   extraction 0.453 / 0.464
//   f64  fused-greedy 1.269 / 1.192   tie-mask 1.033 / 1.057   uniform 0.707 / 0.715   extraction 0.738 / 0.699
// The fp32 kernels are bound by instruction issue and the ALU pipe, not by the mover, and the ring's
// 65 KB per block costs the tie-mask kernels a resident block; the fp64 fused-greedy sweep and
// extraction (8 warps of 164+ registers per SM, little latency hiding of their own) gain 6 %.
// Default: TMA where it wins.  GU_SWEEP_TMA=1 / 0 forces one path everywhere (both are bit-identical).
static bool sweep_tma_enabled(size_t elem, int kind, bool write_tie) {
  static const char* e = getenv("GU_SWEEP_TMA");
  if (e && (e[0] == '0' || e[0] == '1')) return e[0] == '1';
  return elem == 8 && (write_tie || kind == GU_POLICY_GREEDY);
}
template <typename T>
static bool make_sweep_maps(const gu_grid* g, const T* vin, CUtensorMap* vmap, CUtensorMap* imap) {
  using TG = TmaGeom<T>;
  const uint64_t arows = static_cast<uint64_t>(g->row_end - g->row_begin + 2);
  const uint64_t pitch = static_cast<uint64_t>(g->pitch);
  return sweep_map(vmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, vin, pitch * sizeof(T) / 8, arows, pitch * sizeof(T),
                   TG::kVRowBytes / 8, kTmaRows) &&
         sweep_map(imap, CU_TENSOR_MAP_DATA_TYPE_UINT32, g->info, pitch / 4, arows, pitch, TG::kIRowBytes / 4, kTmaRows);
}
template <typename T>
static size_t sweep_ring_bytes() { return static_cast<size_t>(kTiledWarps) * TmaGeom<T>::kWarpBytes + 128; }

template <typename T>
static int make_peer_args(const gu_peer_links* pl, PeerArgs<T>* out) {
  if (!pl || !pl->done_counter || !pl->error_flag || !pl->halo_flags || !pl->edge_counters || !pl->stop_flag)
    return GU_ERR_NULL;
  if (pl->world < 1 || pl->world > GU_MAX_PEERS || pl->rank < 0 || pl->rank >= pl->world || pl->slot < 0 ||
      pl->slot >= pl->n_slots || pl->gate_lag < 1 || pl->gate_lag > 2 || pl->first_slot < 0)
    return GU_ERR_SHAPE;
  out->rank = pl->rank; out->world = pl->world; out->slot = pl->slot;
  out->lag = pl->gate_lag; out->first_slot = pl->first_slot;
  out->up_ghost = static_cast<T*>(pl->up_ghost);
  out->down_ghost = static_cast<T*>(pl->down_ghost);
  for (int r = 0; r < GU_MAX_PEERS; ++r) {
    out->tables[r] = r < pl->world ? static_cast<T*>(pl->res_tables[r]) : nullptr;
    out->aborts[r] = r < pl->world ? pl->abort_flags[r] : nullptr;
  }
  for (int r = 0; r < pl->world; ++r)
    if (!out->tables[r] || !out->aborts[r]) return GU_ERR_NULL;
  // a neighbour is described by its ghost row AND its flag word, or by neither
  if ((pl->up_ghost == nullptr) != (pl->up_flag == nullptr) || (pl->down_ghost == nullptr) != (pl->down_flag == nullptr))
    return GU_ERR_NULL;
  out->done = pl->done_counter; out->err = pl->error_flag;
  out->halo = pl->halo_flags; out->up_flag = pl->up_flag; out->down_flag = pl->down_flag;
  out->edge = pl->edge_counters; out->stop = pl->stop_flag; out->slot_base = pl->slot_base;
  out->timeout = pl->timeout_cycles > 0 ? pl->timeout_cycles : kPeerTimeoutCycles;
  out->thr = static_cast<T>(pl->threshold);
  return GU_OK;
}

// Optional launch with the programmatic-stream-serialization attribute (GU_SWEEP_PDL=1): the kernel's
// own griddepcontrol.wait then orders it after the previous kernel of the stream while its blocks may
// already take the SM slots the previous sweep frees.  Measured on B200 it changes nothing (2048-row
// shard: 0.0723 -> 0.0721 ms per sweep back to back, 9.48 -> 9.68 ms per 130-sweep solve): the time a
// small shard loses is inside the kernel (first loads, two halo rows per block, uneven finish), not
// between launches.  Off by default.
static bool sweep_pdl_enabled() {
  static const char* e = getenv("GU_SWEEP_PDL");
  return e && e[0] == '1';
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_sweep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                Args... args) {
  if (smem > 0) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = sweep_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename T>
static int launch_tiled_peer(const gu_grid* g, const T* vin, T* vout, int kind, const void* policy, T gamma,
                             T* residual, const gu_peer_links* pl, cudaStream_t st) {
  PeerArgs<T> pa;
  int rc = make_peer_args(pl, &pa);
  if (rc) return rc;
  if (!residual) return GU_ERR_NULL;
  constexpr int NV = TiledNV<T>::value;
  constexpr int CPT = Vec<T>::W * NV;
  const int rows = g->row_end - g->row_begin;
  const int cols_per_block = kTiledWarps * 32 * CPT;
  const int bx = (g->X + cols_per_block - 1) / cols_per_block;
  const int rpb = choose_rows_per_block(sweep_tiled_kernel<T, GU_POLICY_GREEDY, false, NV, true>, rows, bx);
  dim3 grid(bx, (rows + rpb - 1) / rpb);
  if (grid.y > 65535u) return GU_ERR_SHAPE;
  const GridView v = tview(g);
  CUtensorMap vmap = {}, imap = {};
  const bool tma = NV == 2 && sweep_tma_enabled(sizeof(T), kind, false) && make_sweep_maps<T>(g, vin, &vmap, &imap);
#define GU_LAUNCH_PEER_T(KIND, TMA)                                                                                   \
  launch_sweep(sweep_tiled_kernel<T, KIND, false, NV, true, TMA>, grid, dim3(kTiledWarps * 32),                       \
               TMA ? sweep_ring_bytes<T>() : 0, st, v, g->info, vin, vout, static_cast<uint8_t*>(nullptr), policy, gamma, \
               residual, static_cast<const T*>(nullptr), T(0), rpb, pa, vmap, imap)
#define GU_LAUNCH_PEER(KIND)          \
  do {                                \
    if (tma) GU_LAUNCH_PEER_T(KIND, true); \
    else GU_LAUNCH_PEER_T(KIND, false);    \
  } while (0)
  switch (kind) {
    case GU_POLICY_PROBS: GU_LAUNCH_PEER(GU_POLICY_PROBS); break;
    case GU_POLICY_MASK: GU_LAUNCH_PEER(GU_POLICY_MASK); break;
    case GU_POLICY_UNIFORM: GU_LAUNCH_PEER(GU_POLICY_UNIFORM); break;
    case GU_POLICY_GREEDY: GU_LAUNCH_PEER(GU_POLICY_GREEDY); break;
    default: return GU_ERR_MODE;
  }
#undef GU_LAUNCH_PEER
#undef GU_LAUNCH_PEER_T
  GU_CHECK_LAUNCH();
  return GU_OK;
}

template <typename T, bool WRITE_TIE>
static int launch_tiled(const gu_grid* g, const T* vin, T* vout, uint8_t* tie, int kind, const void* policy,
                        T gamma, T* residual, const T* gate, T gate_thr, cudaStream_t st) {
  constexpr int NV = TiledNV<T>::value;
  constexpr int CPT = Vec<T>::W * NV;
  const int rows = g->row_end - g->row_begin;
  const int cols_per_block = kTiledWarps * 32 * CPT;
  const int bx = (g->X + cols_per_block - 1) / cols_per_block;
  const int rpb = choose_rows_per_block(sweep_tiled_kernel<T, GU_POLICY_GREEDY, WRITE_TIE, NV, false>, rows, bx);
  dim3 grid(bx, (rows + rpb - 1) / rpb);
  if (grid.y > 65535u) return GU_ERR_SHAPE;
  const GridView v = tview(g);
  const uint8_t* info = g->info;
  CUtensorMap vmap = {}, imap = {};
  const bool tma = NV == 2 && sweep_tma_enabled(sizeof(T), kind, WRITE_TIE) && make_sweep_maps<T>(g, vin, &vmap, &imap);
#define GU_LAUNCH_T(KIND, TMA)                                                                                      \
  launch_sweep(sweep_tiled_kernel<T, KIND, WRITE_TIE, NV, false, TMA>, grid, dim3(kTiledWarps * 32),                \
               TMA ? sweep_ring_bytes<T>() : 0, st, v, info, vin, vout, tie, policy, gamma, residual, gate, gate_thr, rpb, \
               PeerArgs<T>(), vmap, imap)
#define GU_LAUNCH(KIND)               \
  do {                                \
    if (tma) GU_LAUNCH_T(KIND, true); \
    else GU_LAUNCH_T(KIND, false);    \
  } while (0)
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
#undef GU_LAUNCH_T
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
  if (kind == GU_POLICY_PROBS && (reinterpret_cast<uintptr_t>(policy) & 15u)) return GU_ERR_UNSUPPORTED;
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

extern "C" __attribute__((visibility("default"))) int gu_sweep_peer_f32(
    const gu_grid* g, const float* v_in, float* v_out, int policy_kind, const void* policy, float gamma,
    float* residual, const gu_peer_links* peer, void* stream) {
  if (!g || !v_in || !v_out) return GU_ERR_NULL;
  if (!tiled_ok(g, sizeof(float), v_in, v_out)) return GU_ERR_UNSUPPORTED;
  if (policy_kind == GU_POLICY_PROBS && (!policy || (reinterpret_cast<uintptr_t>(policy) & 15u))) return GU_ERR_ALIGN;
  return launch_tiled_peer<float>(g, v_in, v_out, policy_kind, policy, gamma, residual, peer,
                                  static_cast<cudaStream_t>(stream));
}

extern "C" __attribute__((visibility("default"))) int gu_sweep_peer_f64(
    const gu_grid* g, const double* v_in, double* v_out, int policy_kind, const void* policy, double gamma,
    double* residual, const gu_peer_links* peer, void* stream) {
  if (!g || !v_in || !v_out) return GU_ERR_NULL;
  if (!tiled_ok(g, sizeof(double), v_in, v_out)) return GU_ERR_UNSUPPORTED;
  if (policy_kind == GU_POLICY_PROBS && (!policy || (reinterpret_cast<uintptr_t>(policy) & 15u))) return GU_ERR_ALIGN;
  return launch_tiled_peer<double>(g, v_in, v_out, policy_kind, policy, gamma, residual, peer,
                                   static_cast<cudaStream_t>(stream));
}

extern "C" __attribute__((visibility("default"))) int gu_peer_wait(const gu_peer_links* peer, int is_f64,
                                                                     void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_f64) {
    PeerArgs<double> pa;
    int rc = make_peer_args(peer, &pa);
    if (rc) return rc;
    peer_wait_kernel<double><<<1, 1, 0, st>>>(pa);
  } else {
    PeerArgs<float> pa;
    int rc = make_peer_args(peer, &pa);
    if (rc) return rc;
    peer_wait_kernel<float><<<1, 1, 0, st>>>(pa);
  }
  GU_CHECK_LAUNCH();
  return GU_OK;
}

// signed max of (a - b) over the owned cells of a shard (dynamic_programming.py:44)
namespace gu {
template <typename T>
__global__ void __launch_bounds__(256)
max_diff_kernel(int X, int rows, int pitch, const T* __restrict__ a, const T* __restrict__ b, T* out) {
  __shared__ T scratch[8];
  T m = Num<T>::neg_inf();
  const int x4 = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (x4 < X) {
    for (int ry = blockIdx.y; ry < rows; ry += gridDim.y) {
      const size_t o = static_cast<size_t>(ry + 1) * pitch + x4;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (x4 + j < X) m = max_nn(m, Num<T>::add(a[o + j], -b[o + j]));
    }
  }
  block_max_to_global(m, scratch, out);
}
template <typename T>
static int launch_max_diff(const gu_grid* g, const T* a, const T* b, T* out, cudaStream_t st) {
  if (!g || !a || !b || !out) return GU_ERR_NULL;
  if (g->X <= 0 || g->row_end <= g->row_begin || g->pitch < g->X) return GU_ERR_SHAPE;
  const int rows = g->row_end - g->row_begin;
  dim3 grid((g->X + 1023) / 1024, rows < 592 ? rows : 592);
  max_diff_kernel<T><<<grid, 256, 0, st>>>(g->X, rows, g->pitch, a, b, out);
  GU_CHECK_LAUNCH();
  return GU_OK;
}
}  // namespace gu

extern "C" __attribute__((visibility("default"))) int gu_max_diff_f32(const gu_grid* g, const float* a, const float* b,
                                                                        float* out, void* stream) {
  return gu::launch_max_diff<float>(g, a, b, out, static_cast<cudaStream_t>(stream));
}
extern "C" __attribute__((visibility("default"))) int gu_max_diff_f64(const gu_grid* g, const double* a,
                                                                        const double* b, double* out, void* stream) {
  return gu::launch_max_diff<double>(g, a, b, out, static_cast<cudaStream_t>(stream));
}

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
