// gu_env_tables.cu -- table-driven rollout kernels (transition tables staged in shared memory).
//
// A level is static, so look_step_ahead (core/envs/griduniverse_env.py:136-155) can be
// tabulated once per level.  gu_pack_tables builds the tables on the device with the same
// transition() code the layout-agnostic kernels use; the rollout kernels stage each env's table in
// shared memory once per launch and then pay one conflict-free shared load per env step.
//
// Formats
//   INFO8 (per-env levels, X*Y <= 256, X <= 127): one byte per cell, bits 0-3 = action a moves the
//        agent (not a grid edge, not a wall, cell not terminal), bit 6 goal, bit 7 lava of the cell
//        itself (goal cleared when lava is set: lava wins, griduniverse_env.py:86-90); four cells per
//        32-bit word, WORD-MAJOR uint32[ceil(cells/4)][N].  64 bytes per 8x8 env, so a dozen warps of
//        envs stay resident per SM.  (A next-state table -- landing cell per (cell, action), 256 B per
//        8x8 env -- needs fewer instructions per step but leaves one warp per scheduler; it measured
//        6.5 ms against 2.8 ms on BASELINE cfg 4, see profiles/r1_history.md.)
//   NT16 (shared level, X*Y <= 16383): uint16 per (cell, action): landing in bits 0-13, goal
//        bit 14, lava bit 15.  uint16[cells][4], read by every env of the batch.
#include <cuda.h>

#include <cstdlib>

#include "gu_env.cuh"

namespace gu {

enum TableFormat { kTableNone = 0, kTableNT16 = 2, kTableINFO8 = 3 };

static TableFormat table_format(const gu_levels* lv) {
  const int64_t cells = static_cast<int64_t>(lv->X) * lv->Y;
  if (lv->per_env && cells <= 256 && lv->X <= 127) return kTableINFO8;
  if (!lv->per_env && cells <= 16383) return kTableNT16;
  return kTableNone;
}

// ---- table construction -------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pack_info8_kernel(LevelsView lv, uint32_t* __restrict__ tables) {
  const int64_t env = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (env >= lv.N) return;
  const int cells = lv.X * lv.Y;
  const int words = (cells + 3) / 4;
  for (int w = 0; w < words; ++w) {
    uint32_t word = 0;
    for (int j = 0; j < 4; ++j) {
      const int s = 4 * w + j;
      if (s >= cells) break;
      uint32_t b = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        int n, r;
        bool term;
        transition(lv, env, s, a, true, n, r, term);
        b |= (n != s ? 1u : 0u) << a;
      }
      const bool lava = plane_bit(lv.lava, lv, env, s), goal = plane_bit(lv.goal, lv, env, s);
      b |= lava ? 0x80u : (goal ? 0x40u : 0u);
      word |= b << (8 * j);
    }
    tables[static_cast<int64_t>(w) * lv.N + env] = word;
  }
}

__global__ void __launch_bounds__(128)
pack_nt16_kernel(LevelsView lv, uint16_t* __restrict__ tables) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (cell, action)
  const int cells = lv.X * lv.Y;
  if (i >= cells * 4) return;
  int n, r;
  bool term;
  transition(lv, 0, i >> 2, i & 3, true, n, r, term);
  const uint32_t f = r == kRewardLava ? 0x8000u : (r == kRewardGoal ? 0x4000u : 0u);
  tables[i] = static_cast<uint16_t>(static_cast<uint32_t>(n) | f);
}

#ifndef GU_ROLLOUT_WARPS
#define GU_ROLLOUT_WARPS 4
#endif

// ---- rollout over INFO8 tables, TMA-tiled staging ---------------------------------------------------
// A warp owns 32*EPT consecutive envs and lane l steps envs l, l+32, ...  All HBM traffic is 2-D
// TMA tile copies (cp.async.bulk.tensor, SASS UTMALDG) issued by one elected lane and completed
// on per-warp mbarriers: the env range's info table once ([words x EPW] box of the
// [words][N] table), then the action stream as [kInfoRows x EPW] boxes of the [T][N] action
// matrix through a kInfoStages-deep ring, so several KB per warp are always in flight while the
// lanes step.  Tiles land in shared memory exactly as stored, and lane l only ever reads column
// k*32+l, so every shared load of a warp hits 32 different banks.  No block-level barrier.
//
// Per step: allowed = bit a of the current cell's info byte, the landing cell is
// pos + allowed * delta[a] (delta = -X, +1, +X, -1 from a byte LUT), then one shared load fetches
// the landing cell's info byte, whose goal / lava bits give reward and done (griduniverse_env.py:155).
#ifndef GU_INFO8_ROWS
#define GU_INFO8_ROWS 8
#endif
#ifndef GU_INFO8_STAGES
#define GU_INFO8_STAGES 2
#endif
// Ring geometry (template parameters of the kernel): action rows per TMA box, boxes in flight per
// warp, warps per block.  The shallow ring is the default everywhere.  The deep one (4 x 16 rows = 8 KB
// in flight per warp) was built for batches that leave an SM with only a dozen warps (BASELINE cfg 3:
// 65,536 envs = 14 warps per SM) on the theory that bytes in flight bound them; measured on B200 it is
// SLOWER there (0.100 against 0.081 ms): with one env per lane that workload is bound by the dependent
// chain of a step (position -> table word -> landing cell, ~150 cycles) times 1024 sequential steps, not
// by bandwidth -- the packed action stream, a sixteenth of the bytes, still takes 0.065 ms.  It stays
// selectable (GU_INFO8_RING=deep) for experiments.
struct RingStd { static constexpr int kRows = GU_INFO8_ROWS, kStages = GU_INFO8_STAGES, kWarps = GU_ROLLOUT_WARPS; };
struct RingDeep { static constexpr int kRows = 16, kStages = 4, kWarps = 2; };
// One env per lane (small batches): a box of 8 steps is consumed in ~600 cycles, less than a DRAM round
// trip under load, so with two stages every box wait is exposed.  Four stages of the same 8-row box put
// three box times between a tile's issue and its use and still leave 16 warps per SM resident for a
// 16x16 level (12 KB per warp), which the 16-row ring above does not.
struct RingMid { static constexpr int kRows = 8, kStages = 4, kWarps = 4; };
#ifndef GU_INFO8_PATCH
#define GU_INFO8_PATCH 1
#endif

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}

// PACKED: the action stream holds 2 bits per step, 16 steps per 32-bit word (uint32[ceil(T/16)][N],
// step t of env n in bits 2*(t % 16) of word [t / 16][n]): a sixteenth of the bytes, the same steps.
// SC: the start state an env continues from after a done step comes from a host-supplied stream
// start_choice[T][N] (the stand-in for random.choice over several 'x' cells, griduniverse_env.py:189),
// which travels through the same ring as the actions, one box behind each action box.
template <int EPT, bool TRAJ, bool AUTO_RESET, typename RING, bool PACKED, bool SC>
__global__ void __launch_bounds__(RING::kWarps * 32)
rollout_info8_tma_kernel(const __grid_constant__ CUtensorMap act_map, const __grid_constant__ CUtensorMap tab_map,
                         const __grid_constant__ CUtensorMap sc_map, int N, int T, int X, int words, int32_t* __restrict__ pos, int32_t* __restrict__ obs,
                         int32_t* __restrict__ reward, uint8_t* __restrict__ done,
                         const int32_t* __restrict__ start, int32_t* __restrict__ env_return,
                         int32_t* __restrict__ env_done, int64_t* stats, uint32_t flags) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kInfoRows = RING::kRows, kInfoStages = RING::kStages, kBulkWarps = RING::kWarps;
  constexpr int EPW = 32 * EPT, ROWB = EPW * 4;
  constexpr uint32_t kActBytes = kInfoRows * ROWB;              // one box of actions
  constexpr uint32_t kBoxBytes = (SC ? 2 : 1) * kActBytes;       // one ring stage: actions [+ start choices]
  static_assert(!(SC && PACKED), "start-choice streams are per step: int32 actions only");
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);       // warp-uniform for the TMA operands
  const uint32_t per_warp = static_cast<uint32_t>(words) * ROWB + kInfoStages * kBoxBytes;
  const uint32_t tab_s = ((smem_u32(smem_raw) + 127u) & ~127u) + warp * per_warp;   // [words][EPW] words
  const uint32_t act_s = tab_s + static_cast<uint32_t>(words) * ROWB;      // [stages][rows][EPW] words
  __shared__ __align__(8) uint64_t bars[kBulkWarps][kInfoStages + 1];
  const uint32_t bar0 = smem_u32(&bars[warp][0]);
  const uint32_t bar_tab = bar0 + 8 * kInfoStages;
  const int env0 = (blockIdx.x * kBulkWarps + warp) * EPW;               // N % EPW == 0
  const bool accumulate = flags & GU_FLAG_ACCUMULATE;
  const int Trows = PACKED ? (T + 15) / 16 : T;                       // rows of the action matrix
  const int nbatch = (Trows + kInfoRows - 1) / kInfoRows;
  // per action: {bit of the info byte that says "a moves", scaled position delta}: UP -X, RIGHT +1, DOWN +X, LEFT -1
  __shared__ __align__(8) int2 act_lut[4];
  if (threadIdx.x < 4) {
    const int a = threadIdx.x;
    act_lut[a] = make_int2(1 << a, 8 * (a == 0 ? -X : (a == 1 ? 1 : (a == 2 ? X : -1))));
  }
  __syncthreads();
  const uint32_t lut_s = smem_u32(&act_lut[0]);
  long long rsum = 0, dcnt = 0;

  if (env0 < N) {
    if (elect_one()) {
      for (int i = 0; i <= kInfoStages; ++i) mbar_init(bar0 + 8 * i, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (elect_one()) {
      mbar_expect_tx(bar_tab, static_cast<uint32_t>(words) * ROWB);
      tma_load_2d(tab_s, &tab_map, env0, 0, bar_tab);
      for (int b = 0; b < kInfoStages && b < nbatch; ++b) {
        mbar_expect_tx(bar0 + 8 * b, kBoxBytes);
        tma_load_2d(act_s + b * kBoxBytes, &act_map, env0, b * kInfoRows, bar0 + 8 * b);
        if (SC) tma_load_2d(act_s + b * kBoxBytes + kActBytes, &sc_map, env0, b * kInfoRows, bar0 + 8 * b);
      }
    }
    // The position travels pre-scaled, q = 8 * cell: the funnel shift that brings the cell's byte down takes
    // q itself (amount modulo 32) and the row offset of its table word is (q & ~31) * (ROWB / 32) -- one
    // mask and one multiply-add (FMA pipe) on the dependent chain position -> table word -> landing cell.
    // The kernels are bound by the half-rate ALU pipe (cfg 4, packed actions: 92 %) or by that chain under
    // contention for it (BASELINE cfg 3: 14 warps per SM, 1024 sequential steps), so ALU-pipe instructions
    // per step are what counts.
    constexpr int QS = 8, QLOG = 3;
    // With auto-reset to a fixed start the reset of the info byte leaves the chain as well: once the
    // table is staged every lane rewrites the "action moves" nibble of the TERMINAL cells of its envs to
    // the start cell's nibble, so the byte fetched for a terminal landing cell already is the byte the
    // next step needs (its goal / lava bits still count the episode end); only the position is selected.
    constexpr bool kPatch = GU_INFO8_PATCH && AUTO_RESET && !SC;
    int q[EPT], stq[EPT];
    uint32_t inf[EPT], inf_st[EPT], fsum[EPT], fsq[EPT];
    uint32_t tabk[EPT];                              // shared-memory byte address of the env's column
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      q[k] = pos[env0 + k * 32 + lane] * QS;
      stq[k] = (AUTO_RESET && !SC) ? start[env0 + k * 32 + lane] * QS : 0;      // lv->start may be NULL otherwise
      fsum[k] = 0;
      fsq[k] = 0;
      tabk[k] = tab_s + (k * 32 + lane) * 4;
    }
    mbar_wait(bar_tab, 0);
    // info byte of the cell at scaled position c = 8 * cell: word cell >> 2 of the env's column (rows are
    // ROWB bytes apart), byte cell & 3
    auto info_at = [&](int k, int c) -> uint32_t {
      uint32_t a;                                   // kept as mask + multiply-add: the compiler's shift + mask + add is one deeper
      asm("{\n\t"
          ".reg .b32 t;\n\t"
          "and.b32 t, %1, 0xffffffe0;\n\t"
          "mad.lo.u32 %0, t, %2, %3;\n\t"
          "}"
          : "=r"(a)
          : "r"(c), "n"(ROWB / 32), "r"(tabk[k]));
      const uint32_t word = lds_u32(a);
      return __funnelshift_r(word, 0u, static_cast<uint32_t>(c));
    };
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      inf[k] = info_at(k, q[k]);
      inf_st[k] = (AUTO_RESET && !SC) ? info_at(k, stq[k]) : 0u;
    }
    if constexpr (kPatch) {
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const uint32_t nib = (inf_st[k] & 0xfu) * 0x01010101u;
        for (int w = 0; w < words; ++w) {
          const uint32_t a = tabk[k] + static_cast<uint32_t>(w) * ROWB;
          const uint32_t word = lds_u32(a);
          const uint32_t term = (word | (word << 1)) & 0x80808080u;      // goal 0x40 | lava 0x80 per byte
          if (term) {
            const uint32_t m = (term >> 7) * 0x0fu;
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"((word & ~m) | (nib & m)) : "memory");
          }
        }
      }
    }

    // `la` = 8 * action: byte offset into the block's {1 << a, 8 * delta[a]} table (one 8-byte shared load
    // instead of a shift, a multiply and a PRMT per step)
    auto step_one = [&](int k, uint32_t la, int t, uint32_t sc_addr) {
      uint32_t abit;
      int ds;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(abit), "=r"(ds) : "r"(lut_s + la));
      // bit a of the current cell's info byte: the action moves.  One predicate-setting LOP3 and one
      // predicated add on the chain (the compiler's own form, shift + mask + compare + select + add, is 5 deep).
      int n = q[k];
      asm("{\n\t"
          ".reg .pred mv;\n\t"
          ".reg .b32 t;\n\t"
          "and.b32 t, %1, %2;\n\t"
          "setp.ne.u32 mv, t, 0;\n\t"
          "@mv add.s32 %0, %0, %3;\n\t"
          "}"
          : "+r"(n)
          : "r"(inf[k]), "r"(abit), "r"(ds));
      uint32_t i2 = info_at(k, n);
      const uint32_t f = i2 & 0xc0u;               // goal 0x40 / lava 0x80 of the landing cell
      if (TRAJ) {
        const int64_t o = static_cast<int64_t>(t) * N + env0 + k * 32 + lane;
        if (obs) obs[o] = n >> QLOG;
        if (reward) reward[o] = (f & 0x80u) ? kRewardLava : ((f & 0x40u) ? kRewardGoal : kRewardStep);
        if (done) done[o] = f ? 1 : 0;
      }
      fsum[k] += f;                               // 64*goals + 128*lavas
      fsq[k] += f * f;                            // 4096*goals + 16384*lavas
      if (AUTO_RESET && f) {
        if (SC) { n = static_cast<int>(lds_u32(sc_addr + k * 128)) * QS; i2 = info_at(k, n); }
        else { n = stq[k]; if (!kPatch) i2 = inf_st[k]; }
      }
      q[k] = n;
      inf[k] = i2;
    };
    // one row of the action matrix: one step (int32 actions) or sixteen (packed)
    auto step_row = [&](uint32_t arow, int row, bool full) {   // arow: smem address of this lane's word in the row
      if (!PACKED) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) step_one(k, (lds_u32(arow + k * 128) << 3) & 24u, row, arow + kActBytes);
      } else {
        uint32_t w[EPT];
#pragma unroll
        for (int k = 0; k < EPT; ++k) w[k] = lds_u32(arow + k * 128);
        if (full) {
#pragma unroll
          for (int s = 0; s < 16; ++s) {
#pragma unroll
            for (int k = 0; k < EPT; ++k)
              step_one(k, (s == 0 ? w[k] << 3 : (s == 1 ? w[k] << 1 : w[k] >> (2 * s - 3))) & 24u, row * 16 + s, 0u);
          }
        } else {
          for (int s = 0; row * 16 + s < T; ++s) {
#pragma unroll
            for (int k = 0; k < EPT; ++k) step_one(k, ((w[k] >> (2 * s)) & 3u) << 3, row * 16 + s, 0u);
          }
        }
      }
    };

    int stage = 0;
    uint32_t parity = 0;
    for (int b = 0; b < nbatch; ++b) {
      const int t0 = b * kInfoRows;
      mbar_wait(bar0 + 8 * stage, parity);
      const uint32_t arow = act_s + stage * kBoxBytes + lane * 4;
      if (PACKED ? (t0 + kInfoRows) * 16 <= T : t0 + kInfoRows <= T) {
#pragma unroll (PACKED ? 1 : kInfoRows)
        for (int r = 0; r < kInfoRows; ++r) step_row(arow + r * ROWB, t0 + r, true);
      } else {                                         // rows past the end are zero fill and never stepped
        for (int r = 0; t0 + r < Trows; ++r) step_row(arow + r * ROWB, t0 + r, !PACKED || (t0 + r + 1) * 16 <= T);
      }
      __syncwarp();                                   // every lane is done with this stage
      if (b + kInfoStages < nbatch && elect_one()) {
        mbar_expect_tx(bar0 + 8 * stage, kBoxBytes);
        tma_load_2d(act_s + stage * kBoxBytes, &act_map, env0, (b + kInfoStages) * kInfoRows, bar0 + 8 * stage);
        if (SC) tma_load_2d(act_s + stage * kBoxBytes + kActBytes, &sc_map, env0, (b + kInfoStages) * kInfoRows, bar0 + 8 * stage);
      }
      if (++stage == kInfoStages) { stage = 0; parity ^= 1u; }
    }
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      const uint32_t lavas = (fsq[k] - 64u * fsum[k]) >> 13;
      const uint32_t goals = (fsum[k] - 128u * lavas) >> 6;
      const long long dones = static_cast<long long>(goals) + lavas;
      const long long ret = -(static_cast<long long>(T) - dones) + 10ll * goals - 10ll * lavas;
      rsum += ret;
      dcnt += dones;
      const int e = env0 + k * 32 + lane;
      pos[e] = q[k] >> QLOG;
      if (env_return) env_return[e] = static_cast<int>(ret) + (accumulate ? env_return[e] : 0);
      if (env_done) env_done[e] = static_cast<int>(dones) + (accumulate ? env_done[e] : 0);
    }
  }
  publish_stats(rsum, dcnt, stats);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
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
// 2-D map of a row-major int32 [rows][cols] matrix, box = [box_rows][box_cols], out-of-range rows zero-filled
static bool make_map_i32(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_rows, int box_cols) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 4};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename RING>
static size_t info8_smem_bytes(int words, int ept, bool sc = false) {
  return static_cast<size_t>(RING::kWarps) * (static_cast<size_t>(words) + (sc ? 2 : 1) * RING::kStages * RING::kRows) * 32 * ept * 4 + 128;
}

template <int EPT, bool TRAJ, bool AR, typename RING, bool PACKED, bool SC = false>
static int launch_info8(const gu_levels* lv, int64_t n, int64_t T, const int32_t* actions, int32_t* pos, int32_t* obs,
                        int32_t* reward, uint8_t* done, int32_t* env_return, int32_t* env_done, int64_t* stats,
                        const uint32_t* tables, uint32_t flags, cudaStream_t st, const int32_t* start_choice = nullptr) {
  const int cells = lv->X * lv->Y, words = (cells + 3) / 4;
  constexpr int EPW = 32 * EPT;
  const int64_t act_rows = PACKED ? (T + 15) / 16 : T;
  CUtensorMap act_map, tab_map, sc_map;
  if (!make_map_i32(&act_map, actions, act_rows, n, RING::kRows, EPW) ||
      !make_map_i32(&tab_map, tables, words, n, words, EPW))
    return GU_ERR_UNSUPPORTED;
  sc_map = act_map;
  if (SC && !make_map_i32(&sc_map, start_choice, T, n, RING::kRows, EPW)) return GU_ERR_UNSUPPORTED;
  const size_t smem = info8_smem_bytes<RING>(words, EPT, SC);
  const unsigned blocks = static_cast<unsigned>((n / EPW + RING::kWarps - 1) / RING::kWarps);
  auto kernel = rollout_info8_tma_kernel<EPT, TRAJ, AR, RING, PACKED, SC>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  kernel<<<blocks, RING::kWarps * 32, smem, st>>>(act_map, tab_map, sc_map, static_cast<int>(n), static_cast<int>(T), lv->X,
                                                 words, pos, obs, reward, done, lv->start, env_return, env_done, stats,
                                                 flags);
  cudaError_t le = cudaGetLastError();
  return le == cudaSuccess ? GU_OK : static_cast<int>(le);
}

// ---- rollout over a shared NT16 table -------------------------------------------------------
constexpr int kNt16Threads = 256;

template <bool TRAJ>
__global__ void __launch_bounds__(kNt16Threads)
rollout_nt16_kernel(int64_t N, int64_t T, int cells, const uint16_t* __restrict__ tables,
                    const int32_t* __restrict__ actions, int32_t* __restrict__ pos,
                    int32_t* __restrict__ obs, int32_t* __restrict__ reward, uint8_t* __restrict__ done,
                    const int32_t* __restrict__ start, const int32_t* __restrict__ start_choice,
                    int32_t* __restrict__ env_return, int32_t* __restrict__ env_done, int64_t* stats,
                    uint32_t flags) {
  extern __shared__ uint32_t tab32[];
  uint16_t* tab = reinterpret_cast<uint16_t*>(tab32);
  for (int i = threadIdx.x; i < cells * 2; i += kNt16Threads)
    tab32[i] = reinterpret_cast<const uint32_t*>(tables)[i];
  __syncthreads();
  const int64_t env = static_cast<int64_t>(blockIdx.x) * kNt16Threads + threadIdx.x;
  const bool auto_reset = flags & GU_FLAG_AUTO_RESET;
  const bool accumulate = flags & GU_FLAG_ACCUMULATE;
  long long rsum = 0, dcnt = 0;
  if (env < N) {
    int p = pos[env];
    const int st = (auto_reset && start_choice == nullptr) ? __ldg(start) : 0;   // lv->start may be NULL otherwise
    auto step = [&](int a, int64_t t) {
      const uint32_t v = tab[p * 4 + (a & 3)];
      const int n = static_cast<int>(v & 0x3fffu);
      const bool lava = v & 0x8000u, goal = v & 0x4000u;
      const int r = lava ? kRewardLava : (goal ? kRewardGoal : kRewardStep);
      const bool d = lava | goal;
      if (TRAJ) {
        const int64_t o = t * N + env;
        if (obs) obs[o] = n;
        if (reward) reward[o] = r;
        if (done) done[o] = d;
      }
      rsum += r;
      dcnt += d ? 1 : 0;
      p = n;
      if (auto_reset && d) p = start_choice != nullptr ? __ldg(start_choice + t * N + env) : st;
    };
    if (flags & GU_FLAG_PACKED_ACTIONS) {            // 16 steps per word, next word loaded a word ahead
      const int64_t nw = (T + 15) / 16;
      uint32_t next = static_cast<uint32_t>(__ldg(actions + env));
      for (int64_t wq = 0; wq < nw; ++wq) {
        const uint32_t w = next;
        if (wq + 1 < nw) next = static_cast<uint32_t>(__ldg(actions + (wq + 1) * N + env));
#pragma unroll
        for (int s16 = 0; s16 < 16; ++s16)
          if (wq * 16 + s16 < T) step(static_cast<int>((w >> (2 * s16)) & 3u), wq * 16 + s16);
      }
    } else {
      constexpr int U = 8;
      int abuf[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (u < T) abuf[u] = __ldg(actions + static_cast<int64_t>(u) * N + env);
      for (int64_t t0 = 0; t0 < T; t0 += U) {
        int acur[U];
#pragma unroll
        for (int u = 0; u < U; ++u) acur[u] = abuf[u];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (t0 + U + u < T) abuf[u] = __ldg(actions + (t0 + U + u) * N + env);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (t0 + u < T) step(acur[u], t0 + u);
      }
    }
    pos[env] = p;
    if (env_return) env_return[env] = static_cast<int>(rsum) + (accumulate ? env_return[env] : 0);
    if (env_done) env_done[env] = static_cast<int>(dcnt) + (accumulate ? env_done[env] : 0);
  }
  publish_stats(rsum, dcnt, stats);
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int rollout_tables(const gu_levels* lv, int64_t n, int64_t T, const int32_t* actions, int32_t* pos, int32_t* obs,
                   int32_t* reward, uint8_t* done, const int32_t* start_choice, int32_t* env_return,
                   int32_t* env_done, int64_t* stats, const uint32_t* tables, uint32_t flags, cudaStream_t st) {
  if (flags & GU_FLAG_NO_CARE_TERMINAL) return GU_ERR_UNSUPPORTED;   // tables tabulate care_about_terminal=True
  if (T >= (1 << 18)) return GU_ERR_UNSUPPORTED;                     // packed goal / lava counters are 32-bit
  const TableFormat fmt = table_format(lv);
  const int cells = lv->X * lv->Y;
  const bool traj = obs || reward || done;
  const bool packed = flags & GU_FLAG_PACKED_ACTIONS;
  if (fmt == kTableINFO8) {
    if (!al16(actions) || !al16(tables) || n % 32 != 0 || n >= (1ll << 31)) return GU_ERR_UNSUPPORTED;
    const bool sc = start_choice != nullptr && (flags & GU_FLAG_AUTO_RESET);
    if (sc && (packed || !al16(start_choice))) return GU_ERR_UNSUPPORTED;   // per-step stream: int32 actions only
    // envs per lane: 2 when that still gives every SM a dozen warps (measured best on B200 for large
    // batches), else 1 so small batches spread over all SMs; 4 only on request
    static const char* force = getenv("GU_INFO8_EPT");
    int ept = (n % 64 == 0 && n / 64 >= 148 * 12) ? 2 : 1;
    if (force && (atoi(force) == 1 || atoi(force) == 2 || atoi(force) == 4)) ept = atoi(force);   // developer switch
    if (n % (32 * ept) != 0) return GU_ERR_UNSUPPORTED;
    const bool ar = flags & GU_FLAG_AUTO_RESET;
    static const char* ring_env = getenv("GU_INFO8_RING");        // developer switch: "deep"
    const bool deep = ring_env && ring_env[0] == 'd' && ept == 1 &&
                      info8_smem_bytes<RingDeep>((cells + 3) / 4, 1) <= 100 * 1024;
#define GU_INFO8_ARGS lv, n, T, actions, pos, obs, reward, done, env_return, env_done, stats, tables, flags, st
#define GU_INFO8_P(EPT, RING, PACKED)                                                                 \
  return traj ? (ar ? launch_info8<EPT, true, true, RING, PACKED>(GU_INFO8_ARGS)                      \
                    : launch_info8<EPT, true, false, RING, PACKED>(GU_INFO8_ARGS))                    \
              : (ar ? launch_info8<EPT, false, true, RING, PACKED>(GU_INFO8_ARGS)                     \
                    : launch_info8<EPT, false, false, RING, PACKED>(GU_INFO8_ARGS))
#define GU_INFO8(EPT, RING)            \
  do {                                 \
    if (packed) {                      \
      GU_INFO8_P(EPT, RING, true);     \
    } else {                           \
      GU_INFO8_P(EPT, RING, false);    \
    }                                  \
  } while (0)
    if (sc) {      // multi-start levels: the start-choice stream rides the ring behind the actions
#define GU_INFO8_SC(EPT)                                                                                      \
  return traj ? launch_info8<EPT, true, true, RingStd, false, true>(GU_INFO8_ARGS, start_choice)              \
              : launch_info8<EPT, false, true, RingStd, false, true>(GU_INFO8_ARGS, start_choice)
      if (ept == 4) GU_INFO8_SC(4);
      if (ept == 2) GU_INFO8_SC(2);
      GU_INFO8_SC(1);
#undef GU_INFO8_SC
    }
    if (ept == 4) GU_INFO8(4, RingStd);
    if (ept == 2) GU_INFO8(2, RingStd);
    if (deep) GU_INFO8(1, RingDeep);
    static const bool mid_off = ring_env && ring_env[0] == 's';   // developer switch: "std"
    if (!mid_off && info8_smem_bytes<RingMid>((cells + 3) / 4, 1) <= 56 * 1024) GU_INFO8(1, RingMid);
    GU_INFO8(1, RingStd);
#undef GU_INFO8
#undef GU_INFO8_P
#undef GU_INFO8_ARGS
  }
  if (fmt == kTableNT16) {
    const size_t smem = static_cast<size_t>(cells) * 8;
    const unsigned blocks = static_cast<unsigned>((n + kNt16Threads - 1) / kNt16Threads);
    const uint16_t* t16 = reinterpret_cast<const uint16_t*>(tables);
    if (traj) {
      cudaError_t e = cudaFuncSetAttribute(rollout_nt16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
      if (e != cudaSuccess) return static_cast<int>(e);
      rollout_nt16_kernel<true><<<blocks, kNt16Threads, smem, st>>>(n, T, cells, t16, actions, pos, obs, reward, done,
                                                                   lv->start, start_choice, env_return, env_done,
                                                                   stats, flags);
    } else {
      cudaError_t e = cudaFuncSetAttribute(rollout_nt16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
      if (e != cudaSuccess) return static_cast<int>(e);
      rollout_nt16_kernel<false><<<blocks, kNt16Threads, smem, st>>>(n, T, cells, t16, actions, pos, obs, reward, done,
                                                                    lv->start, start_choice, env_return, env_done,
                                                                    stats, flags);
    }
    GU_CHECK_LAUNCH();
    return GU_OK;
  }
  return GU_ERR_UNSUPPORTED;
}

}  // namespace gu

using namespace gu;

extern "C" __attribute__((visibility("default"))) int64_t gu_tables_bytes(const gu_levels* lv, int64_t n) {
  if (!lv || n < 0) return 0;
  const int64_t cells = static_cast<int64_t>(lv->X) * lv->Y;
  switch (table_format(lv)) {
    case kTableINFO8: return ((cells + 3) / 4) * 4 * n;
    case kTableNT16: return cells * 8;
    default: return 0;
  }
}

extern "C" __attribute__((visibility("default"))) int gu_pack_tables(const gu_levels* lv, int64_t n, uint32_t* tables,
                                                                     uint32_t flags, void* stream) {
  int rc = check_levels(lv, n);
  if (rc) return rc;
  if (!tables) return GU_ERR_NULL;
  (void)flags;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const TableFormat fmt = table_format(lv);
  if (fmt == kTableINFO8) {
    if (n == 0) return GU_OK;
    pack_info8_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(view_of(lv, n), tables);
  } else if (fmt == kTableNT16) {
    const int cells = lv->X * lv->Y;
    pack_nt16_kernel<<<(cells * 4 + 127) / 128, 128, 0, st>>>(view_of(lv, 1), reinterpret_cast<uint16_t*>(tables));
  } else {
    return GU_ERR_UNSUPPORTED;
  }
  GU_CHECK_LAUNCH();
  return GU_OK;
}
