// gu_level_text.cu -- batched level-text packing on the device.
//
// The reference turns one text level into python lists with a row-major character scan
// (core/envs/griduniverse_env.py:253-300): 'o' floor, '#' wall, 'G' goal, 'L' lava, 'x' start,
// anything else raises; afterwards a level without a start or without a goal raises.  For a batch
// of same-shaped levels the scan is a ballot: one warp per level, lane l looks at cell 32k + l and
// the warp votes one word of each bit plane per step.  Output is the word-major per-env layout of
// gu_levels; the per-level status carries what the reference would have raised.
#include "gu_common.cuh"

namespace gu {

constexpr int kTextWarps = 4;

__global__ void __launch_bounds__(kTextWarps * 32)
pack_level_text_kernel(const uint8_t* __restrict__ text, long long n, int cells, int words,
                       uint32_t* __restrict__ wall, uint32_t* __restrict__ goal, uint32_t* __restrict__ lava,
                       int32_t* __restrict__ start, int32_t* __restrict__ n_starts, int32_t* __restrict__ status) {
  const long long lvl = static_cast<long long>(blockIdx.x) * kTextWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (lvl >= n) return;
  const uint8_t* t = text + lvl * cells;
  int first_start = -1, starts = 0, goals = 0, first_bad = -1;
  for (int k = 0; k < words; ++k) {
    const int c = k * 32 + lane;
    const int ch = c < cells ? t[c] : 'o';
    const uint32_t w = __ballot_sync(0xffffffffu, ch == '#');
    const uint32_t g = __ballot_sync(0xffffffffu, ch == 'G');
    const uint32_t l = __ballot_sync(0xffffffffu, ch == 'L');
    const uint32_t x = __ballot_sync(0xffffffffu, ch == 'x');
    const uint32_t bad = ~(w | g | l | x | __ballot_sync(0xffffffffu, ch == 'o'));
    if (lane == 0) {
      wall[static_cast<size_t>(k) * n + lvl] = w;
      goal[static_cast<size_t>(k) * n + lvl] = g;
      lava[static_cast<size_t>(k) * n + lvl] = l;
    }
    if (x && first_start < 0) first_start = k * 32 + __ffs(x) - 1;
    starts += __popc(x);
    goals += __popc(g);
    if (bad && first_bad < 0) first_bad = k * 32 + __ffs(bad) - 1;
  }
  if (lane == 0) {
    start[lvl] = first_start < 0 ? 0 : first_start;
    if (n_starts) n_starts[lvl] = starts;
    // the reference's order: invalid character during the scan, then no start, then no goal
    status[lvl] = first_bad >= 0 ? first_bad + 1 : (starts == 0 ? GU_TEXT_NO_START : (goals == 0 ? GU_TEXT_NO_GOAL : 0));
  }
}

}  // namespace gu

using namespace gu;

extern "C" __attribute__((visibility("default"))) int gu_pack_level_text(
    const uint8_t* text, int64_t n_levels, int32_t X, int32_t Y, uint32_t* wall, uint32_t* goal, uint32_t* lava,
    int32_t* start, int32_t* n_starts, int32_t* status, void* stream) {
  if (n_levels == 0) return GU_OK;
  if (!text || !wall || !goal || !lava || !start || !status) return GU_ERR_NULL;
  if (X <= 0 || Y <= 0 || n_levels < 0 || static_cast<int64_t>(X) * Y > (1 << 24)) return GU_ERR_SHAPE;
  const int cells = X * Y, words = (cells + 31) / 32;
  const long long blocks = (n_levels + kTextWarps - 1) / kTextWarps;
  if (blocks > 2147483647LL) return GU_ERR_SHAPE;
  pack_level_text_kernel<<<static_cast<unsigned>(blocks), kTextWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      text, n_levels, cells, words, wall, goal, lava, start, n_starts, status);
  GU_CHECK_LAUNCH();
  return GU_OK;
}

// ---- batched ASCII render (griduniverse_env.py:202-221) --------------------------------------
// One thread per (env, cell): the cell's glyph with precedence x < G < L < # followed by a blank,
// a newline after the last column of a row, and one more newline closing the frame.
namespace gu {

__global__ void __launch_bounds__(256)
render_ansi_kernel(const uint32_t* __restrict__ wall, const uint32_t* __restrict__ goal,
                   const uint32_t* __restrict__ lava, int per_env, long long n, int X, int cells,
                   const int32_t* __restrict__ pos, uint8_t* __restrict__ text, long long frame) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * cells) return;
  const long long env = i / cells;
  const int c = static_cast<int>(i - env * cells);
  const size_t w = per_env ? static_cast<size_t>(c >> 5) * n + env : (c >> 5);
  const uint32_t bit = 1u << (c & 31);
  uint8_t ch = 'o';
  if (pos[env] == c) ch = 'x';
  if (goal[w] & bit) ch = 'G';
  if (lava[w] & bit) ch = 'L';
  if (wall[w] & bit) ch = '#';
  const int y = c / X, x = c - y * X;
  uint8_t* out = text + env * frame + static_cast<size_t>(y) * (2 * X + 1) + 2 * x;
  out[0] = ch;
  out[1] = ' ';
  if (x == X - 1) out[2] = '\n';
  if (c == cells - 1) out[3] = '\n';
}

}  // namespace gu

extern "C" __attribute__((visibility("default"))) int gu_render_ansi(
    const gu_levels* lv, int64_t n, const int32_t* pos, uint8_t* text, void* stream) {
  if (!lv || !lv->wall || !lv->goal || !lv->lava) return GU_ERR_NULL;
  if (n == 0) return GU_OK;
  if (!pos || !text) return GU_ERR_NULL;
  if (lv->X <= 0 || lv->Y <= 0 || n < 0 || static_cast<int64_t>(lv->X) * lv->Y > (1 << 24)) return GU_ERR_SHAPE;
  const int cells = lv->X * lv->Y;
  const long long frame = static_cast<long long>(lv->Y) * (2 * lv->X + 1) + 1;
  const long long blocks = (n * cells + 255) / 256;
  if (blocks > 2147483647LL) return GU_ERR_SHAPE;
  render_ansi_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      lv->wall, lv->goal, lv->lava, lv->per_env, n, lv->X, cells, pos, text, frame);
  GU_CHECK_LAUNCH();
  return GU_OK;
}


// ---- batched headless RGB render ------------------------------------------------------------------
// The reference's viewer (core/envs/rendering.py:121-135) draws one sprite per cell -- ground, wall,
// goal or lava -- and the agent's face on top; render_policy_arrows (:159-212) adds, for every
// non-terminal non-wall cell and every action with probability >= 0.1, a line of round(p * 20) pixels
// from the tile centre and an arrowhead 10 wide and 5 high at its end (tiles of 32 pixels).  Its
// 'rgb_array' mode is half-wired (griduniverse_env.py:223-230) and needs a GL window; here the same
// picture is rasterised directly, flat colours for the sprites, one thread per pixel, all geometry in
// integer half-pixel units scaled by tile / 32 so that the NumPy twin in the tests agrees bit for bit.
namespace gu {

__device__ __forceinline__ bool arrow_hit(int along, int perp, int L2, int W2, int H2, int T2) {
  if (perp < 0) perp = -perp;
  if (along >= 0 && along <= L2 && perp <= T2) return true;                       // shaft
  return along >= L2 && along <= L2 + H2 && perp * H2 <= W2 * (L2 + H2 - along);  // head
}

__global__ void __launch_bounds__(256)
render_rgb_kernel(const uint32_t* __restrict__ wall, const uint32_t* __restrict__ goal,
                  const uint32_t* __restrict__ lava, int per_env, long long n, int X, int Y,
                  const int32_t* __restrict__ pos, const double* __restrict__ policy, int policy_per_env, int tile,
                  uint8_t* __restrict__ rgb) {
  const long long W = static_cast<long long>(X) * tile, H = static_cast<long long>(Y) * tile;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * W * H) return;
  const long long env = i / (W * H);
  const long long r = i - env * W * H;
  const int py = static_cast<int>(r / W), px = static_cast<int>(r - static_cast<long long>(py) * W);
  const int cx = px / tile, cy = py / tile, c = cy * X + cx;
  const int cells = X * Y;
  const size_t w = per_env ? static_cast<size_t>(c >> 5) * n + env : (c >> 5);
  const uint32_t bit = 1u << (c & 31);
  const bool is_wall = wall[w] & bit, is_lava = lava[w] & bit, is_goal = goal[w] & bit;
  uint8_t R = 200, G = 200, B = 200;                       // ground
  if (is_goal) { R = 40; G = 180; B = 60; }
  if (is_lava) { R = 220; G = 80; B = 20; }
  if (is_wall) { R = 60; G = 60; B = 60; }
  // tile-local coordinates in half pixels, origin at the tile centre, y up
  const int lx = 2 * (px - cx * tile) + 1 - tile, ly = tile - (2 * (py - cy * tile) + 1);
  if (policy != nullptr && !is_wall && !is_lava && !is_goal) {
    const double* row = policy + ((policy_per_env ? env * cells : 0) + c) * 4;
    const int W2 = 5 * tile / 16, H2 = 5 * tile / 16, T2 = tile / 32 > 1 ? tile / 32 : 1;
    bool hit = false;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double p = row[a];
      if (!(p >= 0.1)) continue;                            // rendering.py:174-176
      const int L2 = static_cast<int>(rint(p * 20.0)) * tile / 16;
      const int along = a == 0 ? ly : a == 1 ? lx : a == 2 ? -ly : -lx;
      const int perp = (a & 1) ? ly : lx;
      hit = hit || arrow_hit(along, perp, L2, W2, H2, T2);
    }
    if (hit) { R = 0; G = 0; B = 0; }
  }
  if (pos != nullptr && pos[env] == c && 100 * (lx * lx + ly * ly) <= 49 * tile * tile) { R = 250; G = 210; B = 40; }   // agent
  uint8_t* out = rgb + 3 * i;
  out[0] = R; out[1] = G; out[2] = B;
}

}  // namespace gu

extern "C" __attribute__((visibility("default"))) int gu_render_rgb(
    const gu_levels* lv, int64_t n, const int32_t* pos, const double* policy, int32_t policy_per_env, int32_t tile,
    uint8_t* rgb, void* stream) {
  if (!lv || !lv->wall || !lv->goal || !lv->lava) return GU_ERR_NULL;
  if (n == 0) return GU_OK;
  if (!rgb) return GU_ERR_NULL;
  if (lv->X <= 0 || lv->Y <= 0 || n < 0 || tile < 16 || tile % 16 != 0 || tile > 256) return GU_ERR_SHAPE;
  const long long pixels = n * static_cast<long long>(lv->X) * lv->Y * tile * tile;
  const long long blocks = (pixels + 255) / 256;
  if (blocks > 2147483647LL) return GU_ERR_SHAPE;
  gu::render_rgb_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      lv->wall, lv->goal, lv->lava, lv->per_env, n, lv->X, lv->Y, pos, policy, policy_per_env, tile, rgb);
  GU_CHECK_LAUNCH();
  return GU_OK;
}
