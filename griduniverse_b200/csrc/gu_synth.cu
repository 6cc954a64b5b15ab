// gu_synth.cu -- seedable synthetic level generators, evaluated on the device.
//
// Not part of the reference (its only generator is the serial recursive backtracker in
// core/envs/maze_generation.py).  These are pure functions of (seed, index) so the same
// levels can be produced on the host (griduniverse_b200/synth.py is the NumPy twin used by
// the parity tests) and on any shard of a multi-GPU run.
#include "gu_common.cuh"

namespace gu {

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ uint32_t hash3(uint32_t seed, uint32_t a, uint32_t b) {
  uint32_t h = fmix32(seed * 0x9E3779B1u + 0x7F4A7C15u);
  h = fmix32(h ^ a);
  h = fmix32((h + 0x165667B1u) ^ b);
  return h;
}

constexpr uint32_t kWallThreshold = 858993459u;   // 0.2  * 2^32
constexpr uint32_t kMazeThreshold = 1073741824u;  // 0.25 * 2^32
constexpr uint32_t kLavaThreshold = 4294967u;     // 0.001 * 2^32

// One thread per env: border open, 20 % interior walls, one goal, cells/32 lava draws,
// one start on an open non-terminal cell.  Planes are WORD-MAJOR uint32[words][N].
__global__ void __launch_bounds__(128)
synth_env_levels_kernel(int X, int Y, int words, int64_t N, int64_t env0, uint32_t seed,
                        uint32_t* __restrict__ wall, uint32_t* __restrict__ goal,
                        uint32_t* __restrict__ lava, int32_t* __restrict__ start) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t env = static_cast<uint32_t>(env0 + i);
  const int cells = X * Y;
  constexpr int kMaxWords = 8;   // cells <= 256
  uint32_t w[kMaxWords], l[kMaxWords];
#pragma unroll
  for (int k = 0; k < kMaxWords; ++k) { w[k] = 0; l[k] = 0; }
  for (int c = 0; c < cells; ++c) {
    const int y = c / X, x = c - y * X;
    const bool border = x == 0 || y == 0 || x == X - 1 || y == Y - 1;
    const bool is_wall = !border && hash3(seed, env, static_cast<uint32_t>(c)) < kWallThreshold;
#pragma unroll
    for (int k = 0; k < kMaxWords; ++k)
      if (k == (c >> 5) && is_wall) w[k] |= 1u << (c & 31);
  }
  auto bit = [&](const uint32_t (&p)[kMaxWords], int c) -> bool {
    uint32_t word = 0;
#pragma unroll
    for (int k = 0; k < kMaxWords; ++k)
      if (k == (c >> 5)) word = p[k];
    return (word >> (c & 31)) & 1u;
  };
  int g = static_cast<int>(hash3(seed, env, 0x10000u) % static_cast<uint32_t>(cells));
  while (bit(w, g)) g = (g + 1 == cells) ? 0 : g + 1;
  const int n_lava = cells / 32;
  for (int k = 0; k < n_lava; ++k) {
    const int c = static_cast<int>(hash3(seed, env, 0x20000u + k) % static_cast<uint32_t>(cells));
    if (!bit(w, c) && c != g) {
#pragma unroll
      for (int q = 0; q < kMaxWords; ++q)
        if (q == (c >> 5)) l[q] |= 1u << (c & 31);
    }
  }
  int s = static_cast<int>(hash3(seed, env, 0x30000u) % static_cast<uint32_t>(cells));
  while (bit(w, s) || bit(l, s) || s == g) s = (s + 1 == cells) ? 0 : s + 1;
#pragma unroll
  for (int k = 0; k < kMaxWords; ++k) {
    if (k < words) {
      wall[static_cast<int64_t>(k) * N + i] = w[k];
      lava[static_cast<int64_t>(k) * N + i] = l[k];
      goal[static_cast<int64_t>(k) * N + i] = (k == (g >> 5)) ? (1u << (g & 31)) : 0u;
    }
  }
  start[i] = s;
}

// Row-pitched planes of the cfg-5 style maze for rows [row_begin-1, row_end+1) (ghost rows
// included, zero outside the grid).  One thread per 32-cell word.
__global__ void __launch_bounds__(256)
synth_maze_kernel(int X, int Y, int row_begin, int row_end, int pitch_words, uint32_t seed,
                  uint32_t* __restrict__ wall, uint32_t* __restrict__ goal, uint32_t* __restrict__ lava) {
  const int wx = blockIdx.x * blockDim.x + threadIdx.x;
  const int ar = blockIdx.y;                      // array row
  if (wx >= pitch_words) return;
  const int y = row_begin - 1 + ar;
  uint32_t w = 0, g = 0, l = 0;
  if (y >= 0 && y < Y) {
    const int gx = (X / 2) & ~1, gy = (Y / 2) & ~1;   // the (even, even) cell next to the centre
    for (int b = 0; b < 32; ++b) {
      const int x = wx * 32 + b;
      if (x >= X) break;
      const bool xo = x & 1, yo = y & 1;
      bool is_wall = (xo && yo) ||
                     ((xo != yo) && hash3(seed, static_cast<uint32_t>(y), static_cast<uint32_t>(x)) < kMazeThreshold);
      const bool is_goal = (x == gx && y == gy);
      const bool is_lava = !is_wall && !is_goal &&
                           hash3(seed, static_cast<uint32_t>(y), static_cast<uint32_t>(x) | 0x80000000u) < kLavaThreshold;
      w |= static_cast<uint32_t>(is_wall) << b;
      g |= static_cast<uint32_t>(is_goal) << b;
      l |= static_cast<uint32_t>(is_lava) << b;
    }
  }
  const size_t o = static_cast<size_t>(ar) * pitch_words + wx;
  wall[o] = w; goal[o] = g; lava[o] = l;
}

}  // namespace gu

using namespace gu;

extern "C" __attribute__((visibility("default"))) int gu_synth_env_levels(
    int32_t X, int32_t Y, int64_t n_envs, int64_t first_env, uint32_t seed, uint32_t* wall, uint32_t* goal,
    uint32_t* lava, int32_t* start, void* stream) {
  if (!wall || !goal || !lava || !start) return GU_ERR_NULL;
  const int64_t cells = static_cast<int64_t>(X) * Y;
  if (X < 3 || Y < 3 || cells > 256 || n_envs < 0) return GU_ERR_SHAPE;
  if (n_envs == 0) return GU_OK;
  const int words = static_cast<int>((cells + 31) / 32);
  synth_env_levels_kernel<<<static_cast<unsigned>((n_envs + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      X, Y, words, n_envs, first_env, seed, wall, goal, lava, start);
  GU_CHECK_LAUNCH();
  return GU_OK;
}

extern "C" __attribute__((visibility("default"))) int gu_synth_maze(
    int32_t X, int32_t Y, int32_t row_begin, int32_t row_end, int32_t pitch_words, uint32_t seed,
    uint32_t* wall, uint32_t* goal, uint32_t* lava, void* stream) {
  if (!wall || !goal || !lava) return GU_ERR_NULL;
  if (X <= 0 || Y <= 0 || row_begin < 0 || row_end > Y || row_begin >= row_end || pitch_words * 32 < X)
    return GU_ERR_SHAPE;
  const int rows = row_end - row_begin + 2;
  if (rows > 65535) return GU_ERR_SHAPE;
  dim3 grid((pitch_words + 255) / 256, rows);
  synth_maze_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(X, Y, row_begin, row_end, pitch_words,
                                                                          seed, wall, goal, lava);
  GU_CHECK_LAUNCH();
  return GU_OK;
}
