// gu_bfs.cu -- breadth-first distances on the bit-plane grid.
//
// The reference builds an adjacency list with look_step_ahead(s, a, care_about_terminal=False)
// over the non-wall cells (core/algorithms/maze_solving.py:43-50) and runs a FIFO search from one
// state to the first terminal it dequeues (:123-168).  On a 4-connected grid that graph is
// undirected, so one multi-source wavefront from a set of source cells gives every cell its
// distance to the nearest source.  Here a level of the search is a pure bit operation: 32 cells
// per uint32, `visited' = visited | (shifted neighbours & open)`, and only the newly set bits
// write their distance.  Levels ping-pong between two visited planes; HBM/L2-bound integer work.
#include "gu_common.cuh"
#include "gu_cell.cuh"

namespace gu {

constexpr int kBfsThreads = 128;
constexpr int kBfsRows = 8;   // rows walked by one thread (register window over up / cur / down words)

__device__ __forceinline__ uint32_t column_mask(const GridView& g, int w) {
  const int first = w << 5;
  if (first + 32 <= g.X) return 0xffffffffu;
  if (first >= g.X) return 0u;
  return (1u << (g.X - first)) - 1u;
}

// cells the wavefront may enter: not a wall (and not lava when `lava_blocks`), inside the grid
__device__ __forceinline__ uint32_t open_word(const GridView& g, size_t idx, int w, bool lava_blocks) {
  uint32_t blocked = g.wall[idx];
  if (lava_blocks) blocked |= g.lava[idx];
  return ~blocked & column_mask(g, w);
}

// visited plane <- sources on enterable cells; dist <- 0 there, -1 elsewhere (ghost rows and padding too)
__global__ void __launch_bounds__(kBfsThreads)
bfs_init_kernel(GridView g, const uint32_t* __restrict__ sources, uint32_t* __restrict__ visited,
                uint32_t* __restrict__ visited_b, int32_t* __restrict__ dist, bool lava_blocks,
                unsigned long long* __restrict__ reached) {
  const int w = blockIdx.x * kBfsThreads + threadIdx.x;
  const int ar = blockIdx.y;
  const int rows = g.row_end - g.row_begin;
  uint32_t src = 0;
  if (w < g.pitch_words) {
    const size_t idx = static_cast<size_t>(ar) * g.pitch_words + w;
    if (ar >= 1 && ar <= rows) src = (sources ? sources[idx] : g.goal[idx]) & open_word(g, idx, w, lava_blocks);
    visited[idx] = src;
    visited_b[idx] = src;
    int4* d = reinterpret_cast<int4*>(dist + static_cast<size_t>(ar) * g.pitch + (w << 5));
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t nib = src >> (4 * q);
      d[q] = make_int4((nib & 1u) ? 0 : -1, (nib & 2u) ? 0 : -1, (nib & 4u) ? 0 : -1, (nib & 8u) ? 0 : -1);
    }
  }
  unsigned n = __popc(src);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(reached, static_cast<unsigned long long>(n));
}

// one search level: vout = vin | (neighbours(vin) & open); new bits get dist = level
__global__ void __launch_bounds__(kBfsThreads)
bfs_expand_kernel(GridView g, const uint32_t* __restrict__ vin, uint32_t* __restrict__ vout,
                  int32_t* __restrict__ dist, int level, bool lava_blocks,
                  unsigned long long* __restrict__ reached) {
  const int w = blockIdx.x * kBfsThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int rows = g.row_end - g.row_begin;
  const int ar0 = 1 + blockIdx.y * kBfsRows;                    // first array row of this thread
  const int ar1 = min(ar0 + kBfsRows, rows + 1);
  const bool in = w < g.pitch_words;
  const int pw = g.pitch_words;
  auto word = [&](int ar, int ww) -> uint32_t {                // ghost rows hold zeros
    return vin[static_cast<size_t>(ar) * pw + ww];
  };
  uint32_t up = in ? word(ar0 - 1, w) : 0u;
  uint32_t cur = in ? word(ar0, w) : 0u;
  unsigned fresh = 0;
  for (int ar = ar0; ar < ar1; ++ar) {
    const uint32_t dn = in ? word(ar + 1, w) : 0u;
    uint32_t lw = __shfl_up_sync(0xffffffffu, cur, 1);
    uint32_t rw = __shfl_down_sync(0xffffffffu, cur, 1);
    if (lane == 0) lw = (in && w > 0) ? word(ar, w - 1) : 0u;
    if (lane == 31) rw = (w + 1 < pw) ? word(ar, w + 1) : 0u;
    if (in) {
      const size_t idx = static_cast<size_t>(ar) * pw + w;
      const uint32_t nb = up | dn | (cur << 1) | (lw >> 31) | (cur >> 1) | (rw << 31);
      uint32_t nw = 0;
      if (nb & ~cur) nw = nb & ~cur & open_word(g, idx, w, lava_blocks);
      vout[idx] = cur | nw;
      if (nw) {
        fresh += __popc(nw);
        int32_t* d = dist + static_cast<size_t>(ar) * g.pitch + (w << 5);
        do {
          const int b = __ffs(nw) - 1;
          nw &= nw - 1;
          d[b] = level;
        } while (nw);
      }
    }
    up = cur;
    cur = dn;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fresh += __shfl_xor_sync(0xffffffffu, fresh, o);
  if (lane == 0 && fresh) atomicAdd(reached, static_cast<unsigned long long>(fresh));
}

// Follow the distance field downhill from one cell: at every step the lowest-numbered action
// (UP < RIGHT < DOWN < LEFT, the np.argmax order of examples/griduniverse_alg_examples.py:76) whose
// landing cell is one level closer.  One thread; a path is a strictly serial object.
__global__ void bfs_walk_kernel(GridView g, const int32_t* __restrict__ dist, int start, int8_t* __restrict__ actions,
                                int max_len, int32_t* __restrict__ len_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int x = start % g.X, y = start / g.X;
  auto at = [&](int xx, int yy) -> int {
    return dist[static_cast<size_t>(yy - g.row_begin + 1) * g.pitch + xx];
  };
  int d = at(x, y);
  if (d < 0) { *len_out = -1; return; }
  int n = 0;
  while (d > 0 && n < max_len) {
    int a = -1;
    if (y > 0 && at(x, y - 1) == d - 1) { a = 0; --y; }
    else if (x < g.X - 1 && at(x + 1, y) == d - 1) { a = 1; ++x; }
    else if (y < g.Y - 1 && at(x, y + 1) == d - 1) { a = 2; ++y; }
    else if (x > 0 && at(x - 1, y) == d - 1) { a = 3; --x; }
    if (a < 0) { *len_out = -2; return; }     // not a distance field of this grid
    actions[n++] = static_cast<int8_t>(a);
    --d;
  }
  *len_out = d == 0 ? n : -3;                // -3: max_len too small
}

static inline GridView bview(const gu_grid* g) {
  GridView v;
  v.X = g->X; v.Y = g->Y; v.row_begin = g->row_begin; v.row_end = g->row_end;
  v.pitch = g->pitch; v.pitch_words = g->pitch_words;
  v.wall = g->wall; v.goal = g->goal; v.lava = g->lava;
  return v;
}

static inline int bfs_args_ok(const gu_grid* g) {
  if (!g || !g->wall || !g->goal || !g->lava) return GU_ERR_NULL;
  if (g->X <= 0 || g->Y <= 0 || g->row_end <= g->row_begin) return GU_ERR_SHAPE;
  if (g->row_begin != 0 || g->row_end != g->Y) return GU_ERR_UNSUPPORTED;     // whole grids only
  if (g->pitch != g->pitch_words * 32 || g->pitch < g->X) return GU_ERR_ALIGN;
  if ((g->Y + kBfsRows - 1) / kBfsRows > 65535 || g->Y + 2 > 65535) return GU_ERR_SHAPE;
  return GU_OK;
}

}  // namespace gu

using namespace gu;

extern "C" __attribute__((visibility("default"))) int gu_bfs_init(
    const gu_grid* g, const uint32_t* sources, uint32_t* visited_a, uint32_t* visited_b, int32_t* dist,
    uint64_t* reached, uint32_t flags, void* stream) {
  const int rc = bfs_args_ok(g);
  if (rc != GU_OK) return rc;
  if (!visited_a || !visited_b || !dist || !reached) return GU_ERR_NULL;
  if (flags & ~static_cast<uint32_t>(GU_BFS_LAVA_BLOCKS)) return GU_ERR_MODE;
  if (reinterpret_cast<uintptr_t>(dist) & 15u) return GU_ERR_ALIGN;
  dim3 grid((g->pitch_words + kBfsThreads - 1) / kBfsThreads, g->Y + 2);
  bfs_init_kernel<<<grid, kBfsThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      bview(g), sources, visited_a, visited_b, dist, (flags & GU_BFS_LAVA_BLOCKS) != 0,
      reinterpret_cast<unsigned long long*>(reached));
  GU_CHECK_LAUNCH();
  return GU_OK;
}

extern "C" __attribute__((visibility("default"))) int gu_bfs_expand(
    const gu_grid* g, uint32_t* visited_a, uint32_t* visited_b, int32_t* dist, int32_t level_begin,
    int32_t n_levels, uint64_t* reached, uint32_t flags, void* stream) {
  const int rc = bfs_args_ok(g);
  if (rc != GU_OK) return rc;
  if (!visited_a || !visited_b || !dist || !reached) return GU_ERR_NULL;
  if (flags & ~static_cast<uint32_t>(GU_BFS_LAVA_BLOCKS)) return GU_ERR_MODE;
  if (level_begin < 1 || n_levels < 0) return GU_ERR_SHAPE;
  dim3 grid((g->pitch_words + kBfsThreads - 1) / kBfsThreads, (g->Y + kBfsRows - 1) / kBfsRows);
  const GridView v = bview(g);
  for (int32_t k = 0; k < n_levels; ++k) {
    const int32_t level = level_begin + k;
    // level L reads plane (L-1)&1 and writes plane L&1; plane 0 is visited_a
    uint32_t* vin = ((level - 1) & 1) ? visited_b : visited_a;
    uint32_t* vout = (level & 1) ? visited_b : visited_a;
    bfs_expand_kernel<<<grid, kBfsThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        v, vin, vout, dist, level, (flags & GU_BFS_LAVA_BLOCKS) != 0, reinterpret_cast<unsigned long long*>(reached));
  }
  GU_CHECK_LAUNCH();
  return GU_OK;
}

extern "C" __attribute__((visibility("default"))) int gu_bfs_walk(
    const gu_grid* g, const int32_t* dist, int64_t start_state, int8_t* actions, int32_t max_len,
    int32_t* length, void* stream) {
  const int rc = bfs_args_ok(g);
  if (rc != GU_OK) return rc;
  if (!dist || !length || (max_len > 0 && !actions)) return GU_ERR_NULL;
  if (start_state < 0 || start_state >= static_cast<int64_t>(g->X) * g->Y || max_len < 0) return GU_ERR_SHAPE;
  bfs_walk_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(bview(g), dist, static_cast<int>(start_state),
                                                                actions, max_len, length);
  GU_CHECK_LAUNCH();
  return GU_OK;
}
