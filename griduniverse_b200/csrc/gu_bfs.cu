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
#ifndef GU_BFS_ROWS
#define GU_BFS_ROWS 4
#endif
constexpr int kBfsRows = GU_BFS_ROWS;   // rows walked by one thread (register window over up / cur / down quads)

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
  if (w >= (g.pitch >> 5) && w < g.pitch_words) {   // row padding words: stay zero, no dist storage
    visited[static_cast<size_t>(ar) * g.pitch_words + w] = 0u;
    visited_b[static_cast<size_t>(ar) * g.pitch_words + w] = 0u;
  }
  if (w < (g.pitch >> 5)) {
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

// one search level: vout = vin | (neighbours(vin) & open); new bits get dist = level.
// A thread owns a quad of 4 words (128 cells, one 16-byte load) and walks kBfsRows rows with the
// up / cur / down quads in registers; the words left and right of the quad come from the
// neighbouring lanes.  Words at and beyond pitch/32 are row padding: open_word() is 0 there.
__global__ void __launch_bounds__(kBfsThreads)
bfs_expand_kernel(GridView g, const uint32_t* __restrict__ vin, uint32_t* __restrict__ vout,
                  int32_t* __restrict__ dist, int level, bool lava_blocks,
                  unsigned long long* __restrict__ reached) {
  const int q = blockIdx.x * kBfsThreads + threadIdx.x;        // quad column
  const int lane = threadIdx.x & 31;
  const int rows = g.row_end - g.row_begin;
  const int ar0 = 1 + blockIdx.y * kBfsRows;                    // first array row of this thread
  const int pw = g.pitch_words, nq = pw >> 2;
  const bool in = q < nq;
  const bool edge_l = in && lane == 0 && q > 0, edge_r = lane == 31 && q + 1 < nq;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  // running pointers: row ar of the input plane (quad view / word view), output plane, distances
  const uint32_t* pin = vin + static_cast<size_t>(ar0) * pw + 4 * q;
  uint32_t* pout = vout + static_cast<size_t>(ar0) * pw + 4 * q;
  int32_t* pd = dist + static_cast<size_t>(ar0) * g.pitch + (q << 7);
  const uint32_t* pwall = g.wall + static_cast<size_t>(ar0) * pw + 4 * q;
  const uint32_t* plava = g.lava + static_cast<size_t>(ar0) * pw + 4 * q;
  uint4 up = in ? *reinterpret_cast<const uint4*>(pin - pw) : zero;   // ghost rows hold zeros
  uint4 cur = in ? *reinterpret_cast<const uint4*>(pin) : zero;
  unsigned fresh = 0;
#pragma unroll
  for (int k = 0; k < kBfsRows; ++k) {
    if (ar0 + k <= rows) {                                       // uniform across the block
      const uint4 dn = in ? *reinterpret_cast<const uint4*>(pin + pw) : zero;
      uint32_t lw = __shfl_up_sync(0xffffffffu, cur.w, 1);
      uint32_t rw = __shfl_down_sync(0xffffffffu, cur.x, 1);
      if (lane == 0) lw = edge_l ? pin[-1] : 0u;
      if (lane == 31) rw = edge_r ? pin[4] : 0u;
      if (in) {
        uint32_t nw[4];
        nw[0] = (up.x | dn.x | (cur.x << 1) | (lw >> 31) | (cur.x >> 1) | (cur.y << 31)) & ~cur.x;
        nw[1] = (up.y | dn.y | (cur.y << 1) | (cur.x >> 31) | (cur.y >> 1) | (cur.z << 31)) & ~cur.y;
        nw[2] = (up.z | dn.z | (cur.z << 1) | (cur.y >> 31) | (cur.z >> 1) | (cur.w << 31)) & ~cur.z;
        nw[3] = (up.w | dn.w | (cur.w << 1) | (cur.z >> 31) | (cur.w >> 1) | (rw << 31)) & ~cur.w;
        if (nw[0] | nw[1] | nw[2] | nw[3]) {                    // candidates: now look at the masks
          const uint4 wl = *reinterpret_cast<const uint4*>(pwall);
          uint4 bl = wl;
          if (lava_blocks) {
            const uint4 lv = *reinterpret_cast<const uint4*>(plava);
            bl = make_uint4(wl.x | lv.x, wl.y | lv.y, wl.z | lv.z, wl.w | lv.w);
          }
          const uint32_t blocked[4] = {bl.x, bl.y, bl.z, bl.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t m = nw[c] & ~blocked[c] & column_mask(g, 4 * q + c);
            nw[c] = m;
            if (m) {
              fresh += __popc(m);
              int32_t* d = pd + (c << 5);
              do {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                d[b] = level;
              } while (m);
            }
          }
        }
        *reinterpret_cast<uint4*>(pout) = make_uint4(cur.x | nw[0], cur.y | nw[1], cur.z | nw[2], cur.w | nw[3]);
      }
      up = cur;
      cur = dn;
      pin += pw; pout += pw; pwall += pw; plava += pw;
      pd += g.pitch;
    }
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
  if (g->pitch % 32 != 0 || g->pitch > g->pitch_words * 32 || g->pitch < g->X) return GU_ERR_ALIGN;
  if (g->pitch_words % 4 != 0) return GU_ERR_ALIGN;                            // rows move as 16-byte quads
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
  if ((reinterpret_cast<uintptr_t>(visited_a) | reinterpret_cast<uintptr_t>(visited_b)) & 15u) return GU_ERR_ALIGN;
  dim3 grid((g->pitch_words / 4 + kBfsThreads - 1) / kBfsThreads, (g->Y + kBfsRows - 1) / kBfsRows);
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
