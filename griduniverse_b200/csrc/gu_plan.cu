// gu_plan.cu -- Bellman sweep, greedy extraction and the single-block VI loop (sm_100a).
//
// Reference: core/algorithms/utils.py:15-27,55-72 and dynamic_programming.py:8-28.
// This file holds the layout-agnostic kernels (any X / pitch); the TMA-tiled kernels for
// large aligned grids live in gu_plan_tiled.cu and share gu_cell.cuh's arithmetic.
#include "gu_cell.cuh"

namespace gu {

static inline GridView view_of(const gu_grid* g) {
  GridView v;
  v.X = g->X; v.Y = g->Y; v.row_begin = g->row_begin; v.row_end = g->row_end;
  v.pitch = g->pitch; v.pitch_words = g->pitch_words;
  v.wall = g->wall; v.goal = g->goal; v.lava = g->lava;
  return v;
}

int check_grid(const gu_grid* g) {
  if (!g || !g->wall || !g->goal || !g->lava) return GU_ERR_NULL;
  if (g->X <= 0 || g->Y <= 0 || g->row_begin < 0 || g->row_end > g->Y || g->row_begin >= g->row_end)
    return GU_ERR_SHAPE;
  if (g->pitch < g->X || g->pitch_words * 32 < g->X) return GU_ERR_SHAPE;
  return GU_OK;
}

// ---- generic sweep: one thread per cell, 32x8 cells per block -----------------------
template <typename T, int KIND>
__global__ void __launch_bounds__(256)
sweep_generic_kernel(GridView g, const T* __restrict__ vin, T* __restrict__ vout,
                     const void* __restrict__ policy, T gamma, T* residual, const T* gate, T gate_thr) {
  __shared__ T scratch[8];
  if (gate != nullptr && *gate < gate_thr) return;   // converged earlier: keep V where it is
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = g.row_begin + blockIdx.y * 8 + threadIdx.y;
  T delta = Num<T>::neg_inf();
  if (x < g.X && y < g.row_end) {
    CellIn<T> c;
    gather_cell(g, vin, x, y, c);
    const size_t cell = static_cast<size_t>(y - g.row_begin + 1) * g.pitch + x;
    const T vnew = cell_update<T, KIND>(c, gamma, policy, cell);
    vout[cell] = vnew;
    delta = Num<T>::add(c.vs, -vnew);   // V - V_new, signed (dynamic_programming.py:17)
  }
  block_max_to_global(delta, scratch, residual);
}

template <typename T>
__global__ void __launch_bounds__(256)
greedy_generic_kernel(GridView g, const T* __restrict__ v, uint8_t* __restrict__ tie, T gamma) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = g.row_begin + blockIdx.y * 8 + threadIdx.y;
  if (x < g.X && y < g.row_end) {
    CellIn<T> c;
    gather_cell(g, v, x, y, c);
    T gn[4];
    discounted_next(c, gamma, gn);
    tie[static_cast<size_t>(y - g.row_begin + 1) * g.pitch + x] = static_cast<uint8_t>(tie_mask_of(c, gn));
  }
}

template <typename T>
int sweep_generic(const gu_grid* g, const T* vin, T* vout, int kind, const void* policy, T gamma,
                  T* residual, const T* gate, T gate_thr, cudaStream_t st) {
  const GridView v = view_of(g);
  const int rows = g->row_end - g->row_begin;
  dim3 block(32, 8), grid((g->X + 31) / 32, (rows + 7) / 8);
  if (grid.y > 65535u) return GU_ERR_SHAPE;
  switch (kind) {
    case GU_POLICY_PROBS:
      sweep_generic_kernel<T, GU_POLICY_PROBS><<<grid, block, 0, st>>>(v, vin, vout, policy, gamma, residual, gate, gate_thr);
      break;
    case GU_POLICY_MASK:
      sweep_generic_kernel<T, GU_POLICY_MASK><<<grid, block, 0, st>>>(v, vin, vout, policy, gamma, residual, gate, gate_thr);
      break;
    case GU_POLICY_UNIFORM:
      sweep_generic_kernel<T, GU_POLICY_UNIFORM><<<grid, block, 0, st>>>(v, vin, vout, policy, gamma, residual, gate, gate_thr);
      break;
    case GU_POLICY_GREEDY:
      sweep_generic_kernel<T, GU_POLICY_GREEDY><<<grid, block, 0, st>>>(v, vin, vout, policy, gamma, residual, gate, gate_thr);
      break;
    default:
      return GU_ERR_MODE;
  }
  GU_CHECK_LAUNCH();
  return GU_OK;
}

template <typename T>
int greedy_generic(const gu_grid* g, const T* v, uint8_t* tie, T gamma, cudaStream_t st) {
  const GridView gv = view_of(g);
  const int rows = g->row_end - g->row_begin;
  dim3 block(32, 8), grid((g->X + 31) / 32, (rows + 7) / 8);
  if (grid.y > 65535u) return GU_ERR_SHAPE;
  greedy_generic_kernel<T><<<grid, block, 0, st>>>(gv, v, tie, gamma);
  GU_CHECK_LAUNCH();
  return GU_OK;
}

// Tiled fast paths (gu_plan_tiled.cu); return GU_ERR_UNSUPPORTED when the shape does not qualify.
int sweep_tiled_f32(const gu_grid*, const float*, float*, int, const void*, float, float*, const float*,
                    float, cudaStream_t);
int sweep_tiled_f64(const gu_grid*, const double*, double*, int, const void*, double, double*, const double*,
                    double, cudaStream_t);
int greedy_tiled_f32(const gu_grid*, const float*, uint8_t*, float, cudaStream_t);
int greedy_tiled_f64(const gu_grid*, const double*, uint8_t*, double, cudaStream_t);

// ---- whole VI loop in one thread block (grids that fit in shared memory) ---------------
constexpr int64_t kSmallMaxCells = 13000;   // 2 x f64 + 1 info byte per cell within 227 KB

// A batch of same-shape mazes, one thread block each (blockIdx.x = maze): maze m's planes start
// plane_stride words, its per-cell arrays cell_stride elements after maze m-1's (0 for a single grid).
struct BatchStrides {
  long long plane, cell;
};
__device__ __forceinline__ GridView maze_view(GridView g, const BatchStrides& b) {
  const long long o = static_cast<long long>(blockIdx.x) * b.plane;
  g.wall += o; g.goal += o; g.lava += o;
  return g;
}

// info byte: bits 0-3 = blocked per action, bit 4 = goal, bit 5 = lava
__device__ __forceinline__ void small_cell(const double* va, const uint8_t* info, int s, int X,
                                           CellIn<double>& c) {
  const uint32_t i = info[s];
  c.vs = va[s];
  c.blk = i & 15u;
  c.term = (i & 0x30u) != 0;
  c.rs = reward_of(i & 0x10u, i & 0x20u);
  const int off[4] = {-X, 1, X, -1};
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (!((c.blk >> a) & 1u)) {
      const int n = s + off[a];
      const uint32_t j = info[n];
      c.vn[a] = va[n];
      c.rn[a] = reward_of(j & 0x10u, j & 0x20u);
    } else {
      c.vn[a] = c.vs;
      c.rn[a] = c.rs;
    }
  }
}

template <int kSmallThreads>
__global__ void __launch_bounds__(kSmallThreads)
vi_small_kernel(GridView g0, BatchStrides bs, const double* __restrict__ v0, double* __restrict__ vout,
                uint8_t* __restrict__ tie, int kind0, const void* __restrict__ policy, double gamma,
                double threshold, int max_steps, int32_t* sweeps_out, double* last_delta) {
  extern __shared__ double smem_d[];
  const GridView g = maze_view(g0, bs);
  {                                             // this maze's slice of every per-cell array
    const long long o = static_cast<long long>(blockIdx.x) * bs.cell;
    if (v0 != nullptr) v0 += o;
    vout += o;
    tie += o;
    if (policy != nullptr)
      policy = static_cast<const char*>(policy) + o * (kind0 == GU_POLICY_PROBS ? 4 * sizeof(double) : 1);
    sweeps_out += blockIdx.x;
    last_delta += blockIdx.x;
  }
  const int N = g.X * g.Y;
  double* va = smem_d;
  double* vb = smem_d + N;
  uint8_t* info = reinterpret_cast<uint8_t*>(vb + N);
  __shared__ double red[32];
  __shared__ double delta_sh;
  const int tid = threadIdx.x;

  for (int s = tid; s < N; s += kSmallThreads) {
    const int y = s / g.X, x = s - y * g.X;
    CellIn<double> c;
    gather_cell(g, v0, x, y, c);
    const size_t w = static_cast<size_t>(y + 1) * g.pitch_words + (x >> 5);
    const uint32_t goal = (g.goal[w] >> (x & 31)) & 1u, lava = (g.lava[w] >> (x & 31)) & 1u;
    info[s] = static_cast<uint8_t>(c.blk | (goal << 4) | (lava << 5));
    va[s] = c.vs;
  }
  __syncthreads();

  int sweeps = 0;
  double delta = 0.0;
  for (int it = 0; it < max_steps; ++it) {
    double dmax = -CUDART_INF;
    for (int s = tid; s < N; s += kSmallThreads) {
      CellIn<double> c;
      small_cell(va, info, s, g.X, c);
      const int y = s / g.X, x = s - y * g.X;
      const size_t cell = static_cast<size_t>(y + 1) * g.pitch + x;
      double vnew;
      if (it == 0 && kind0 == GU_POLICY_PROBS) vnew = cell_update<double, GU_POLICY_PROBS>(c, gamma, policy, cell);
      else if (it == 0 && kind0 == GU_POLICY_MASK) vnew = cell_update<double, GU_POLICY_MASK>(c, gamma, policy, cell);
      else if (it == 0 && kind0 == GU_POLICY_UNIFORM) vnew = cell_update<double, GU_POLICY_UNIFORM>(c, gamma, policy, cell);
      else vnew = cell_update<double, GU_POLICY_GREEDY>(c, gamma, policy, cell);
      vb[s] = vnew;
      const double d = __dadd_rn(c.vs, -vnew);
      dmax = d > dmax ? d : dmax;
    }
    dmax = warp_max(dmax);
    if ((tid & 31) == 0) red[tid >> 5] = dmax;
    __syncthreads();
    if (tid < 32) {
      double w = warp_max(tid < kSmallThreads / 32 ? red[tid] : -CUDART_INF);
      if (tid == 0) delta_sh = w;
    }
    __syncthreads();
    delta = delta_sh;
    double* t = va; va = vb; vb = t;
    ++sweeps;
    if (delta < threshold) break;   // dynamic_programming.py:22-23 (uniform across the block)
  }

  for (int s = tid; s < N; s += kSmallThreads) {
    CellIn<double> c;
    small_cell(va, info, s, g.X, c);
    double gn[4];
    discounted_next(c, gamma, gn);
    const int y = s / g.X, x = s - y * g.X;
    const size_t cell = static_cast<size_t>(y + 1) * g.pitch + x;
    vout[cell] = c.vs;
    tie[cell] = static_cast<uint8_t>(tie_mask_of(c, gn));
  }
  if (tid == 0) {
    *sweeps_out = sweeps;
    *last_delta = delta;
  }
}

// ---- whole policy_iteration loop (dynamic_programming.py:31-57) in one thread block ----------
// Shared memory: three value arrays (current, scratch, last converged), the info bytes and the
// tie masks of the current greedy policy.  The caller's policy (kind0) is evaluated until the
// first improvement; from then on the policy is the tie masks held in shared memory.
constexpr int64_t kPiSmallMaxCells = 8500;   // 3 x f64 + 2 bytes per cell within 227 KB

__device__ __forceinline__ double block_max(double v, double* red, double* out) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    const double w = warp_max(threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -CUDART_INF);
    if (threadIdx.x == 0) *out = w;
  }
  __syncthreads();
  return *out;
}

template <int kSmallThreads>
__global__ void __launch_bounds__(kSmallThreads)
pi_small_kernel(GridView g0, BatchStrides bs, const double* __restrict__ v0, double* __restrict__ vout,
                uint8_t* __restrict__ tie, int kind0, const void* __restrict__ policy, double gamma,
                double threshold, int max_steps, int32_t* meta, double* last_delta_eval) {
  extern __shared__ double smem_d[];
  const GridView g = maze_view(g0, bs);
  {
    const long long o = static_cast<long long>(blockIdx.x) * bs.cell;
    if (v0 != nullptr) v0 += o;
    vout += o;
    tie += o;
    if (policy != nullptr)
      policy = static_cast<const char*>(policy) + o * (kind0 == GU_POLICY_PROBS ? 4 * sizeof(double) : 1);
    meta += 3 * blockIdx.x;
    last_delta_eval += blockIdx.x;
  }
  const int N = g.X * g.Y;
  double* v = smem_d;               // current value function
  double* scratch = smem_d + N;
  double* last = smem_d + 2 * N;    // last converged value function (:36,45)
  uint8_t* info = reinterpret_cast<uint8_t*>(smem_d + 3 * N);
  uint8_t* pmask = info + N;
  __shared__ double red[32];
  __shared__ double red_out;
  const int tid = threadIdx.x;

  for (int s = tid; s < N; s += kSmallThreads) {
    const int y = s / g.X, x = s - y * g.X;
    CellIn<double> c;
    gather_cell(g, v0, x, y, c);
    const size_t w = static_cast<size_t>(y + 1) * g.pitch_words + (x >> 5);
    const uint32_t goal = (g.goal[w] >> (x & 31)) & 1u, lava = (g.lava[w] >> (x & 31)) & 1u;
    info[s] = static_cast<uint8_t>(c.blk | (goal << 4) | (lava << 5));
    v[s] = c.vs;
    last[s] = c.vs;
  }
  __syncthreads();

  auto greedy_of = [&](const double* from) {             // utils.py:55-72 on `from` -> pmask
    for (int s = tid; s < N; s += kSmallThreads) {
      CellIn<double> c;
      small_cell(from, info, s, g.X, c);
      double gn[4];
      discounted_next(c, gamma, gn);
      pmask[s] = static_cast<uint8_t>(tie_mask_of(c, gn));
    }
    __syncthreads();
  };

  int sweeps = 0, improved = 0, exhausted = 0;
  int kind = kind0;
  double delta_eval = 0.0;
  for (int step = 0; step < max_steps; ++step) {
    double dmax = -CUDART_INF;
    for (int s = tid; s < N; s += kSmallThreads) {
      CellIn<double> c;
      small_cell(v, info, s, g.X, c);
      const int y = s / g.X, x = s - y * g.X;
      const size_t cell = static_cast<size_t>(y + 1) * g.pitch + x;
      double vnew;
      if (improved) {
        double gn[4];
        discounted_next(c, gamma, gn);
        vnew = backup_mask(c, gn, pmask[s]);
      } else if (kind == GU_POLICY_PROBS) vnew = cell_update<double, GU_POLICY_PROBS>(c, gamma, policy, cell);
      else if (kind == GU_POLICY_MASK) vnew = cell_update<double, GU_POLICY_MASK>(c, gamma, policy, cell);
      else if (kind == GU_POLICY_UNIFORM) vnew = cell_update<double, GU_POLICY_UNIFORM>(c, gamma, policy, cell);
      else vnew = cell_update<double, GU_POLICY_GREEDY>(c, gamma, policy, cell);
      scratch[s] = vnew;
      const double d = __dadd_rn(c.vs, -vnew);
      dmax = d > dmax ? d : dmax;
    }
    delta_eval = block_max(dmax, red, &red_out);          // np.max(v - v_new), :40
    { double* t = v; v = scratch; scratch = t; }
    ++sweeps;
    if (delta_eval < threshold) {                          // evaluation converged, :42
      greedy_of(v);                                        // :43 (in place, utils.py:69)
      improved = 1;
      double lmax = -CUDART_INF;
      for (int s = tid; s < N; s += kSmallThreads) {
        const double d = __dadd_rn(last[s], -v[s]);
        lmax = d > lmax ? d : lmax;
        last[s] = v[s];
      }
      const double delta = block_max(lmax, red, &red_out); // np.max(last_converged - v_new), :44
      if (delta < threshold) break;                        // :46-47
    } else if (step == max_steps - 1) {                    // :48-54
      greedy_of(last);
      improved = 1;
      exhausted = 1;
    }
  }

  for (int s = tid; s < N; s += kSmallThreads) {
    const int y = s / g.X, x = s - y * g.X;
    const size_t cell = static_cast<size_t>(y + 1) * g.pitch + x;
    vout[cell] = last[s];
    if (improved) tie[cell] = pmask[s];
  }
  if (tid == 0) {
    meta[0] = sweeps;
    meta[1] = improved;
    meta[2] = exhausted;
    *last_delta_eval = delta_eval;
  }
}

}  // namespace gu

using namespace gu;

extern "C" __attribute__((visibility("default"))) int gu_sweep_f64(const gu_grid* g, const double* v_in, double* v_out, int policy_kind,
                            const void* policy, double gamma, double* residual, const double* gate,
                            double gate_threshold, void* stream) {
  int rc = check_grid(g);
  if (rc) return rc;
  if (!v_in || !v_out) return GU_ERR_NULL;
  if ((policy_kind == GU_POLICY_PROBS || policy_kind == GU_POLICY_MASK) && !policy) return GU_ERR_NULL;
  rc = sweep_tiled_f64(g, v_in, v_out, policy_kind, policy, gamma, residual, gate, gate_threshold,
                       static_cast<cudaStream_t>(stream));
  if (rc != GU_ERR_UNSUPPORTED) return rc;
  return sweep_generic<double>(g, v_in, v_out, policy_kind, policy, gamma, residual, gate, gate_threshold,
                               static_cast<cudaStream_t>(stream));
}

extern "C" __attribute__((visibility("default"))) int gu_sweep_f32(const gu_grid* g, const float* v_in, float* v_out, int policy_kind,
                            const void* policy, float gamma, float* residual, const float* gate,
                            float gate_threshold, void* stream) {
  int rc = check_grid(g);
  if (rc) return rc;
  if (!v_in || !v_out) return GU_ERR_NULL;
  if ((policy_kind == GU_POLICY_PROBS || policy_kind == GU_POLICY_MASK) && !policy) return GU_ERR_NULL;
  rc = sweep_tiled_f32(g, v_in, v_out, policy_kind, policy, gamma, residual, gate, gate_threshold,
                       static_cast<cudaStream_t>(stream));
  if (rc != GU_ERR_UNSUPPORTED) return rc;
  return sweep_generic<float>(g, v_in, v_out, policy_kind, policy, gamma, residual, gate, gate_threshold,
                              static_cast<cudaStream_t>(stream));
}

extern "C" __attribute__((visibility("default"))) int gu_greedy_f64(const gu_grid* g, const double* v, uint8_t* tie_mask, double gamma, void* stream) {
  int rc = check_grid(g);
  if (rc) return rc;
  if (!v || !tie_mask) return GU_ERR_NULL;
  rc = greedy_tiled_f64(g, v, tie_mask, gamma, static_cast<cudaStream_t>(stream));
  if (rc != GU_ERR_UNSUPPORTED) return rc;
  return greedy_generic<double>(g, v, tie_mask, gamma, static_cast<cudaStream_t>(stream));
}

extern "C" __attribute__((visibility("default"))) int gu_greedy_f32(const gu_grid* g, const float* v, uint8_t* tie_mask, float gamma, void* stream) {
  int rc = check_grid(g);
  if (rc) return rc;
  if (!v || !tie_mask) return GU_ERR_NULL;
  rc = greedy_tiled_f32(g, v, tie_mask, gamma, static_cast<cudaStream_t>(stream));
  if (rc != GU_ERR_UNSUPPORTED) return rc;
  return greedy_generic<float>(g, v, tie_mask, gamma, static_cast<cudaStream_t>(stream));
}

namespace gu {
// Block size for a maze of N cells: enough threads for one cell each up to 1024, so that small mazes
// leave room for many resident blocks (= mazes in flight) per SM.
template <bool PI, int THREADS>
static int launch_small_t(const GridView& v, BatchStrides bs, int n_mazes, size_t smem, const double* v0,
                          double* v_out, uint8_t* tie, int kind, const void* policy, double gamma, double threshold,
                          int max_steps, int32_t* meta, double* delta, cudaStream_t st) {
  if (PI) {
    cudaError_t e = cudaFuncSetAttribute(pi_small_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    pi_small_kernel<THREADS><<<n_mazes, THREADS, smem, st>>>(v, bs, v0, v_out, tie, kind, policy, gamma, threshold,
                                                             max_steps, meta, delta);
  } else {
    cudaError_t e = cudaFuncSetAttribute(vi_small_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    vi_small_kernel<THREADS><<<n_mazes, THREADS, smem, st>>>(v, bs, v0, v_out, tie, kind, policy, gamma, threshold,
                                                             max_steps, meta, delta);
  }
  GU_CHECK_LAUNCH();
  return GU_OK;
}

template <bool PI>
static int launch_small(const GridView& v, BatchStrides bs, int n_mazes, bool single, const double* v0, double* v_out,
                        uint8_t* tie, int kind, const void* policy, double gamma, double threshold, int max_steps,
                        int32_t* meta, double* delta, cudaStream_t st) {
  const int64_t N = static_cast<int64_t>(v.X) * v.Y;
  const size_t smem = static_cast<size_t>(N) * (PI ? 26 : 17) + 16;
#define GU_SMALL(T) launch_small_t<PI, T>(v, bs, n_mazes, smem, v0, v_out, tie, kind, policy, gamma, threshold, max_steps, meta, delta, st)
  if (single || N > 2048) return GU_SMALL(1024);     // one grid: the whole SM works on it
  if (N > 512) return GU_SMALL(512);
  if (N > 128) return GU_SMALL(256);
  return GU_SMALL(128);
#undef GU_SMALL
}

static int check_batch(const gu_grid_batch* b, int64_t max_cells) {
  if (!b || !b->wall || !b->goal || !b->lava) return GU_ERR_NULL;
  const int64_t N = static_cast<int64_t>(b->X) * b->Y;
  if (b->X <= 0 || b->Y <= 0 || N > max_cells || b->n_mazes < 0 || b->pitch < b->X || b->pitch_words * 32 < b->X ||
      b->cell_stride < static_cast<int64_t>(b->Y + 2) * b->pitch ||
      b->plane_stride < static_cast<int64_t>(b->Y + 2) * b->pitch_words)
    return GU_ERR_SHAPE;
  return GU_OK;
}
static GridView batch_view(const gu_grid_batch* b) {
  GridView v;
  v.X = b->X; v.Y = b->Y; v.row_begin = 0; v.row_end = b->Y; v.pitch = b->pitch; v.pitch_words = b->pitch_words;
  v.wall = b->wall; v.goal = b->goal; v.lava = b->lava;
  return v;
}
}  // namespace gu

extern "C" __attribute__((visibility("default"))) int64_t gu_vi_small_max_cells(void) { return kSmallMaxCells; }

extern "C" __attribute__((visibility("default"))) int gu_vi_small_f64(const gu_grid* g, const double* v0, double* v_out, uint8_t* tie_mask,
                               int policy_kind, const void* policy, double gamma, double threshold,
                               int32_t max_steps, int32_t* sweeps_out, double* last_delta, void* stream) {
  int rc = check_grid(g);
  if (rc) return rc;
  if (!v0 || !v_out || !tie_mask || !sweeps_out || !last_delta) return GU_ERR_NULL;
  if ((policy_kind == GU_POLICY_PROBS || policy_kind == GU_POLICY_MASK) && !policy) return GU_ERR_NULL;
  if (policy_kind < 0 || policy_kind > GU_POLICY_GREEDY) return GU_ERR_MODE;
  const int64_t N = static_cast<int64_t>(g->X) * g->Y;
  if (g->row_begin != 0 || g->row_end != g->Y || N > kSmallMaxCells || max_steps < 0) return GU_ERR_SHAPE;
  return launch_small<false>(view_of(g), BatchStrides{0, 0}, 1, true, v0, v_out, tie_mask, policy_kind, policy, gamma,
                             threshold, max_steps, sweeps_out, last_delta, static_cast<cudaStream_t>(stream));
}

extern "C" __attribute__((visibility("default"))) int gu_vi_batch_f64(
    const gu_grid_batch* b, const double* v0, double* v_out, uint8_t* tie_mask, int policy_kind, const void* policy,
    double gamma, double threshold, int32_t max_steps, int32_t* sweeps_out, double* last_delta, void* stream) {
  int rc = check_batch(b, kSmallMaxCells);
  if (rc) return rc;
  if (!v_out || !tie_mask || !sweeps_out || !last_delta) return GU_ERR_NULL;
  if (policy_kind < GU_POLICY_PROBS || policy_kind > GU_POLICY_GREEDY) return GU_ERR_MODE;
  if ((policy_kind == GU_POLICY_PROBS || policy_kind == GU_POLICY_MASK) && !policy) return GU_ERR_NULL;
  if (max_steps < 0) return GU_ERR_SHAPE;
  if (b->n_mazes == 0) return GU_OK;
  return launch_small<false>(batch_view(b), BatchStrides{b->plane_stride, b->cell_stride}, b->n_mazes, false, v0, v_out,
                             tie_mask, policy_kind, policy, gamma, threshold, max_steps, sweeps_out, last_delta,
                             static_cast<cudaStream_t>(stream));
}

extern "C" __attribute__((visibility("default"))) int64_t gu_pi_small_max_cells(void) { return kPiSmallMaxCells; }

extern "C" __attribute__((visibility("default"))) int gu_pi_small_f64(
    const gu_grid* g, const double* v0, double* v_out, uint8_t* tie_mask, int policy_kind, const void* policy,
    double gamma, double threshold, int32_t max_steps, int32_t* meta, double* last_delta_eval, void* stream) {
  if (!g || !v0 || !v_out || !tie_mask || !meta || !last_delta_eval) return GU_ERR_NULL;
  if (policy_kind < GU_POLICY_PROBS || policy_kind > GU_POLICY_GREEDY) return GU_ERR_MODE;
  if ((policy_kind == GU_POLICY_PROBS || policy_kind == GU_POLICY_MASK) && !policy) return GU_ERR_NULL;
  const int64_t N = static_cast<int64_t>(g->X) * g->Y;
  if (g->row_begin != 0 || g->row_end != g->Y || N > kPiSmallMaxCells || max_steps < 0) return GU_ERR_SHAPE;
  return launch_small<true>(view_of(g), BatchStrides{0, 0}, 1, true, v0, v_out, tie_mask, policy_kind, policy, gamma,
                            threshold, max_steps, meta, last_delta_eval, static_cast<cudaStream_t>(stream));
}

extern "C" __attribute__((visibility("default"))) int gu_pi_batch_f64(
    const gu_grid_batch* b, const double* v0, double* v_out, uint8_t* tie_mask, int policy_kind, const void* policy,
    double gamma, double threshold, int32_t max_steps, int32_t* meta, double* last_delta_eval, void* stream) {
  int rc = check_batch(b, kPiSmallMaxCells);
  if (rc) return rc;
  if (!v_out || !tie_mask || !meta || !last_delta_eval) return GU_ERR_NULL;
  if (policy_kind < GU_POLICY_PROBS || policy_kind > GU_POLICY_GREEDY) return GU_ERR_MODE;
  if ((policy_kind == GU_POLICY_PROBS || policy_kind == GU_POLICY_MASK) && !policy) return GU_ERR_NULL;
  if (max_steps < 0) return GU_ERR_SHAPE;
  if (b->n_mazes == 0) return GU_OK;
  return launch_small<true>(batch_view(b), BatchStrides{b->plane_stride, b->cell_stride}, b->n_mazes, false, v0, v_out,
                            tie_mask, policy_kind, policy, gamma, threshold, max_steps, meta, last_delta_eval,
                            static_cast<cudaStream_t>(stream));
}
