/* gu_b200.h -- C ABI of the B200-native GridUniverse hot path (libgu_b200.so).
 *
 * The reference (TheMTank/GridUniverse) is pure Python and has no FFI; the
 * boundary a replacement sits behind is its Python API.  Each entry point below
 * replaces the reference function cited next to it (paths relative to the
 * reference root) for a whole batch / whole grid at once.  INTEGRATION.md shows
 * the ctypes stub a maintainer would add on the reference side.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes.  Every pointer inside the descriptor
 *    structs and every array argument is a DEVICE-ACCESSIBLE pointer owned by the
 *    caller (device memory; pinned host memory also works under unified addressing
 *    and is what the single-env step path passes); the library never allocates,
 *    frees or keeps a pointer after the call.
 *  - Every call only enqueues work on `stream` (a cudaStream_t passed as void*).
 *  - Return value: 0 = ok, <0 = argument error (GU_ERR_*), >0 = cudaError_t.
 *  - Grid geometry (core/envs/griduniverse_env.py:44-56): X = x_max columns,
 *    Y = y_max rows, state s = y*X + x, actions 0=UP 1=RIGHT 2=DOWN 3=LEFT.
 *  - Rewards are derived from the masks: lava -10, else goal +10, else -1
 *    (griduniverse_env.py:80-90; lava is written last and wins).
 */
#ifndef GU_B200_H
#define GU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GU_OK 0
#define GU_ERR_NULL (-1)      /* a required pointer is NULL */
#define GU_ERR_SHAPE (-2)     /* X/Y/N/T/pitch out of the supported range */
#define GU_ERR_ALIGN (-3)     /* pointer or pitch violates the documented alignment */
#define GU_ERR_MODE (-4)      /* unknown policy kind / flag / table format */
#define GU_ERR_UNSUPPORTED (-5)

/* ---- environment batches ------------------------------------------------ */

/* Levels for a batch of N envs that all have the same X x Y shape.
 * Bit planes are dense: bit (s & 31) of word (s >> 5) is cell s; `words` =
 * ceil(X*Y/32).  per_env = 0: one shared level, planes are uint32[words],
 * `start` is int32[1].  per_env = 1: one level per env, planes are WORD-MAJOR
 * uint32[words][N] (word w of env n at w*N + n, so a warp reading one word for
 * 32 consecutive envs is one coalesced request), `start` is int32[N].
 * Replaces wall_grid / goal_states / lava_states / reward_matrix /
 * starting_states (griduniverse_env.py:61-90,120-134). */
typedef struct {
  int32_t X, Y;
  int32_t per_env;
  int32_t words;
  const uint32_t* wall;
  const uint32_t* goal;
  const uint32_t* lava;
  const int32_t* start;
} gu_levels;

#define GU_FLAG_AUTO_RESET 1u      /* after a done step the env continues from its start state */
#define GU_FLAG_NO_CARE_TERMINAL 2u /* care_about_terminal=False (griduniverse_env.py:150-153) */
#define GU_FLAG_ACCUMULATE 4u      /* gu_rollout: env_return / env_done += instead of = (streamed slabs) */
#define GU_FLAG_PACKED_ACTIONS 8u  /* gu_rollout: `actions` holds 2 bits per step, 16 steps per 32-bit word:
                                    * uint32[ceil(T/16)][N], step t of env n in bits 2*(t%16).. of word
                                    * [t/16][n] -- a sixteenth of the bytes of the int32 stream */

/* One step for N envs: GridUniverseEnv._step (griduniverse_env.py:176-185) =
 * look_step_ahead(current_state, action) (:136-155) for every env.
 *   actions int32[N]   (only the two low bits are used: -1 is LEFT like the
 *                       reference's negative list index; the host validates >3)
 *   pos     int32[N]   in: current states, out: states the envs continue from
 *   obs     int32[N]   out: the landing cell `step` returns (may be NULL)
 *   reward  int32[N]   out, done uint8[N] out (either may be NULL)
 *   start_choice int32[N] or NULL: start state to use if env n resets now
 *                       (host-supplied stand-in for random.choice, :189)
 *   stats   int64[2] or NULL: += sum of rewards, += number of done flags
 */
int gu_step(const gu_levels* lv, int64_t n_envs, const int32_t* actions, int32_t* pos,
            int32_t* obs, int32_t* reward, uint8_t* done, const int32_t* start_choice,
            int64_t* stats, uint32_t flags, void* stream);

/* T steps for N envs in one launch (run_episode's inner loop,
 * core/algorithms/monte_carlo.py:19-25, with host-supplied action streams).
 *   actions int32[T][N];  pos int32[N] in/out
 *   obs int32[T][N], reward int32[T][N], done uint8[T][N]: trajectories, each may be NULL
 *   start_choice int32[T][N] or NULL
 *   env_return int32[N] / env_done int32[N] or NULL: per-env reward sum / done count (overwritten)
 *   stats int64[2] or NULL as in gu_step (warp-shuffle reduced, one atomic per warp)
 *   tables: NULL, or transition tables built by gu_pack_tables for the same levels
 */
int gu_rollout(const gu_levels* lv, int64_t n_envs, int64_t n_steps, const int32_t* actions,
               int32_t* pos, int32_t* obs, int32_t* reward, uint8_t* done,
               const int32_t* start_choice, int32_t* env_return, int32_t* env_done,
               int64_t* stats, const uint32_t* tables, uint32_t flags, void* stream);

/* int32 actions [T][N] (two low bits used, like gu_rollout) -> the packed stream of
 * GU_FLAG_PACKED_ACTIONS, uint32[ceil(T/16)][N].  gu_pack_actions: device pointers, enqueued on
 * `stream`.  gu_pack_actions_host: HOST pointers, runs on n_threads CPU threads (0 = all cores) and
 * returns when done -- for callers whose action source is a host int32 array.  Full 16-step rows are
 * packed with AVX2 (run-time check) or SSE2 vector code, 16 envs per iteration. */
int gu_pack_actions(const int32_t* actions, int64_t n_steps, int64_t n_envs, uint32_t* packed, void* stream);
int gu_pack_actions_host(const int32_t* actions, int64_t n_steps, int64_t n_envs, uint32_t* packed,
                         int32_t n_threads);

/* Policy-driven episodes: run_episode (core/algorithms/monte_carlo.py:7-26) for N
 * episodes on one SHARED level, the randomness host-supplied as uniform draws.
 *   cdf      f64[cells][4]  cumulative action probabilities per state, built exactly like
 *                           np.random.choice does (p.cumsum(); cdf /= cdf[-1])
 *   uniforms f64[T][N]      draw t of episode n; action = #{a : cdf[s][a] <= u}
 *                           (searchsorted side='right', monte_carlo.py:20)
 *   pos      int32[N]       in: start states, out: final states
 *   obs int32[T][N], reward int32[T][N] trajectories (entries past an episode's end are
 *   left untouched), length int32[N] = steps taken (<= T), done uint8[N]; any may be NULL.
 * An episode stops at its first done step (monte_carlo.py:24-25).  A cdf row of NaN marks a state
 * whose probabilities np.random.choice would reject: an episode that has to sample there stops with
 * length = -1 - (steps taken) and done = 0. */
int gu_rollout_policy(const gu_levels* lv, int64_t n_envs, int64_t n_steps, const double* cdf,
                      const double* uniforms, int32_t* pos, int32_t* obs, int32_t* reward,
                      int32_t* length, uint8_t* done, void* stream);

/* One episode's contribution to Monte-Carlo evaluation (monte_carlo.py:53-91).
 *   start int32[1] device, obs / rewards: the episode's trajectory as written by
 *   gu_rollout_policy (element t at t*stride), episode_len = steps taken (L)
 *   weights f64[>=L]: discount**i, keep uint8[>=L]: discount**i > threshold (built on the
 *   host with the interpreter's float pow, like the reference's expression at :69-70)
 *   every_visit: 0 = first-visit, 1 = every-visit
 *   mode: 0 = incremental mean, stationary (V += (G-V)/N for every state with N > 0)
 *         1 = incremental, constant alpha  (V += alpha*(G-V) for every state)
 *         2 = batch: only total_visits / total_return are accumulated (finalize later)
 *   g_scratch f64[L+1]; total_visits / total_return / value f64[cells], updated in place.
 * Sums run left to right in fp64 (the reference's pinned CPython 3.6 `sum`; CPython >= 3.12
 * compensates float sums and differs in the last bits). */
int gu_mc_episode_f64(int32_t cells, int32_t episode_len, const int32_t* start, const int32_t* obs,
                      const int32_t* rewards, int64_t stride, const double* weights,
                      const uint8_t* keep, int32_t every_visit, int32_t mode, double alpha,
                      double* g_scratch, double* total_visits, double* total_return, double* value,
                      void* stream);
/* monte_carlo_evaluation's episode loop (monte_carlo.py:49-91) for n_episodes episodes on one SHARED
 * level in ONE launch: episode e starts in starts[e] (the host's random.choice, griduniverse_env.py:189),
 * takes its actions from the uniform draws that follow episode e-1's (uniforms f64[n_uniforms], one
 * stream for the whole batch, consumed exactly like np.random.choice would, :20), and is folded into
 * total_visits / total_return / value like gu_mc_episode_f64 does, episode after episode.
 *   obs_scratch / rew_scratch int32[max_steps], g_scratch f64[max_steps + 1]
 *   lengths int32[n_episodes], done uint8[n_episodes]: per-episode results
 *   meta int64[4]: episodes completed, draws consumed, status (0 ok; 1 ran out of draws; 2 + s: state
 *   s was visited and its cdf row is NaN = np.random.choice would raise ValueError there), state the
 *   last completed episode ended in.  Everything device memory. */
int gu_mc_evaluate_f64(const gu_levels* lv, const double* cdf, const double* uniforms, int64_t n_uniforms,
                       const int32_t* starts, int32_t n_episodes, int32_t max_steps, const double* weights,
                       const uint8_t* keep, int32_t every_visit, int32_t mode, double alpha, int32_t* obs_scratch,
                       int32_t* rew_scratch, double* g_scratch, double* total_visits, double* total_return,
                       double* value, int32_t* lengths, uint8_t* done, int64_t* meta, void* stream);
/* V(s) = S(s) / N(s) where N(s) > 0 (monte_carlo.py:93-97). */
int gu_mc_finalize_f64(int32_t cells, const double* total_visits, const double* total_return,
                       double* value, void* stream);

/* Size in bytes of the transition tables for (lv, n_envs), 0 if the shape has no table format. */
int64_t gu_tables_bytes(const gu_levels* lv, int64_t n_envs);
/* Build transition tables (device) from the bit planes.  Per-env levels (X*Y <= 256, X <= 127):
 * one info byte per cell ("action a moves" bits + goal / lava), four cells per word, word-major
 * like the planes.  Shared level (X*Y <= 16383): landing cell + goal / lava flags per
 * (cell, action) as uint16. */
int gu_pack_tables(const gu_levels* lv, int64_t n_envs, uint32_t* tables, uint32_t flags, void* stream);

/* Resident look_step_ahead service for a one-env step loop (griduniverse_env.py:176-185): launches a
 * one-warp kernel that answers (state, action) requests written into `mailbox` -- 32 x uint32, 128-byte
 * aligned, pinned host memory -- until `idle_cycles` SM cycles pass without a request or `max_cycles`
 * since the launch, then clears the alive word and leaves (relaunch on demand).  Shared level only.
 *   mailbox[0:2]   request, ONE 8-byte host store: bits 0-31 sequence number (any value different from
 *                  the previous request's; `seq0` = the last one already answered), 32-33 action,
 *                  34 = care_about_terminal is False, 35-63 state
 *   mailbox[16:20] answer, one 16-byte device store: {sequence number, next state, reward, terminal}
 *   mailbox[20]    alive: the host sets it to 1 before calling, the kernel clears it when it leaves */
int gu_look_server_start(const gu_levels* lv, void* mailbox, uint32_t seq0, int64_t idle_cycles,
                         int64_t max_cycles, void* stream);

/* look_step_ahead for M arbitrary (state, action) pairs (griduniverse_env.py:136-155).
 * Shared level: any M.  per_env levels: pair i is evaluated on level i (M == N).
 *   states int32[M], actions int32[M] -> next int32[M], reward int32[M], terminal uint8[M] */
int gu_look_step_ahead(const gu_levels* lv, int64_t m, const int32_t* states, const int32_t* actions,
                       int32_t* next, int32_t* reward, uint8_t* terminal, uint32_t flags, void* stream);

/* Batched level text -> per-env bit planes (griduniverse_env.py:253-300): `text` holds n_levels
 * levels of X*Y characters each, whitespace already removed (:248-249), row-major.  Writes the
 * word-major planes of a per-env `gu_levels` -- uint32[words][n_levels] --, the first start state of
 * each level (`start`), optionally how many 'x' it has (`n_starts`, may be NULL), and per level
 * `status`: 0 = ok; k > 0 = character k-1 is not one of "o#GLx" (ValueError, :292-293);
 * GU_TEXT_NO_START / GU_TEXT_NO_GOAL = the two ValueErrors raised after the scan (:297-300). */
#define GU_TEXT_NO_START (-1)
#define GU_TEXT_NO_GOAL (-2)
int gu_pack_level_text(const uint8_t* text, int64_t n_levels, int32_t X, int32_t Y, uint32_t* wall,
                       uint32_t* goal, uint32_t* lava, int32_t* start, int32_t* n_starts,
                       int32_t* status, void* stream);

/* Batched render(mode='ansi') (griduniverse_env.py:202-221): for each env the Y text rows of
 * "<glyph><blank>" pairs plus '\n', then one closing '\n'; glyph precedence x (agent at pos[env])
 * < G < L < #.  `text`: uint8[n][Y*(2X+1)+1]. */
int gu_render_ansi(const gu_levels* lv, int64_t n, const int32_t* pos, uint8_t* text, void* stream);

/* Batched headless RGB frames: what the reference's viewer draws (core/envs/rendering.py:121-135: one
 * tile per cell -- ground / wall / goal / lava, the agent on top) and, with `policy`, what
 * render_policy_arrows adds (:159-212: per non-terminal non-wall cell and action with p >= 0.1 a line of
 * round(p * 20) pixels from the tile centre plus an arrowhead, for 32-pixel tiles), rasterised without
 * a GL window (the reference's 'rgb_array' mode is half-wired, griduniverse_env.py:223-230).  Flat
 * colours stand in for the sprites.
 *   pos int32[n] or NULL (no agent); policy f64[cells][4] shared, f64[n][cells][4] with
 *   policy_per_env = 1, or NULL (no arrows); tile = pixels per cell, a multiple of 16;
 *   rgb uint8[n][Y * tile][X * tile][3], image row 0 = grid row 0. */
int gu_render_rgb(const gu_levels* lv, int64_t n, const int32_t* pos, const double* policy, int32_t policy_per_env,
                  int32_t tile, uint8_t* rgb, void* stream);

/* ---- whole-grid planning ------------------------------------------------ */

/* One grid, or one row shard of it, for the sweep / greedy kernels.
 * The shard owns rows [row_begin, row_end) of a Y-row grid.  EVERY per-cell
 * array (value functions, policies, tie masks) holds (row_end-row_begin+2) rows
 * of `pitch` elements: array row 0 is the ghost row above the shard, the last
 * row the ghost row below (filled by the halo exchange; never read at the true
 * grid edges).  The bit planes hold the same rows with `pitch_words` uint32 per
 * row, bit (x & 31) of word (x >> 5); bits at x >= X are zero.
 * pitch >= X.
 * Limits (GU_ERR_SHAPE otherwise): a shard owns at least one row (row_begin < row_end: a grid of Y rows
 * shards over at most Y ranks) and its arrays hold at most 65,535 rows (rows + 2 ghost rows); the
 * layout-agnostic kernels take one block row per grid row, so they too stop at 65,535 rows per call.
 * Taller grids are split into row shards by the caller -- on one GPU as well: shards are independent
 * launches joined by their ghost rows (griduniverse_b200/sharded.py).
 * `info` (optional, may be NULL): derived per-cell byte plane built once per level by
 * gu_pack_info, same padded layout as the per-cell arrays: bit0/1/2/5 = UP/RIGHT/DOWN/LEFT is
 * blocked (grid edge | wall at the target | cell terminal), bit 3 goal, bit 4 lava.  With it, and
 * with pitch % 4 == 0, pitch*sizeof(T) % 16 == 0 and 16-byte aligned arrays, the sweep /
 * greedy entry points run the register-tiled kernels; otherwise the layout-agnostic ones. */
typedef struct {
  int32_t X, Y;
  int32_t row_begin, row_end;
  int32_t pitch;
  int32_t pitch_words;
  const uint32_t* wall;
  const uint32_t* goal;
  const uint32_t* lava;
  const uint8_t* info;
} gu_grid;

/* Build the `info` plane (uint8[(rows+2)*pitch]) of a grid / shard from its bit planes. */
int gu_pack_info(const gu_grid* g, uint8_t* info, void* stream);

#define GU_POLICY_PROBS 0   /* policy = T[cells][4] probabilities (any stochastic policy) */
#define GU_POLICY_MASK 1    /* policy = uint8[cells] tie masks: prob 1/popcount on set bits */
#define GU_POLICY_UNIFORM 2 /* policy = NULL: 1/4 everywhere (policy0 of the examples) */
#define GU_POLICY_GREEDY 3  /* policy = NULL: tie set recomputed from v_in in the same pass (VI) */

/* One synchronous sweep, single_step_policy_evaluation (core/algorithms/utils.py:15-27):
 *   v_out[s] = (((R[s] + p0*(g*v[n0])) + p1*(g*v[n1])) + p2*(g*v[n2])) + p3*(g*v[n3])
 * in the reference's left-to-right order without fused multiply-add.  With
 * GU_POLICY_GREEDY the policy is greedy_policy_from_value_function(v_in)
 * (utils.py:55-72), i.e. one value-iteration pass (dynamic_programming.py:16-20).
 *   residual: T* device scalar, combined with max(v_in - v_out) (signed,
 *   dynamic_programming.py:17) by atomic max; caller initialises it to -inf. May be NULL.
 *   gate / gate_threshold: if gate != NULL and *gate < gate_threshold when the kernel
 *   starts, the sweep is a no-op (v_out and residual untouched).  Passing the previous
 *   sweep's residual and the convergence threshold lets a caller enqueue many sweeps
 *   without a host round trip: the ones after convergence (dynamic_programming.py:22-23)
 *   do nothing and the converged V stays where it was written. */
int gu_sweep_f64(const gu_grid* g, const double* v_in, double* v_out, int policy_kind,
                 const void* policy, double gamma, double* residual, const double* gate,
                 double gate_threshold, void* stream);
int gu_sweep_f32(const gu_grid* g, const float* v_in, float* v_out, int policy_kind,
                 const void* policy, float gamma, float* residual, const float* gate,
                 float gate_threshold, void* stream);

/* Row-sharded sweeps fused with their collectives over NVLink peer memory (no NCCL in the
 * loop).  Each rank passes pointers into its neighbours' and peers' memory (e.g. from
 * torch.distributed._symmetric_memory).  Sweep number `slot` (0, 1, 2, ... within one solve):
 *   halo exchange   the blocks that own the shard's first / last row also store their output rows
 *     into `up_ghost` / `down_ghost`, the ghost row of the neighbour's v_out that mirrors them (NULL at
 *     the grid edge).  When the last of those blocks is done it stores slot+1 into the neighbour's
 *     flag word (`up_flag` / `down_flag`, pointing at the neighbour's halo_flags[1] / halo_flags[0]).
 *     Only the first / last block row of sweep slot >= 1 waits, and only for its own neighbour:
 *     halo_flags[0] >= slot (rows from above have arrived, and the neighbour above has finished
 *     reading the ghost row this sweep is about to overwrite), halo_flags[1] >= slot likewise.
 *   residual / stopping rule   when the last block of the shard finishes, max(v_in - v_out) of the
 *     shard goes to entry [slot][rank] of EVERY rank's residual table res_tables[r]
 *     (T[n_slots][world], NaN = not written yet; res_tables[rank] is this rank's own).  Sweep `slot`
 *     checks the slot `slot - gate_lag` (only slots >= first_slot): it waits until all `world` entries
 *     have arrived locally and, if their max is below `threshold`, is a no-op and sets the sticky word
 *     `stop_flag`, which turns every later sweep into a no-op too.  This is the reference's stopping
 *     rule (dynamic_programming.py:17,22-23) evaluated identically on every rank.  gate_lag = 1 stops
 *     exactly after the converged sweep; gate_lag = 2 (the default driver) lets one more sweep run
 *     into the OTHER ping-pong buffer, so the converged V is intact and no rank ever stalls on the
 *     slowest rank's residual.
 *   errors   every wait gives up after `timeout_cycles` (0 = about 3 s): the rank sets `error_flag`,
 *     stores 1 into every rank's abort word (`abort_flags[r]`, this rank's own included) and stops
 *     writing; a rank that finds its abort word set skips all remaining work, so the host of EVERY
 *     rank sees the failure at its next check instead of consuming stale ghost rows.
 *   slot_base (optional, device int32): the effective slot is slot + *slot_base, so a captured CUDA
 *     graph of a chunk of sweeps can be replayed with only that word changing.
 *   residual: local T scalar of this sweep, initialised to -inf; done_counter int32 and
 *     edge_counters int32[2]: local, zero between launches.
 * Needs the tiled layout (info plane, pitch % 32 == 0).  gu_peer_wait blocks the stream until every
 * entry of slot `slot` has arrived and both neighbours have delivered the rows of that sweep (use it
 * before reading the table or the ghost rows on the host / in another kernel). */
#define GU_MAX_PEERS 16
typedef struct {
  int32_t rank, world;
  int32_t slot, n_slots;
  void* up_ghost;
  void* down_ghost;
  void* res_tables[GU_MAX_PEERS];
  int32_t* done_counter;
  int32_t* error_flag;
  double threshold;
  int32_t gate_lag;             /* 1 or 2 */
  int32_t first_slot;           /* slots below this one belong to an earlier phase and are never gated on */
  int32_t* halo_flags;          /* local int32[2]: sweeps delivered by the neighbour above / below */
  int32_t* up_flag;             /* neighbour above: its halo_flags + 1, or NULL */
  int32_t* down_flag;           /* neighbour below: its halo_flags + 0, or NULL */
  int32_t* edge_counters;       /* local int32[2] */
  int32_t* stop_flag;           /* local int32, sticky "converged" */
  int32_t* abort_flags[GU_MAX_PEERS];
  int64_t timeout_cycles;
  const int32_t* slot_base;
} gu_peer_links;

int gu_sweep_peer_f32(const gu_grid* g, const float* v_in, float* v_out, int policy_kind,
                      const void* policy, float gamma, float* residual, const gu_peer_links* peer,
                      void* stream);
int gu_sweep_peer_f64(const gu_grid* g, const double* v_in, double* v_out, int policy_kind,
                      const void* policy, double gamma, double* residual, const gu_peer_links* peer,
                      void* stream);
int gu_peer_wait(const gu_peer_links* peer, int is_f64, void* stream);

/* Signed max of (a - b) over the shard's owned cells (x < X), combined into *out by atomic max
 * (caller initialises *out to -inf): np.max(last_converged_v_fun - new_value_function),
 * dynamic_programming.py:44.  a, b: padded per-cell arrays of the grid layout. */
int gu_max_diff_f32(const gu_grid* g, const float* a, const float* b, float* out, void* stream);
int gu_max_diff_f64(const gu_grid* g, const double* a, const double* b, double* out, void* stream);

/* greedy_policy_from_value_function (utils.py:55-72) as a tie mask per cell:
 * bit a set <=> rint(q[s,a]*1e8) == rint(max_a q[s,a]*1e8) and s is not terminal,
 * q[s,a] = R[next] + g*v[next].  np.argmax of the expanded row is ctz(mask). */
int gu_greedy_f64(const gu_grid* g, const double* v, uint8_t* tie_mask, double gamma, void* stream);
int gu_greedy_f32(const gu_grid* g, const float* v, uint8_t* tie_mask, float gamma, void* stream);

/* ---- shortest paths (cross-check of greedy policies) --------------------- */

/* Breadth-first distances over the graph the reference builds with
 * look_step_ahead(s, a, care_about_terminal=False) on the non-wall cells
 * (core/algorithms/maze_solving.py:43-50) -- searched there by a FIFO queue from one state to the
 * first terminal it reaches (:123-168).  Here: a multi-source wavefront, one level per kernel
 * launch, 32 cells per word.  Whole grids only (row_begin == 0, row_end == Y), pitch % 32 == 0,
 * pitch <= 32*pitch_words, pitch_words % 4 == 0, planes 16-byte aligned.  `visited_a` / `visited_b`: uint32[(Y+2)*pitch_words] ping-pong planes;
 * `dist`: int32[(Y+2)*pitch], 16-byte aligned, -1 = not reached; `reached` (uint64, caller zeroes
 * it) is incremented by the number of cells each level reaches.
 *   gu_bfs_init    sources (NULL = the goal plane) restricted to enterable cells get distance 0.
 *   gu_bfs_expand  runs levels level_begin .. level_begin+n_levels-1 (level_begin >= 1, continuing
 *                  from the previous call); extra levels after the wavefront died are no-ops, so
 *                  the caller polls `reached` every chunk of levels instead of every level.
 *   gu_bfs_walk    the action list of a shortest path from start_state to the nearest source
 *                  (replaces construct_path, maze_solving.py:170-193): at every step the
 *                  lowest-numbered action that lands one level closer.  *length = number of
 *                  actions, -1 if start_state was not reached, -3 if max_len is too small.
 * GU_BFS_LAVA_BLOCKS: lava cells are not entered (distance to the goal along cells an optimal
 * policy may use); default: only walls block, as in the reference graph. */
#define GU_BFS_LAVA_BLOCKS 1u
int gu_bfs_init(const gu_grid* g, const uint32_t* sources, uint32_t* visited_a, uint32_t* visited_b,
                int32_t* dist, uint64_t* reached, uint32_t flags, void* stream);
int gu_bfs_expand(const gu_grid* g, uint32_t* visited_a, uint32_t* visited_b, int32_t* dist,
                  int32_t level_begin, int32_t n_levels, uint64_t* reached, uint32_t flags, void* stream);
int gu_bfs_walk(const gu_grid* g, const int32_t* dist, int64_t start_state, int8_t* actions,
                int32_t max_len, int32_t* length, void* stream);

/* Whole value_iteration loop (dynamic_programming.py:8-28) for a grid small enough to
 * live in one thread block's shared memory (see gu_vi_small_max_cells): sweeps until
 * max(V - V') < threshold or max_steps, then writes V, the tie masks of the final V,
 * and the number of sweeps done.  Arrays use the padded layout of gu_grid; the shard
 * must be the whole grid.
 *   policy_kind/policy describe the caller's initial policy (used by the first sweep). */
int gu_vi_small_f64(const gu_grid* g, const double* v0, double* v_out, uint8_t* tie_mask,
                    int policy_kind, const void* policy, double gamma, double threshold,
                    int32_t max_steps, int32_t* sweeps_out, double* last_delta, void* stream);
int64_t gu_vi_small_max_cells(void);

/* Whole policy_iteration loop (dynamic_programming.py:31-57) in one thread block, for whole
 * grids of at most gu_pi_small_max_cells() cells: evaluate the current policy until
 * max(V - V') < threshold, take the greedy policy of the result, stop when
 * max(V_last_converged - V') < threshold or after max_steps sweeps.  Writes the last converged V,
 * the tie masks of the final policy (only if an improvement ran: meta[1]) and
 * meta = {sweeps, improved, exhausted}; exhausted = the reference's non-convergence warning. */
int gu_pi_small_f64(const gu_grid* g, const double* v0, double* v_out, uint8_t* tie_mask,
                    int policy_kind, const void* policy, double gamma, double threshold,
                    int32_t max_steps, int32_t* meta, double* last_delta_eval, void* stream);
int64_t gu_pi_small_max_cells(void);

/* A batch of n_mazes same-shape grids solved in ONE launch, one thread block per maze (the many
 * small mazes of the reference's examples, examples/griduniverse_alg_examples.py:29-59; independent
 * units: shard them over GPUs by maze range, no collective).  Maze m's bit planes start
 * m * plane_stride uint32 words and its per-cell arrays (v0, v_out, tie_mask, policy) m * cell_stride
 * elements after maze 0's, each laid out like a whole grid in struct gu_grid: Y + 2 rows of pitch / pitch_words.
 * Per-maze outputs: sweeps_out / last_delta (value iteration), meta[m][3] = {sweeps, improved,
 * exhausted} / last_delta_eval (policy iteration) -- same meaning as gu_vi_small_f64 / gu_pi_small_f64,
 * to which every maze's result is bit-identical.  v0 = NULL: value functions of zeros. */
typedef struct {
  int32_t X, Y;
  int32_t n_mazes;
  int32_t pitch, pitch_words;
  int64_t cell_stride;
  int64_t plane_stride;
  const uint32_t* wall;
  const uint32_t* goal;
  const uint32_t* lava;
} gu_grid_batch;
int gu_vi_batch_f64(const gu_grid_batch* b, const double* v0, double* v_out, uint8_t* tie_mask, int policy_kind,
                    const void* policy, double gamma, double threshold, int32_t max_steps, int32_t* sweeps_out,
                    double* last_delta, void* stream);
int gu_pi_batch_f64(const gu_grid_batch* b, const double* v0, double* v_out, uint8_t* tie_mask, int policy_kind,
                    const void* policy, double gamma, double threshold, int32_t max_steps, int32_t* meta,
                    double* last_delta_eval, void* stream);

/* ---- synthetic levels (not in the reference; pure functions of seed and index) -------- */

/* Per-env levels for envs [first_env, first_env + n_envs): border open, 20 % interior walls,
 * one goal, X*Y/32 lava draws, one start on an open non-terminal cell.  X*Y <= 256.
 * Planes WORD-MAJOR uint32[words][n_envs], start int32[n_envs] (see gu_levels). */
int gu_synth_env_levels(int32_t X, int32_t Y, int64_t n_envs, int64_t first_env, uint32_t seed,
                        uint32_t* wall, uint32_t* goal, uint32_t* lava, int32_t* start, void* stream);

/* Maze planes in the gu_grid layout for rows [row_begin-1, row_end+1): (x,y) is a wall iff
 * x,y both odd, or exactly one is odd and hash(seed,y,x) < 1/4; goal at the (even,even) cell
 * next to the centre; open cells are lava with probability 0.001. */
int gu_synth_maze(int32_t X, int32_t Y, int32_t row_begin, int32_t row_end, int32_t pitch_words,
                  uint32_t seed, uint32_t* wall, uint32_t* goal, uint32_t* lava, void* stream);

/* ---- queries -------------------------------------------------------------- */
int gu_version(void);                 /* 10000*major + 100*minor + patch */
const char* gu_arch(void);            /* "sm_100a" */
const char* gu_error_string(int code);

#ifdef __cplusplus
}
#endif
#endif /* GU_B200_H */
