"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's hot path.

Every function cites the reference lines it restates (paths relative to
/root/reference).  Vectorised NumPy is used for whole-grid work, plain Python
loops only for the small scalar cases.  Parity status: pinned (see
``oracle/__init__.py``).

Conventions (core/envs/griduniverse_env.py:44-56): ``X = x_max`` columns,
``Y = y_max`` rows, state ``s = y*X + x``, actions 0=UP 1=RIGHT 2=DOWN 3=LEFT.
"""
import warnings

import numpy as np

UP, RIGHT, DOWN, LEFT = 0, 1, 2, 3
REWARD_STEP, REWARD_GOAL, REWARD_LAVA = -1, 10, -10  # griduniverse_env.py:80,83,88


# --------------------------------------------------------------------------
# Level model
# --------------------------------------------------------------------------
class Level(object):
    """Static description of one grid (griduniverse_env.py:44-90)."""

    def __init__(self, X, Y, walls=(), goals=None, lavas=(), starts=(0,)):
        self.X, self.Y = int(X), int(Y)
        self.N = self.X * self.Y
        # default goal = last cell when none given (griduniverse_env.py:66-67)
        if goals is None or len(goals) == 0:
            goals = [self.N - 1]
        self.goals = [int(g) for g in goals]
        self.lavas = [int(l) for l in lavas]
        self.walls = [int(w) for w in walls]
        self.starts = [int(s) for s in starts]
        self.wall = np.zeros(self.N, dtype=bool)
        self.goal = np.zeros(self.N, dtype=bool)
        self.lava = np.zeros(self.N, dtype=bool)
        for w in self.walls:
            if w < 0 or w > self.N - 1:  # griduniverse_env.py:130-131
                raise ValueError("Wall state {} is out of grid bounds".format(w))
            self.wall[w] = True
        self.goal[self.goals] = True
        if self.lavas:
            self.lava[self.lavas] = True
        self.term = self.goal | self.lava  # griduniverse_env.py:163-174
        # reward: -1, goals +10, then lava -10 (lava written last wins) :80-90
        self.reward = np.full(self.N, REWARD_STEP, dtype=np.int64)
        self.reward[self.goal] = REWARD_GOAL
        self.reward[self.lava] = REWARD_LAVA

    @classmethod
    def from_masks(cls, X, Y, wall, goal, lava, starts=(0,)):
        wall = np.asarray(wall).reshape(-1).astype(bool)
        goal = np.asarray(goal).reshape(-1).astype(bool)
        lava = np.asarray(lava).reshape(-1).astype(bool)
        lv = cls.__new__(cls)
        lv.X, lv.Y, lv.N = int(X), int(Y), int(X) * int(Y)
        lv.wall, lv.goal, lv.lava = wall, goal, lava
        lv.term = goal | lava
        lv.reward = np.full(lv.N, REWARD_STEP, dtype=np.int64)
        lv.reward[goal] = REWARD_GOAL
        lv.reward[lava] = REWARD_LAVA
        lv.starts = [int(s) for s in starts]
        lv.goals = lv.lavas = lv.walls = None  # not materialised for big grids
        return lv


def strip_level_lines(raw_lines):
    """griduniverse_env.py:246-251: rstrip, drop blank lines, remove all whitespace."""
    lines = [line.rstrip() for line in raw_lines]
    return ["".join(line.split()) for line in lines if line]


def parse_level_text(text_world_lines):
    """griduniverse_env.py:253-300 -- row-major scan of 'G L o # x'."""
    goals, starts, lavas, walls = [], [], [], []
    idx = 0
    width = len(text_world_lines[0])
    for line in text_world_lines:
        if len(line) != width:
            raise ValueError("Input text file is not a rectangle")
        for ch in line:
            if ch == 'G':
                goals.append(idx)
            elif ch == 'L':
                lavas.append(idx)
            elif ch == 'o':
                pass
            elif ch == '#':
                walls.append(idx)
            elif ch == 'x':
                starts.append(idx)
            else:
                raise ValueError('Invalid Character "{}". Returning'.format(ch))
            idx += 1
    if len(starts) == 0:
        raise ValueError("No starting states set in text file. Place \"x\" within grid. ")
    if len(goals) == 0:
        raise ValueError("No terminal goal states set in text file. Place \"T\" within grid. ")
    return Level(width, len(text_world_lines), walls=walls, goals=goals, lavas=lavas, starts=starts)


# --------------------------------------------------------------------------
# Transition function
# --------------------------------------------------------------------------
def clamp_move(level, s, a):
    """The four lambdas at griduniverse_env.py:51-54 (scalar)."""
    X, Y = level.X, level.Y
    x, y = s % X, s // X
    if a == UP:
        return s - X if y > 0 else s
    if a == RIGHT:
        return s + 1 if x < X - 1 else s
    if a == DOWN:
        return s + X if y < Y - 1 else s
    if a == LEFT:
        return s - 1 if x > 0 else s
    raise IndexError("list index out of range")


def look_step_ahead(level, s, a, care_about_terminal=True):
    """griduniverse_env.py:136-155 (scalar) -> (next, reward, terminal)."""
    if care_about_terminal and level.term[s]:
        n = s
    else:
        c = clamp_move(level, s, a)
        n = s if level.wall[c] else c
    return int(n), int(level.reward[n]), bool(level.term[n])


def next_table(level, care_about_terminal=True):
    """Vectorised griduniverse_env.py:136-153 for every (s, a): int64 [N,4]."""
    X, Y, N = level.X, level.Y, level.N
    s = np.arange(N, dtype=np.int64)
    x, y = s % X, s // X
    cand = np.stack([np.where(y > 0, s - X, s),
                     np.where(x < X - 1, s + 1, s),
                     np.where(y < Y - 1, s + X, s),
                     np.where(x > 0, s - 1, s)], axis=1)
    nxt = np.where(level.wall[cand], s[:, None], cand)
    if care_about_terminal:
        nxt = np.where(level.term[:, None], s[:, None], nxt)
    return nxt


def look_step_ahead_batch(level, states, actions, care_about_terminal=True):
    """Batched a6: arbitrary (states[M], actions[M]) on one level."""
    states = np.asarray(states, dtype=np.int64)
    actions = np.asarray(actions, dtype=np.int64)
    nxt = next_table(level, care_about_terminal)[states, actions]
    return nxt, level.reward[nxt], level.term[nxt]


def rollout(levels, pos0, actions, auto_reset=False, start_choice=None):
    """Batched griduniverse_env.py:176-193 for N envs over T steps.

    ``levels``: one Level (shared) or a list of N Levels.  ``actions`` int[T,N].
    Reference semantics: a terminal state is absorbing (:145-146), ``step``
    returns the landing cell.  ``auto_reset=True`` restates the callers'
    ``if done: env.reset()`` (examples/griduniverse_env_examples.py:15,22-24):
    the returned observation is still the landing cell, the env then continues
    from its start state.  With several start states the choice stream is
    host-supplied: ``start_choice`` int[T,N] holds the start *state* to use if
    env n resets after step t (default: the level's first start).
    Returns obs[T,N], reward[T,N], done[T,N], final_pos[N].
    """
    actions = np.asarray(actions, dtype=np.int64)
    T, N = actions.shape
    shared = isinstance(levels, Level)
    pos = np.array(pos0, dtype=np.int64).copy()
    obs = np.zeros((T, N), dtype=np.int64)
    rew = np.zeros((T, N), dtype=np.int64)
    done = np.zeros((T, N), dtype=bool)
    if shared:
        tables = next_table(levels)
        first_start = np.full(N, levels.starts[0], dtype=np.int64)
    else:
        assert len(levels) == N
        tables = [next_table(lv) for lv in levels]
        first_start = np.array([lv.starts[0] for lv in levels], dtype=np.int64)
    for t in range(T):
        if shared:
            nxt = tables[pos, actions[t]]
            r = levels.reward[nxt]
            d = levels.term[nxt]
        else:
            nxt = np.array([tables[n][pos[n], actions[t, n]] for n in range(N)], dtype=np.int64)
            r = np.array([levels[n].reward[nxt[n]] for n in range(N)], dtype=np.int64)
            d = np.array([levels[n].term[nxt[n]] for n in range(N)], dtype=bool)
        obs[t], rew[t], done[t] = nxt, r, d
        pos = nxt
        if auto_reset:
            st = first_start if start_choice is None else np.asarray(start_choice[t], dtype=np.int64)
            pos = np.where(d, st, pos)
    return obs, rew, done, pos


# --------------------------------------------------------------------------
# Bellman sweep / greedy extraction
# --------------------------------------------------------------------------
def _ftype(dtype):
    dt = np.dtype(dtype)
    assert dt in (np.dtype(np.float64), np.dtype(np.float32))
    return dt.type


def sweep(level, policy, v, gamma=1.0, dtype=np.float64, nxt=None):
    """core/algorithms/utils.py:15-27 -- one synchronous policy-evaluation sweep.

    ``v_new[s] = ((((0 + R[s]) + p0*(g*v[n0])) + p1*(g*v[n1])) + p2*..) + p3*..``
    accumulated left to right exactly like utils.py:23-26; ``dtype=float32``
    evaluates the same expression in IEEE single (no fused multiply-add).
    """
    F = _ftype(dtype)
    if nxt is None:
        nxt = next_table(level)
    v = np.asarray(v, dtype=F)
    policy = np.asarray(policy, dtype=F)
    g = F(gamma)
    acc = np.zeros(level.N, dtype=F) + level.reward.astype(F)
    for a in range(4):
        acc = acc + policy[:, a] * (g * v[nxt[:, a]])
    return acc


def greedy_masks(level, v, gamma=1.0, dtype=np.float64, nxt=None):
    """core/algorithms/utils.py:62-71 as a 4-bit tie-set mask per state.

    bit a set <=> around(q[s,a],8) == around(max q[s],8) and s not terminal;
    ``np.around(x, 8)`` is ``rint(x*1e8)/1e8`` so ties are decided on
    ``rint(q*1e8)``.  q uses the reward of the *landing* state (utils.py:65-66).
    """
    F = _ftype(dtype)
    if nxt is None:
        nxt = next_table(level)
    v = np.asarray(v, dtype=F)
    g = F(gamma)
    q = level.reward[nxt].astype(F) + g * v[nxt]          # 0.0 + (reward + g*v)
    r = np.rint(q * F(1e8))
    tie = r == r.max(axis=1, keepdims=True)
    tie &= ~level.term[:, None]
    return (tie * np.array([1, 2, 4, 8])).sum(axis=1).astype(np.uint8)


def masks_to_policy(masks, dtype=np.float64):
    """Expand tie masks to the [N,4] rows utils.py:69-71 writes (1/len on ties)."""
    masks = np.asarray(masks, dtype=np.uint8)
    bits = ((masks[:, None] >> np.arange(4)) & 1).astype(dtype)
    cnt = bits.sum(axis=1, keepdims=True)
    with np.errstate(divide='ignore', invalid='ignore'):
        p = np.where(cnt > 0, bits * (np.asarray(1, dtype=dtype) / cnt), 0)
    return p.astype(dtype)


def policy_to_masks(policy):
    """Inverse of masks_to_policy, or None if policy is not a uniform-on-subset policy."""
    policy = np.asarray(policy, dtype=np.float64)
    bits = policy > 0
    masks = (bits * np.array([1, 2, 4, 8])).sum(axis=1).astype(np.uint8)
    if np.array_equal(masks_to_policy(masks), policy):
        return masks
    return None


def greedy_policy_from_value_function(policy, level, v, gamma=1.0, dtype=np.float64):
    """utils.py:55-72 -- writes into the caller's ``policy`` and returns it."""
    policy[...] = masks_to_policy(greedy_masks(level, v, gamma, dtype))
    return policy


def greedy_action(masks):
    """np.argmax(policy[s]) (examples/griduniverse_alg_examples.py:76,121): lowest set bit, 0 if none."""
    masks = np.asarray(masks, dtype=np.uint8)
    out = np.zeros(masks.shape, dtype=np.int64)
    for a in (3, 2, 1, 0):
        out = np.where((masks >> a) & 1, a, out)
    return out


def value_iteration(policy, level, value_function=None, threshold=0.00001, max_steps=1000,
                    discount_factor=1.0, dtype=np.float64):
    """core/algorithms/dynamic_programming.py:8-28.  Returns (V, policy, sweeps)."""
    F = _ftype(dtype)
    nxt = next_table(level)
    v = np.zeros(level.N, dtype=F) if value_function is None else np.asarray(value_function, dtype=F)
    greedy = policy
    sweeps = 0
    for step_number in range(max_steps):
        v_new = sweep(level, greedy, v, discount_factor, dtype, nxt)
        delta = np.max(v - v_new)             # signed, not abs (:17)
        v = v_new
        sweeps += 1
        greedy[...] = masks_to_policy(greedy_masks(level, v, discount_factor, dtype, nxt))
        if delta < threshold:
            break
        elif step_number == max_steps - 1:
            warnings.warn('Value iteration did not reach the selected threshold. Finished after reaching '
                          'the maximum {} steps'.format(step_number + 1), UserWarning)
    return v, greedy, sweeps


def policy_iteration(policy, level, value_function=None, threshold=0.00001, max_steps=1000,
                     discount_factor=1.0, dtype=np.float64):
    """core/algorithms/dynamic_programming.py:31-57.  Returns (V_lastconv, policy, sweeps)."""
    F = _ftype(dtype)
    nxt = next_table(level)
    v = last = np.zeros(level.N, dtype=F) if value_function is None else np.asarray(value_function, dtype=F)
    greedy = policy
    sweeps = 0
    for step_number in range(max_steps):
        v_new = sweep(level, greedy, v, discount_factor, dtype, nxt)
        delta_eval = np.max(v - v_new)
        v = v_new
        sweeps += 1
        if delta_eval < threshold:
            # utils.greedy writes into the *same* array object greedy_policy refers to (:43,
            # utils.py:69), so the policy changes in place even when the loop then breaks.
            new_masks = greedy_masks(level, v, discount_factor, dtype, nxt)
            greedy[...] = masks_to_policy(new_masks)
            delta = np.max(last - v_new)
            last = v_new
            if delta < threshold:
                break
        elif step_number == max_steps - 1:
            greedy[...] = masks_to_policy(greedy_masks(level, last, discount_factor, dtype, nxt))
            warnings.warn('Policy iteration did not reach the selected threshold. Finished after reaching '
                          'the maximum {} steps with delta_eval {}'.format(step_number + 1, delta_eval),
                          UserWarning)
    return last, greedy, sweeps


# --------------------------------------------------------------------------
# Monte-Carlo (core/algorithms/monte_carlo.py)
# --------------------------------------------------------------------------
def run_episode(policy, level, start, max_steps_per_episode=1000, rng=np.random):
    """monte_carlo.py:7-26 with the start state host-supplied (reset's random.choice,
    griduniverse_env.py:189) and actions drawn by the same ``np.random.choice`` call (:20)."""
    states, rewards = [int(start)], []
    obs, done = int(start), False
    for _ in range(max_steps_per_episode):
        a = rng.choice(policy[obs].size, p=policy[obs])
        obs, r, done = look_step_ahead(level, obs, int(a))
        states.append(obs)
        rewards.append(r)
        if done:
            break
    return states, rewards, done


def mc_accumulate_episode(states, rewards, N, every_visit, discount_factor, threshold):
    """monte_carlo.py:53-71: per-episode visit counts and truncated discounted returns."""
    visits = np.zeros(N)
    returns = np.zeros(N)
    for idx, s in enumerate(states):
        if visits[s] == 0:
            pass
        elif not every_visit:
            continue
        visits[s] += 1
        # monte_carlo.py:69-70.  Plain left-to-right adds = `sum` of the reference's pinned
        # CPython 3.6 (CPython >= 3.12 compensates float sums; see tests/golden/make_golden.py).
        g = 0
        for i, r in enumerate(rewards[idx:]):
            if (discount_factor ** i) > threshold:
                g = g + (discount_factor ** i) * r
        returns[s] += g
    return visits, returns


def monte_carlo_evaluation(policy, level, starts_stream, every_visit=False, incremental_mean=True,
                           stationary_env=True, discount_factor=0.99, threshold=0.0001, alpha=0.001,
                           num_episodes=100, rng=np.random, episodes=None):
    """monte_carlo.py:29-99.  ``episodes`` (list of (states, rewards)) replays recorded rollouts."""
    N = level.N
    total_visits = np.zeros(N)
    total_return = np.zeros(N)
    V = np.zeros(N)
    for ep in range(num_episodes):
        if episodes is not None:
            st, rw = episodes[ep]
        else:
            st, rw, _ = run_episode(policy, level, starts_stream[ep], rng=rng)
        visits, returns = mc_accumulate_episode(st, rw, N, every_visit, discount_factor, threshold)
        for s in range(N):
            total_visits[s] += visits[s]
            if not incremental_mean:
                total_return[s] += returns[s]
            else:
                if stationary_env:
                    if total_visits[s] > 0.0:
                        V[s] += (1 / total_visits[s]) * (returns[s] - V[s])
                else:
                    V[s] += alpha * (returns[s] - V[s])
    if not incremental_mean:
        for s in range(N):
            if total_visits[s] > 0.0:
                V[s] = total_return[s] / total_visits[s]
    return V


# --------------------------------------------------------------------------
# Shortest paths (core/algorithms/maze_solving.py)
# --------------------------------------------------------------------------
def bfs_graph(level):
    """create_graph, maze_solving.py:43-50: for every non-wall state the list of states reached
    by the four actions with care_about_terminal=False, in action order, self-loops dropped."""
    nxt = next_table(level, care_about_terminal=False)
    graph = {}
    for s in range(level.N):
        if not level.wall[s]:
            graph[s] = [int(n) for n in nxt[s] if n != s]
    return graph


def bfs_reference_path(level, start):
    """breadth_first_search + construct_path + calculate_action, maze_solving.py:110-193: FIFO
    search from ``start``, stops at the first terminal dequeued, returns the action list (None if
    no terminal is reachable -- the script then has nothing to return either)."""
    graph = bfs_graph(level)
    open_set, closed_set, meta = [start], set(), {start: (None, None)}
    while open_set:
        parent = open_set.pop(0)
        if level.term[parent]:
            actions = []
            state = parent
            while meta[state][0] is not None:
                state, a = meta[state]
                actions.append(a)
            actions.reverse()
            return actions
        for child in graph[parent]:
            if child in closed_set or child in open_set:
                continue
            diff = parent - child        # calculate_action, :110-121
            a = LEFT if diff == 1 else RIGHT if diff == -1 else DOWN if diff < -1 else UP
            meta[child] = (parent, a)
            open_set.append(child)
        closed_set.add(parent)
    return None


def bfs_distances(level, sources, lava_blocks=False):
    """Distance (number of actions) from every state to the nearest of ``sources`` over the
    graph of bfs_graph; -1 where no source is reachable and on walls.  ``lava_blocks`` also
    removes the lava cells from the graph.  Queue version (small levels)."""
    from collections import deque
    graph = bfs_graph(level)
    blocked = level.wall | level.lava if lava_blocks else level.wall
    dist = np.full(level.N, -1, dtype=np.int32)
    queue = deque()
    for s in sources:
        if not blocked[s] and dist[s] < 0:
            dist[s] = 0
            queue.append(int(s))
    while queue:
        s = queue.popleft()
        for n in graph[s]:
            if dist[n] < 0 and not blocked[n]:
                dist[n] = dist[s] + 1
                queue.append(n)
    return dist


def bfs_distances_dense(level, sources, lava_blocks=False):
    """Same result as bfs_distances by whole-grid wavefront expansion (bigger levels)."""
    X, Y = level.X, level.Y
    blocked = (level.wall | level.lava if lava_blocks else level.wall).reshape(Y, X)
    seen = np.zeros((Y, X), dtype=bool)
    seen.reshape(-1)[np.asarray(list(sources), dtype=np.int64)] = True
    seen &= ~blocked
    dist = np.where(seen, 0, -1).astype(np.int32)
    d = 0
    while True:
        d += 1
        grow = np.zeros_like(seen)
        grow[1:] |= seen[:-1]
        grow[:-1] |= seen[1:]
        grow[:, 1:] |= seen[:, :-1]
        grow[:, :-1] |= seen[:, 1:]
        new = grow & ~seen & ~blocked
        if not new.any():
            return dist.reshape(-1)
        dist[new] = d
        seen |= new


def bfs_descent_path(level, dist, start):
    """Shortest path from ``start`` down a distance field: lowest-numbered action that lands one
    level closer (the np.argmax order, examples/griduniverse_alg_examples.py:76)."""
    if dist[start] < 0:
        return None
    s, actions = int(start), []
    while dist[s] > 0:
        for a in range(4):
            n = clamp_move(level, s, a)
            if n != s and dist[n] == dist[s] - 1:
                actions.append(a)
                s = n
                break
        else:
            raise ValueError("not a distance field of this level")
    return actions


def value_from_goal_distance(d, gamma=0.9):
    """Fixed point of value_iteration (dynamic_programming.py:8-28) in closed form for a state
    ``d`` actions from the nearest goal along non-lava cells: d step rewards of -1 discounted by
    gamma (utils.py:23 pays the reward of the CURRENT state), then the goal's own value 10 (its
    greedy row is all zero, utils.py:69-71).  A state that cannot reach a goal pays -1 forever."""
    d = np.asarray(d, dtype=np.float64)
    reach = d >= 0
    gd = np.power(gamma, np.where(reach, d, 0.0))
    return np.where(reach, -(1.0 - gd) / (1.0 - gamma) + gd * REWARD_GOAL, REWARD_STEP / (1.0 - gamma))


# --------------------------------------------------------------------------
# ASCII render (griduniverse_env.py:202-221) -- host glue, kept for the unit-test goldens
# --------------------------------------------------------------------------
def render_ansi(level, current_state):
    cells = ['o'] * level.N
    cells[current_state] = 'x'
    for s in np.flatnonzero(level.goal):
        cells[s] = 'G'
    for s in np.flatnonzero(level.lava):
        cells[s] = 'L'
    for s in np.flatnonzero(level.wall):
        cells[s] = '#'
    out = []
    for y in range(level.Y):
        out.append(''.join(c + ' ' for c in cells[y * level.X:(y + 1) * level.X]) + '\n')
    return ''.join(out) + '\n'
