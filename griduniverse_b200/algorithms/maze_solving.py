"""Shortest paths on a GridUniverse level (core/algorithms/maze_solving.py of the reference).

The reference file is a script: it builds the adjacency list of the non-wall cells with
``look_step_ahead(s, a, False)`` (:43-50) and searches it first-in-first-out from the start state
until a terminal is dequeued (:123-168), then lists the actions back to the start (:170-193).
Here the search is a wavefront over bit planes on the GPU (csrc/gu_bfs.cu).  The action list has
the reference's length and ends on a terminal at the same distance; where several shortest paths
exist the reference returns the one its queue order happens to find, this returns the one that
takes the lowest-numbered action at every step.
"""
import numpy as np

from ..device import PlanGrid
from ..paths import ShortestPaths
from .utils import level_of


def _solver(env, device=None):
    level = level_of(env)
    return level, ShortestPaths(PlanGrid(level, np.float32, device or getattr(env, "_device", "cuda")))


def shortest_distances(env, sources=None, lava_blocks=False):
    """int32[N]: number of actions from every state to the nearest source (default: the goal
    states); -1 for walls and for states that cannot reach one."""
    level, sp = _solver(env)
    if sources is None:
        sources = np.flatnonzero(level.goal)
    dist = sp.solve(sources, lava_blocks=lava_blocks)
    return sp.grid.dense(dist).cpu().numpy()


def breadth_first_search(env, start_state=None):
    """Action list from ``start_state`` (default: the env's current initial state) to the nearest
    terminal state, walls blocking -- maze_solving.py:123-168,201.  None if no terminal is
    reachable."""
    level, sp = _solver(env)
    if start_state is None:
        start_state = getattr(env, "initial_state", level.starting_states[0])
    sp.solve(np.flatnonzero(level.goal | level.lava))
    return sp.walk(start_state)
