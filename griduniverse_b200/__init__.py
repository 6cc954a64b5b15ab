"""griduniverse_b200 -- B200-native GridUniverse simulation and tabular-planning core.

Drop-in for the hot path of TheMTank/GridUniverse (``core.envs`` / ``core.algorithms``):

    from griduniverse_b200.envs import GridUniverseEnv, GridUniverseVecEnv
    from griduniverse_b200.algorithms import utils, monte_carlo
    import griduniverse_b200.algorithms.dynamic_programming as dp

All arithmetic on the path runs in hand-written sm_100a CUDA kernels reached through the
C ABI in include/gu_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"
