from . import utils, dynamic_programming, monte_carlo, maze_solving  # noqa: F401
