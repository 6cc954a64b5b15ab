from . import utils, dynamic_programming, monte_carlo  # noqa: F401
