from .griduniverse_env import GridUniverseEnv  # noqa: F401
from .vec_env import GridUniverseVecEnv  # noqa: F401
