"""reference utils/utils.py -> fvdbm_jax_b200.utils (NumPy helpers + CustomArray)."""
import numpy as np  # noqa: F401  (the reference module star-exports np as well)
from fvdbm_jax_b200.utils import *  # noqa: F401,F403
from fvdbm_jax_b200.utils import __all__ as _names

__all__ = list(_names) + ["np"]
