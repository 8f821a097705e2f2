"""reference src/environment.py -> fvdbm_jax_b200.Environment (the sm_100a engine behind init()/step())."""
from src.containers import *  # noqa: F401,F403
from src.containers import __all__ as _c
from fvdbm_jax_b200.dynamics import D2Q9  # noqa: F401
from fvdbm_jax_b200.environment import Environment  # noqa: F401

__all__ = list(_c) + ["D2Q9", "Environment"]
