"""reference src/dynamics.py -> fvdbm_jax_b200.dynamics."""
from fvdbm_jax_b200.dynamics import Dynamics, D2Q9, D2Q13  # noqa: F401

__all__ = ["Dynamics", "D2Q9", "D2Q13"]
