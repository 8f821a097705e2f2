"""reference src/faces.py -> the cell-centre-stencil face classes."""
from src.containers import *  # noqa: F401,F403
from src.containers import __all__ as _c
from fvdbm_jax_b200.containers import CCStencilFaces, CCStencilKsiFaces  # noqa: F401

__all__ = list(_c) + ["CCStencilFaces", "CCStencilKsiFaces"]
