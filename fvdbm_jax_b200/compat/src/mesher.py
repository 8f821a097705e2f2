"""reference src/mesher.py -> fvdbm_jax_b200.Mesher (plus what that module star-exports)."""
import pickle  # noqa: F401
import time  # noqa: F401
from src.containers import *  # noqa: F401,F403
from src.dynamics import *  # noqa: F401,F403
from src.faces import *  # noqa: F401,F403
from src.environment import *  # noqa: F401,F403
from src.environment import __all__ as _e
from src.faces import __all__ as _f
from fvdbm_jax_b200.mesher import Mesher  # noqa: F401

__all__ = sorted(set(_e) | set(_f) | {"D2Q13", "Mesher", "pickle", "time"})
