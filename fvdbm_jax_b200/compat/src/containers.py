"""reference src/containers.py -> fvdbm_jax_b200.containers (plus the names that module star-exports)."""
import numpy as np  # noqa: F401
from fvdbm_jax_b200.containers import Container, Cells, Faces, Nodes  # noqa: F401
from fvdbm_jax_b200.dynamics import Dynamics  # noqa: F401
from utils.utils import *  # noqa: F401,F403
from utils.utils import __all__ as _utils_names

__all__ = ["Container", "Cells", "Faces", "Nodes", "Dynamics"] + list(_utils_names)
