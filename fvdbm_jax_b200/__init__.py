"""fvdbm_jax_b200 -- B200-native drop-in for FVDBM-JAX's ``Environment.step()`` hot path.

(The repo brief calls the package ``fvdbm-jax_b200``; a hyphen is not importable, hence the
underscore.)  Public surface mirrors the reference modules:

    reference                      here
    src/environment.py Environment fvdbm_jax_b200.Environment
    src/containers.py  Cells/...   fvdbm_jax_b200.Cells / Faces / Nodes / CustomArray
    src/dynamics.py    D2Q9/D2Q13  fvdbm_jax_b200.D2Q9 / D2Q13
    src/mesher.py      Mesher      fvdbm_jax_b200.Mesher (vectorised producer of the statics)
"""
from .dynamics import Dynamics, D2Q9, D2Q13
from .containers import CustomArray, Container, Cells, Faces, Nodes
from .environment import Environment
from .mesher import Mesher
from . import meshgen

__all__ = ["Dynamics", "D2Q9", "D2Q13", "CustomArray", "Container", "Cells", "Faces", "Nodes",
           "Environment", "Mesher", "meshgen"]
__version__ = "0.1.0"
