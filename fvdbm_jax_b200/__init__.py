"""fvdbm_jax_b200 -- B200-native drop-in for FVDBM-JAX's ``Environment.step()`` hot path.

(The repo brief calls the package ``fvdbm-jax_b200``; a hyphen is not importable, hence the
underscore.)  Public surface mirrors the reference modules:

    reference                      here
    src/environment.py Environment fvdbm_jax_b200.Environment
    src/containers.py  Cells/...   fvdbm_jax_b200.Cells / Faces / Nodes / CustomArray
    src/dynamics.py    D2Q9/D2Q13  fvdbm_jax_b200.D2Q9 / D2Q13
    src/mesher.py      Mesher      fvdbm_jax_b200.Mesher (native / vectorised producer of the statics)
    src/faces.py       CCStencil*  fvdbm_jax_b200.CCStencilFaces / CCStencilKsiFaces
    utils/utils.py     helpers     fvdbm_jax_b200.utils
fvdbm_jax_b200.compat.install() puts packages named ``src`` / ``utils`` with these contents first on sys.path, so the
reference notebooks' own ``from src.environment import *`` lines resolve here without being edited.
"""
from .dynamics import Dynamics, D2Q9, D2Q13
from .containers import CustomArray, Container, Cells, Faces, Nodes, CCStencilFaces, CCStencilKsiFaces
from .environment import Environment
from .mesher import Mesher
from . import meshgen

__all__ = ["Dynamics", "D2Q9", "D2Q13", "CustomArray", "Container", "Cells", "Faces", "Nodes", "CCStencilFaces",
           "CCStencilKsiFaces", "Environment", "Mesher", "meshgen"]
__version__ = "0.1.0"
