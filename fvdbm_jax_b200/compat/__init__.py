"""Import-level drop-in for the reference's own module names.

The reference notebooks begin with ``sys.path.append('..')`` followed by ``from src.environment import *``,
``from src.mesher import *``, ``from src.dynamics import *`` ... (tests/flow_over_cyl.ipynb c1, tests/ldcFVDBM.ipynb c1,
tests/porous_flow.ipynb c1).  This directory holds packages called ``src`` and ``utils`` whose modules carry the same
names and re-export this framework's classes, so ONE added line ahead of those imports switches a notebook over:

    import fvdbm_jax_b200.compat; fvdbm_jax_b200.compat.install()

``install()`` puts this directory first on ``sys.path`` (and forgets any ``src`` / ``utils`` modules imported before);
``uninstall()`` undoes it.  Nothing here computes anything: every name is the one exported by ``fvdbm_jax_b200``.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
_SHADOWED = ("src", "utils")


def _forget():
    for name in [m for m in sys.modules if m in _SHADOWED or m.startswith(tuple(s + "." for s in _SHADOWED))]:
        del sys.modules[name]


def install():
    """Make ``import src.*`` / ``import utils.*`` resolve to this framework."""
    if HERE in sys.path:
        sys.path.remove(HERE)
    sys.path.insert(0, HERE)
    _forget()
    return HERE


def uninstall():
    if HERE in sys.path:
        sys.path.remove(HERE)
    _forget()
