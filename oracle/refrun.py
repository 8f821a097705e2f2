"""Run the UNMODIFIED reference sources (/root/reference/src, utils) under the NumPy `jax` shim.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build container); it is
used by oracle/make_golden.py to mint tests/golden/*.npz and by tests that pin the NumPy
restatement (oracle/step_numpy.py) and the vectorised Mesher against the reference's own code.
Never imported by the product package, bench.py's GPU arm or the -m gpu tests.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("FVDBM_REFERENCE", "/root/reference")
_SHIM = os.path.join(HERE, "jaxshim")
_mods = None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "src"))


def load():
    """Import the reference modules once; returns a namespace with Mesher, Environment, D2Q9, ..."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE}")
    for p in (REFERENCE, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    # make sure a real jax (if one ever appears) does not shadow the shim for this process
    for name in [m for m in sys.modules if m == "jax" or m.startswith("jax.")]:
        if not getattr(sys.modules[name], "__shim__", False) and name == "jax":
            raise RuntimeError("a non-shim jax is already imported")
    ns = types.SimpleNamespace()
    ns.jax = importlib.import_module("jax")
    assert getattr(ns.jax, "__shim__", False)
    ns.jnp = importlib.import_module("jax.numpy")
    ns.dynamics = importlib.import_module("src.dynamics")
    ns.containers = importlib.import_module("src.containers")
    ns.environment = importlib.import_module("src.environment")
    ns.mesher = importlib.import_module("src.mesher")
    ns.utils = importlib.import_module("utils.utils")
    ns.Mesher = ns.mesher.Mesher
    ns.Environment = ns.environment.Environment
    ns.D2Q9, ns.D2Q13 = ns.dynamics.D2Q9, ns.dynamics.D2Q13
    ns.float_dtype = np.dtype(ns.jnp.float32)
    _mods = ns
    return ns


def ref_mesher(raw):
    """Reference Mesher over a RawMesh-like object (points/elements/faces/point_markers)."""
    ns = load()
    m = ns.Mesher()
    with contextlib.redirect_stdout(io.StringIO()):
        m.import_meshpy(raw)
        m.calc_mesh_properties()
    return m


MESHER_FIELDS = ("points", "cells", "faces", "point_markers", "cell_centers", "face_centers",
                 "face_normals", "face_lengths", "cell_face_indices", "cell_face_normal_signs",
                 "face_cell_indices", "face_cell_center_distances", "point_cell_indices",
                 "point_cell_center_distances", "stencil_norms", "cc_stencil_dist", "face_stencil_angles")


def mesher_arrays(m) -> dict:
    return {k: np.array(getattr(m, k)) for k in MESHER_FIELDS}


STATE = ("cells.pdf", "cells.rho", "cells.vel", "cells.pdf_eq", "faces.pdf",
         "nodes.pdf", "nodes.rho", "nodes.vel")
STATIC = ("cells.face_indices", "cells.face_normals", "faces.nodes_index",
          "faces.stencil_cells_index", "faces.stencil_dists", "faces.n", "faces.L",
          "nodes.type", "nodes.cells_index", "nodes.cell_dists")


def _get(env, dotted):
    obj, attr = dotted.split(".")
    return np.array(np.asarray(getattr(getattr(env, obj), attr)))


def snapshot(env, names=STATE) -> dict:
    out = {n: _get(env, n) for n in names}
    if names is STATIC and hasattr(env.faces, "alpha"):            # CCStencilFaces (src/faces.py:10-76)
        out["faces.alpha"] = _get(env, "faces.alpha")
    return out


def run_steps(env, checkpoints):
    """Step the reference Environment; returns {step: {state arrays}} at the given step counts."""
    out = {}
    done = 0
    for target in sorted(checkpoints):
        while done < target:
            env = env.step()
            done += 1
        out[target] = snapshot(env)
    return out
