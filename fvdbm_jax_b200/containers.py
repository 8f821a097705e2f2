"""Host-side containers with the attribute surface of /root/reference/src/containers.py.

``Cells`` / ``Faces`` / ``Nodes`` only *describe* a problem (NumPy arrays in the reference's AoS
shapes); all arithmetic of the reference methods (calc_macros, calc_eqs, calc_pdfs, calc_fluxes,
Nodes.calc_pdfs) lives in the CUDA kernels behind ``Environment.step``.  ``Environment`` also
accepts the reference's own container objects or any duck-typed object with these attributes.
"""
from __future__ import annotations

import numpy as np

from .dynamics import Dynamics

__all__ = ["CustomArray", "Container", "Cells", "Faces", "Nodes"]


class _AtIndex:
    """``x.at[idx]`` -> functional updates that return a changed COPY, the JAX idiom the reference notebooks use on
    container attributes while hand-building a mesh (tests/ldcFVDBM.ipynb c6-c7:
    ``env.nodes.type = env.nodes.type.at[i].set(1)``, ``env.faces.n = env.faces.n.at[j].set(n)``)."""

    def __init__(self, arr, idx=None):
        self._arr, self._idx = arr, idx

    def __getitem__(self, idx):
        return _AtIndex(self._arr, idx)

    def _updated(self, fn):
        out = np.array(self._arr, copy=True)
        out[self._idx] = fn(out[self._idx])
        return out.view(Array)

    def set(self, values):
        return self._updated(lambda _: np.asarray(values))

    def add(self, values):
        return self._updated(lambda cur: cur + np.asarray(values))

    def multiply(self, values):
        return self._updated(lambda cur: cur * np.asarray(values))

    def min(self, values):
        return self._updated(lambda cur: np.minimum(cur, np.asarray(values)))

    def max(self, values):
        return self._updated(lambda cur: np.maximum(cur, np.asarray(values)))

    def get(self):
        return np.asarray(self._arr)[self._idx]


class Array(np.ndarray):
    """NumPy array with JAX's ``.at[...]`` functional-update property; what the containers hold before the engine is
    built, so that notebook code written against ``jax.Array`` attributes keeps working.  ``np.asarray`` strips it."""

    @property
    def at(self):
        return _AtIndex(self)


def _arr(a):
    return np.asarray(a).view(Array)


class CustomArray:
    """Padded ragged array used while hand-building meshes (reference utils/utils.py:176-230)."""

    def __init__(self, size, dtype=np.float64, default_value=-1):
        self.data = default_value * np.ones((size, 1), dtype=dtype)
        self.default_value = default_value

    def add_item(self, index, item):
        row = self.data[index]
        free = np.nonzero(row == self.default_value)[0]
        if free.size:
            self.data[index, free[0]] = item
        else:
            col = -np.ones_like(self.data[..., 0:1])
            col[index] = item
            self.data = np.concatenate((self.data, col), axis=-1)

    def add_items(self, index, items):
        for item in np.asarray(items).reshape(-1):
            self.add_item(index, item)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data, dtype=dtype)

    def __getitem__(self, idx):
        return self.data[idx]

    def shape(self):
        return self.data.shape

    def dtype(self):
        return self.data.dtype

    def __repr__(self):
        return f"CustomArray(size={self.data.shape[0]}, data={self.data})"


class Container:
    """reference containers.py:19-45: ``pdf`` starts at the rest equilibrium feq(rho=1, u=0) = W."""
    dynamics: Dynamics

    def __init__(self, size, dynamics: Dynamics):
        eq = dynamics.calc_eq(np.float64(1), np.zeros(dynamics.DIM))
        # a real C-contiguous array (not np.broadcast_to: copies of a broadcast view come out Fortran-ordered), filled by
        # broadcast assignment, which is faster than np.repeat at 10^7 rows
        eq = np.asarray(eq)
        self.pdf = np.empty((size,) + eq.shape, dtype=eq.dtype).view(Array)
        self.pdf[...] = eq
        self.dynamics = dynamics

    def __repr__(self):
        return repr(self.__dict__)


class Cells(Container):
    """reference containers.py:48-134"""

    def __init__(self, size, dynamics: Dynamics):
        super().__init__(size, dynamics)
        self.rho = _arr(np.zeros((size, 1), dtype=np.float64))
        self.vel = _arr(np.zeros((size, dynamics.DIM), dtype=np.float64))
        self.pdf_eq = _arr(np.zeros((size, dynamics.NUM_QUIVERS), dtype=np.float64))
        self.face_indices = CustomArray(size, dtype=np.int32, default_value=-1)
        self.face_normals = CustomArray(size, dtype=np.int32, default_value=-1)

    def init(self):
        self.face_indices = np.asarray(self.face_indices)
        self.face_normals = np.asarray(self.face_normals)


class Faces(Container):
    """reference containers.py:137-291 (``pdf`` holds the face flux)."""

    def __init__(self, size, dynamics: Dynamics, flux_scheme: str = "upwind"):
        super().__init__(size, dynamics)
        self.nodes_index = CustomArray(size, dtype=np.int32, default_value=-1)
        self.stencil_cells_index = CustomArray(size, dtype=np.int32, default_value=-1)
        self.stencil_dists = CustomArray(size, dtype=np.float64, default_value=-1)
        self.n = _arr(np.zeros((size, dynamics.DIM), dtype=np.float64))
        self.L = _arr(np.zeros((size, 1), dtype=np.float64))
        self.flux_scheme = flux_scheme

    def init(self):
        self.nodes_index = np.asarray(self.nodes_index)
        self.stencil_cells_index = np.asarray(self.stencil_cells_index)
        self.stencil_dists = np.asarray(self.stencil_dists)


class CCStencilFaces(Faces):
    """reference src/faces.py:10-76: ``n`` is the cell-centre stencil direction, ``alpha`` its angle to the face normal;
    the flux is the plain scheme's times cos(alpha) (Environment folds it into the face length)."""

    def __init__(self, size, dynamics: Dynamics, flux_scheme: str = "upwind"):
        super().__init__(size, dynamics, flux_scheme=flux_scheme)
        self.alpha = np.zeros((size, 1), dtype=np.float64)

    def init(self):
        super().init()
        self.alpha = np.asarray(self.alpha)


class CCStencilKsiFaces(Faces):
    """reference src/faces.py:78-141 (``cc_alt_upwind``): divides by KSI.n_PQ, which is 0 for axis-aligned stencils, so the
    reference itself produces NaN; ``Environment`` refuses it with a ValueError.  Present for import compatibility."""

    def __init__(self, size, dynamics: Dynamics, flux_scheme: str = "upwind"):
        super().__init__(size, dynamics, flux_scheme=flux_scheme)
        self.npq = np.zeros((size, dynamics.DIM), dtype=np.float64)

    def init(self):
        super().init()
        self.npq = np.asarray(self.npq)


class Nodes(Container):
    """reference containers.py:294-408"""

    def __init__(self, size, dynamics: Dynamics):
        super().__init__(size, dynamics)
        self.rho = _arr(np.zeros((size, 1), dtype=np.float64))
        self.vel = _arr(np.zeros((size, dynamics.DIM), dtype=np.float64))
        self.type = _arr(np.zeros((size, 1), dtype=np.int32))
        self.cells_index = CustomArray(size, dtype=np.int32, default_value=-1)
        self.cell_dists = CustomArray(size, dtype=np.float64, default_value=-1)

    def init(self):
        self.cells_index = np.asarray(self.cells_index)
        self.cell_dists = np.asarray(self.cell_dists)
