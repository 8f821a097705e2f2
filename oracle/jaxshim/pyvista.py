"""Empty stub: src/mesher.py imports pyvista at module top; only to_vtk uses it."""


class CellType:
    TRIANGLE = 5


class UnstructuredGrid:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("pyvista is stubbed in the oracle shim")
