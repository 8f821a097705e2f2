"""Locality renumbering of cells: explicit permutations, never applied behind the user's back.

``perm[i]`` = storage rank of original cell ``i`` (a bijection; results are returned in the original
numbering, so renumbering is invisible at the API -- tests/test_reorder.py checks bijectivity and
permutation invariance).  The reference has no renumbering; north_star asks for
"RCM / space-filling-curve ordering".
"""
from __future__ import annotations

import numpy as np

__all__ = ["hilbert_perm", "rcm_perm", "order_to_perm", "identity_perm", "choose_perm"]


def order_to_perm(order: np.ndarray) -> np.ndarray:
    """order[r] = original cell stored at rank r  ->  perm[i] = rank of original cell i."""
    perm = np.empty(order.shape[0], dtype=np.int32)
    perm[order] = np.arange(order.shape[0], dtype=np.int32)
    return perm


def identity_perm(n: int) -> np.ndarray:
    return np.arange(n, dtype=np.int32)


def hilbert_index(x: np.ndarray, y: np.ndarray, bits: int = 16) -> np.ndarray:
    """Hilbert-curve index of integer grid points (x,y) in [0,2^bits)^2, vectorised."""
    x = x.astype(np.int64).copy()
    y = y.astype(np.int64).copy()
    d = np.zeros(x.shape, dtype=np.int64)
    s = np.int64(1) << (bits - 1)
    n1 = (np.int64(1) << bits) - 1
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        # rotate quadrant
        flip = (ry == 0) & (rx == 1)
        x = np.where(flip, n1 - x, x)
        y = np.where(flip, n1 - y, y)
        swap = ry == 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        s >>= 1
    return d


def hilbert_perm(centers: np.ndarray, bits: int = 16) -> np.ndarray:
    """Space-filling-curve ordering from cell centroids (N,2)."""
    c = np.ascontiguousarray(centers, dtype=np.float64)
    lo = c.min(axis=0)
    span = np.maximum(c.max(axis=0) - lo, 1e-300)
    scale = ((1 << bits) - 1) / span.max()
    key = None
    if c.shape[0] >= 1 << 16:            # large meshes: native OpenMP helper of the C-ABI library (same curve, same keys)
        try:
            from . import _lib
            key = np.empty(c.shape[0], dtype=np.int64)
            _lib.check(_lib.load().fvdbm_sfc_keys(c.ctypes.data, None, c.shape[0], 3, bits, float(lo[0]), float(lo[1]), float(scale),
                                                  key.ctypes.data))
        except (RuntimeError, OSError):
            key = None
    if key is None:
        g = np.minimum(((c - lo) * scale).astype(np.int64), (1 << bits) - 1)
        key = hilbert_index(g[:, 0], g[:, 1], bits)
    order = np.argsort(key, kind="stable")
    return order_to_perm(order)


def rcm_perm(face_cell_idx: np.ndarray, n_cells: int) -> np.ndarray:
    """Reverse Cuthill-McKee on the cell adjacency graph (cells sharing an interior face)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    fc = np.asarray(face_cell_idx).reshape(-1, 2)
    m = (fc[:, 0] >= 0) & (fc[:, 1] >= 0)
    a, b = fc[m, 0], fc[m, 1]
    g = coo_matrix((np.ones(2 * a.size, dtype=np.int8), (np.concatenate([a, b]), np.concatenate([b, a]))),
                   shape=(n_cells, n_cells)).tocsr()
    order = reverse_cuthill_mckee(g, symmetric_mode=True)
    return order_to_perm(np.asarray(order, dtype=np.int64))


def choose_perm(reorder, n_cells: int, centers=None, face_cell_idx=None, small: int = 65536):
    """Policy behind Environment(reorder=...): 'auto' | 'hilbert' | 'rcm' | 'none' | explicit array."""
    if isinstance(reorder, np.ndarray):
        return np.ascontiguousarray(reorder, dtype=np.int32)
    if reorder in (None, "none", False):
        return None
    if reorder == "hilbert":
        if centers is None:
            raise ValueError("hilbert reordering needs cells.centers")
        return hilbert_perm(centers)
    if reorder == "rcm":
        return rcm_perm(face_cell_idx, n_cells)
    if reorder == "auto":
        if n_cells <= small:
            return None                      # whole state fits in L2/L1: order is irrelevant
        if centers is not None:
            return hilbert_perm(centers)
        return rcm_perm(face_cell_idx, n_cells)
    raise ValueError(f"unknown reorder policy {reorder!r}")
