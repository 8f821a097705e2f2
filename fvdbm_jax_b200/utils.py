"""NumPy counterparts of the small helpers the reference's modules star-export from ``utils/utils.py``
(/root/reference/utils/utils.py) -- the hot path itself uses none of them (its arithmetic lives in the CUDA
kernels); they exist so that notebook code calling them keeps working after the import swap
(``fvdbm_jax_b200.compat``).  ``CustomArray`` lives in ``containers``."""
from __future__ import annotations

import numpy as np

from .containers import CustomArray  # noqa: F401  (utils/utils.py:176-230)

__all__ = ["pad_stack", "weighted_avg", "extrapolate", "extrap_pdf", "interp_pdf", "calc_dist", "calc_normal", "CustomArray"]


def pad_stack(args):
    """Stack 1-D arrays of different lengths into one (n, longest) array padded with -1 (utils.py:9-20)."""
    args = [np.asarray(a).reshape(-1) for a in args]
    width = max((a.size for a in args), default=0)
    out = -np.ones((len(args), width), dtype=np.result_type(*args) if args else np.float64)
    for i, a in enumerate(args):
        out[i, :a.size] = a
    return out


def weighted_avg(x, w):
    """sum_j x_j w_j / sum_j w_j over axis 0 (utils.py:34-45)."""
    x, w = np.asarray(x), np.asarray(w)
    return np.sum(x * w[..., np.newaxis], axis=0) / np.sum(w[..., np.newaxis], axis=0)


def extrapolate(x, d):
    """Inverse-distance weighting with negative (padding) distances switched off (utils.py:47-60)."""
    with np.errstate(divide="ignore"):
        w = 1.0 / np.asarray(d, dtype=np.float64)
    return weighted_avg(x, np.where(w < 0, 0, w))


def extrap_pdf(pdf1, pdf2, extrap_dist, pdf2_dist):
    """Linear extrapolation through pdf1 away from pdf2 (utils.py:153-154): the ghost-cell rule."""
    return pdf1 + (pdf1 - pdf2) * (extrap_dist / pdf2_dist)


def interp_pdf(pdf, dist):
    return weighted_avg(pdf, 1.0 / np.asarray(dist, dtype=np.float64))


def calc_dist(p1, p2):
    d = np.asarray(p2, dtype=np.float64) - np.asarray(p1, dtype=np.float64)
    return float(np.sqrt(np.sum(d * d)))


def calc_normal(p1, p2):
    """Unit normal on the left of p1 -> p2 (utils.py:162-173)."""
    t = np.asarray(p2, dtype=np.float64) - np.asarray(p1, dtype=np.float64)
    n = np.array([-t[1], t[0]])
    return n / np.sqrt(n[0] * n[0] + n[1] * n[1])
