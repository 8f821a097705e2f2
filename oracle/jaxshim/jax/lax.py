"""jax.lax stand-in: only `select` is used by the reference (containers.py)."""
import numpy as _np
from .numpy import _wrap


def select(pred, on_true, on_false):
    return _wrap(_np.where(_np.asarray(pred), _np.asarray(on_true), _np.asarray(on_false)))
