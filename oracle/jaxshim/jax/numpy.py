"""jax.numpy stand-in on top of NumPy (see package docstring)."""
import os as _os
import numpy as _np

_FLOAT = _np.float32 if _os.environ.get("JAXSHIM_FLOAT", "64") == "32" else _np.float64

float32 = _FLOAT
float64 = _FLOAT
int32 = _np.int32
newaxis = None
pi = _np.pi
inf = _np.inf


class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self._arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def set(self, value):
        out = _np.array(self._arr, copy=True)
        idx = self._idx
        if isinstance(idx, tuple):
            idx = tuple(_np.asarray(i) if isinstance(i, _np.ndarray) else i for i in idx)
        elif isinstance(idx, _np.ndarray):
            idx = _np.asarray(idx)
        out[idx] = _np.asarray(value)
        return _wrap(out)


class ShimArray(_np.ndarray):
    """ndarray with jax's functional-update `.at[...]` and jax-like promotion."""

    @property
    def at(self):
        return _At(self)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        ins = []
        fdt = None
        for x in inputs:
            if isinstance(x, _np.ndarray) and x.dtype.kind == "f":
                fdt = _FLOAT
        for x in inputs:
            if isinstance(x, ShimArray):
                x = x.view(_np.ndarray)
            if fdt is not None and isinstance(x, _np.ndarray) and x.dtype.kind in "iub" and ufunc.nin > 1 \
                    and ufunc not in (_np.equal, _np.not_equal, _np.less, _np.greater, _np.less_equal, _np.greater_equal):
                x = x.astype(fdt)          # jax: int32 (op) float32 -> float32
            ins.append(x)
        if out is not None:
            kwargs["out"] = tuple(o.view(_np.ndarray) if isinstance(o, ShimArray) else o for o in out)
        res = getattr(ufunc, method)(*ins, **kwargs)
        return _canon(res)


def _canon(res):
    if isinstance(res, tuple):
        return tuple(_canon(r) for r in res)
    if isinstance(res, _np.ndarray) or _np.isscalar(res) or isinstance(res, _np.generic):
        a = _np.asarray(res)
        if a.dtype.kind == "f" and a.dtype != _FLOAT:
            a = a.astype(_FLOAT)
        elif a.dtype.kind == "i" and a.dtype != _np.int32:
            a = a.astype(_np.int32)
        return a.view(ShimArray)
    return res


def _wrap(a):
    return _canon(_np.asarray(a))


def _dt(dtype):
    if dtype is None:
        return None
    d = _np.dtype(dtype)
    if d.kind == "f":
        return _np.dtype(_FLOAT)
    if d.kind == "i":
        return _np.dtype(_np.int32)
    return d


def asarray(x, dtype=None):
    if hasattr(x, "__jax_array__") and not isinstance(x, _np.ndarray):
        x = x.__jax_array__()
    a = _np.asarray(x)
    if dtype is not None:
        a = a.astype(_dt(dtype))
    return _wrap(a)


def array(x, dtype=None):
    return asarray(_np.array(asarray(x), copy=True), dtype)


def zeros(shape, dtype=None):
    return _wrap(_np.zeros(shape, dtype=_dt(dtype) or _FLOAT))


def ones(shape, dtype=None):
    return _wrap(_np.ones(shape, dtype=_dt(dtype) or _FLOAT))


def zeros_like(x, dtype=None):
    return _wrap(_np.zeros_like(_np.asarray(x), dtype=_dt(dtype)))


def ones_like(x, dtype=None):
    return _wrap(_np.ones_like(_np.asarray(x), dtype=_dt(dtype)))


def full_like(x, fill_value, dtype=None):
    return _wrap(_np.full_like(_np.asarray(x), fill_value, dtype=_dt(dtype)))


def arange(*a, **k):
    return _wrap(_np.arange(*a, **k))


def linspace(*a, **k):
    return _wrap(_np.linspace(*a, **k))


def repeat(x, n, axis=None):
    return _wrap(_np.repeat(_np.asarray(x), n, axis=axis))


def _fl(x):
    a = asarray(x)
    if a.dtype.kind in "iub":
        a = asarray(a, _FLOAT)
    return a


def sqrt(x):
    return _wrap(_np.sqrt(_fl(x).view(_np.ndarray)))


def cos(x):
    return _wrap(_np.cos(_fl(x).view(_np.ndarray)))


def sin(x):
    return _wrap(_np.sin(_fl(x).view(_np.ndarray)))


def abs(x):  # noqa: A001
    return _wrap(_np.abs(_np.asarray(x)))


def dot(a, b):
    a, b = asarray(a), asarray(b)
    if a.dtype.kind == "f" or b.dtype.kind == "f":
        a, b = a.astype(_FLOAT), b.astype(_FLOAT)
    return _wrap(_np.dot(a.view(_np.ndarray), b.view(_np.ndarray)))


def sum(x, axis=None, keepdims=False):  # noqa: A001
    return _wrap(_np.sum(_np.asarray(x), axis=axis, keepdims=keepdims))


def mean(x, axis=None, keepdims=False):
    return _wrap(_np.mean(_np.asarray(x), axis=axis, keepdims=keepdims))


def max(x, axis=None):  # noqa: A001
    return _wrap(_np.max(_np.asarray(x), axis=axis))


def min(x, axis=None):  # noqa: A001
    return _wrap(_np.min(_np.asarray(x), axis=axis))


def where(cond, *args):
    if not args:
        return tuple(_wrap(i) for i in _np.where(_np.asarray(cond)))
    a, b = args
    a_is_py = isinstance(a, (int, float)) and not isinstance(a, bool)
    b_is_py = isinstance(b, (int, float)) and not isinstance(b, bool)
    a_, b_ = _np.asarray(a), _np.asarray(b)
    if a_is_py and not b_is_py:       # weak python scalars adopt the array dtype
        a_ = a_.astype(b_.dtype)
    if b_is_py and not a_is_py:
        b_ = b_.astype(a_.dtype)
    return _wrap(_np.where(_np.asarray(cond), a_, b_))


def argwhere(x):
    return _wrap(_np.argwhere(_np.asarray(x)))


def unique(x):
    return _wrap(_np.unique(_np.asarray(x)))


def concatenate(xs, axis=0):
    return _wrap(_np.concatenate([_np.asarray(x) for x in xs], axis=axis))


def stack(xs, axis=0):
    return _wrap(_np.stack([_np.asarray(x) for x in xs], axis=axis))


def hstack(xs):
    return _wrap(_np.hstack([_np.asarray(x) for x in xs]))


def pad(x, pad_width, mode="constant", constant_values=0):
    return _wrap(_np.pad(_np.asarray(x), pad_width, mode=mode, constant_values=constant_values))


def reshape(x, shape):
    return _wrap(_np.reshape(_np.asarray(x), shape))


def expand_dims(x, axis):
    return _wrap(_np.expand_dims(_np.asarray(x), axis))


def meshgrid(*a, **k):
    return tuple(_wrap(m) for m in _np.meshgrid(*a, **k))


ndarray = ShimArray
