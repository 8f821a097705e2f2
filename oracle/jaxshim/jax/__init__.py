"""NumPy-backed stand-in for the slice of the `jax` API that FVDBM-JAX touches.

TEST INFRASTRUCTURE ONLY (part of oracle/): lets the *unmodified* reference
sources under /root/reference/src execute on a box without jax so that their
outputs can be recorded as golden vectors (oracle/make_golden.py).  Nothing in
the product package imports this.

Surface covered: jax.numpy.*, jax.vmap(in_axes),
jax.lax.select, jax.jit, jax.tree_util.register_pytree_node_class,
jax.typing.ArrayLike, jax.Array, ndarray.at[idx].set(v).

Float width: JAXSHIM_FLOAT=64 (default) makes *both* jnp.float32 and
jnp.float64 mean float64 -> a clean fp64 oracle.  JAXSHIM_FLOAT=32 mimics stock
JAX (x64 disabled): every float is float32, ints are int32 and int*float
promotes to float32.
"""
import numpy as _np
from . import numpy            # noqa: F401  (jax.numpy)
from . import lax              # noqa: F401
from . import tree_util        # noqa: F401
from . import typing           # noqa: F401
from .numpy import ShimArray as Array

__shim__ = True


def jit(fun=None, **_kw):
    """Identity: the reference only uses @jax.jit as a decorator."""
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    """Python-loop vmap: slice axis 0 of every mapped argument, stack results."""
    def mapped(*args):
        axes = in_axes
        if isinstance(axes, int) or axes is None:
            axes = (axes,) * len(args)
        if len(axes) != len(args):
            raise ValueError("vmap in_axes length mismatch")
        n = None
        for a, ax in zip(args, axes):
            if ax is None:
                continue
            if ax != 0:
                raise NotImplementedError("shim vmap maps axis 0 only")
            m = numpy.asarray(a).shape[0]
            if n is None:
                n = m
            elif n != m:
                raise ValueError("vmap size mismatch")
        if n is None:
            raise ValueError("vmap needs at least one mapped argument")
        outs = []
        for i in range(n):
            call = [a if ax is None else numpy.asarray(a)[i] for a, ax in zip(args, axes)]
            outs.append(fun(*call))
        if n == 0:
            raise ValueError("vmap over empty axis not supported by the shim")
        if isinstance(outs[0], tuple):
            return tuple(numpy.stack([o[j] for o in outs]) for j in range(len(outs[0])))
        return numpy.stack(outs)
    return mapped


class _Tree:
    @staticmethod
    def map(f, *trees):  # dead code in the reference; kept for import-compat
        raise NotImplementedError

tree = _Tree()
