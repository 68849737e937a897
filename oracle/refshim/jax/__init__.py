"""Stand-in for the part of the `jax` API the Updes hot path uses, on torch.func (CPU, float64).  Test infrastructure:
see oracle/refshim/README.md.  Semantics follow JAX: grad / jacfwd differentiate w.r.t. argument 0, vmap takes
in_axes / out_axes, jit is the identity (no tracing compiler here), non-array outputs are promoted to arrays."""
import functools

import torch
from torch import func as _tf

from . import numpy  # noqa: F401  (jax.numpy)
from . import tree_util, lax, random  # noqa: F401
from . import ffi  # noqa: F401  (custom-call surface used by integration/updes_jax.py)

_F64 = torch.float64


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


class ShapeDtypeStruct:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(int(s) for s in shape), dtype


def jit(fun=None, **kwargs):
    """No tracing: the function itself (static_argnums / static_argnames are accepted and ignored)."""
    if fun is None:
        return lambda f: f
    return fun


def _as_out(v, like=None):
    if isinstance(v, torch.Tensor):
        return v
    if isinstance(v, (tuple, list)):
        return type(v)(_as_out(e, like) for e in v)
    t = torch.as_tensor(v, dtype=_F64)
    if like is not None:                      # keep a (zero) dependence on the input so that grad() returns zeros
        t = t + 0.0 * like.sum()
    return t


def grad(fun, argnums=0):
    assert argnums == 0

    def scalar(*args, **kw):
        return _as_out(fun(*args, **kw), like=args[0])
    return _tf.grad(scalar)


def jacfwd(fun, argnums=0):
    assert argnums == 0

    def f(*args, **kw):
        return _as_out(fun(*args, **kw), like=args[0])
    return _tf.jacfwd(f)


def jacrev(fun, argnums=0):
    assert argnums == 0
    return _tf.jacrev(lambda *a, **k: _as_out(fun(*a, **k), like=a[0]))


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        dims = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        dims = tuple(dims)
        assert len(dims) == len(args), "vmap: in_axes must match the positional arguments"
        args = tuple(numpy.asarray(a) if (d is not None and not isinstance(a, torch.Tensor)) else a for a, d in zip(args, dims))
        return _tf.vmap(lambda *a: _as_out(fun(*a)), in_dims=dims, out_dims=out_axes)(*args)
    return functools.wraps(fun)(mapped) if hasattr(fun, "__name__") else mapped
