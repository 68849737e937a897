"""jax.numpy stand-in on torch (CPU, float64 / int64).  Only what updes/{utils,cloud,assembly,operators}.py call."""
import math

import numpy as _np
import torch

ndarray = torch.Tensor
newaxis = None
inf = math.inf
pi = math.pi
_F64, _I64 = torch.float64, torch.int64
float64, int64, int32 = torch.float64, torch.int64, torch.int32


def _dtype(dt):
    if dt is None:
        return None
    if dt in (int, _np.int64, _np.int32, "int"):
        return _I64
    if dt in (float, _np.float64, "float"):
        return _F64
    if dt is bool:
        return torch.bool
    return dt


def _has_tensor(v):
    if isinstance(v, torch.Tensor):
        return True
    if isinstance(v, (list, tuple)):
        return any(_has_tensor(e) for e in v)
    return False


def array(v, dtype=None):
    dt = _dtype(dtype)
    if isinstance(v, torch.Tensor):
        return v if dt is None else v.to(dt)
    if isinstance(v, (list, tuple)) and _has_tensor(v):
        t = torch.stack([array(e) for e in v])
        return t if dt is None else t.to(dt)
    a = _np.asarray(v)
    if dt is None:
        dt = _F64 if a.dtype.kind == "f" else (_I64 if a.dtype.kind in "iu" else None)
    return torch.as_tensor(a, dtype=dt)


asarray = array


def zeros(shape, dtype=None):
    return torch.zeros(shape, dtype=_dtype(dtype) or _F64)


def ones(shape, dtype=None):
    return torch.ones(shape, dtype=_dtype(dtype) or _F64)


def _as_int(v):
    if callable(v):                    # `x.size` is an int on a jax / numpy array and a method on a tensor
        v = v()
    if isinstance(v, torch.Size):
        return int(_np.prod(tuple(v), dtype=_np.int64))
    return int(v)


def arange(*a, dtype=None):
    return torch.arange(*[_as_int(v) for v in a], dtype=_dtype(dtype) or _I64)


def linspace(a, b, n):
    return torch.as_tensor(_np.linspace(a, b, n), dtype=_F64)       # jnp.linspace == np.linspace on these grids


def meshgrid(*xs):
    return torch.meshgrid(*xs, indexing="xy")


def stack(seq, axis=0):
    return torch.stack([array(e) if not isinstance(e, torch.Tensor) else e for e in seq], dim=axis)


def concatenate(seq, axis=0):
    return torch.cat([array(e) if not isinstance(e, torch.Tensor) else e for e in seq], dim=axis)


def nan_to_num(x, nan=0.0, posinf=None, neginf=None):
    return torch.nan_to_num(array(x), nan=nan, posinf=posinf, neginf=neginf)


def dot(a, b):
    a, b = array(a), array(b)
    return torch.matmul(a, b)


def sum(x, axis=None):  # noqa: A001
    x = array(x)
    return torch.sum(x) if axis is None else torch.sum(x, dim=axis)


def mean(x, axis=None):
    x = array(x)
    return torch.mean(x) if axis is None else torch.mean(x, dim=axis)


def max(x, axis=None):  # noqa: A001
    x = array(x)
    return torch.max(x) if axis is None else torch.max(x, dim=axis).values


def min(x, axis=None):  # noqa: A001
    x = array(x)
    return torch.min(x) if axis is None else torch.min(x, dim=axis).values


def sqrt(x):
    return torch.sqrt(array(x))


def log(x):
    return torch.log(array(x))


def exp(x):
    return torch.exp(array(x))


def sin(x):
    return torch.sin(array(x))


def cos(x):
    return torch.cos(array(x))


def cosh(x):
    return torch.cosh(array(x))


def abs(x):  # noqa: A001
    return torch.abs(array(x))


def allclose(a, b, rtol=1e-05, atol=1e-08):
    a, b = array(a), array(b)
    return bool(torch.allclose(a.to(_F64), b.to(_F64).expand_as(a) if b.dim() == 0 else b.to(_F64), rtol=rtol, atol=atol))


def where(cond, a, b):
    cond = array(cond)
    like = a if isinstance(a, torch.Tensor) else (b if isinstance(b, torch.Tensor) else None)
    dt = like.dtype if like is not None else _F64
    return torch.where(cond, torch.as_tensor(a, dtype=dt), torch.as_tensor(b, dtype=dt))


def trace(x):
    return torch.trace(array(x))


def clip(x, lo, hi):
    return torch.clamp(array(x), lo, hi)


def flip(x, axis=0):
    return torch.flip(array(x), dims=(axis,))


def set_printoptions(**kw):
    pass


class linalg:
    @staticmethod
    def norm(x, axis=None):
        x = array(x)                       # jnp functions accept numpy arrays too
        return torch.linalg.norm(x) if axis is None else torch.linalg.norm(x, dim=axis)

    @staticmethod
    def inv(a):
        return torch.linalg.inv(array(a))   # LAPACK getrf + getri, as jaxlib's CPU path

    @staticmethod
    def solve(a, b):
        return torch.linalg.solve(array(a), array(b))


class _At:
    """x.at[idx].set(v) / .add(v): functional updates (a new array; x is left untouched)."""

    def __init__(self, x):
        self.x = x

    def __getitem__(self, idx):
        return _AtIdx(self.x, idx)


class _AtIdx:
    def __init__(self, x, idx):
        self.x, self.idx = x, idx

    def _whole(self):
        return self.idx is Ellipsis or (isinstance(self.idx, slice) and self.idx == slice(None))

    def set(self, v):
        if self._whole():                      # out of place: also valid when v is a vmapped value and x is not
            return torch.zeros_like(self.x) + v
        y = self.x.clone()
        y[self.idx] = v if not isinstance(v, torch.Tensor) else v.to(y.dtype)
        return y

    def add(self, v):
        if self._whole():
            return self.x + v
        y = self.x.clone()
        y[self.idx] = y[self.idx] + v
        return y


torch.Tensor.at = property(lambda self: _At(self))
torch.Tensor.copy = lambda self: self.clone()          # jax arrays have .copy()


def _numpy_style(name):
    """np.mean(x) / np.sum(x) ... on an array dispatch to x.mean(axis=None, dtype=None, out=None): accept that
    calling convention on tensors, as jax arrays do."""
    orig = getattr(torch.Tensor, name)

    def method(self, *args, axis=None, dtype=None, out=None, keepdims=False, **kw):
        if axis is not None:
            kw["dim"] = axis
            if keepdims:
                kw["keepdim"] = True
        return orig(self, *args, **kw)
    setattr(torch.Tensor, name, method)


for _n in ("mean", "sum"):
    _numpy_style(_n)
