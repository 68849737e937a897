"""Drop-in call surface of the hot path (reference ``updes/operators.py``).

* term set ``nodal_value / nodal_gradient / nodal_laplacian / nodal_div_grad`` (operators.py:15-111)
* field evaluators ``value / gradient / laplacian / divergence`` (+ ``_vec``) (operators.py:118-351)
* BC preparation (operators.py:512-555, :628-647)
* ``pde_solver``, ``pde_solver_jit``, ``pde_solver_jit_with_bc`` (operators.py:559-683), ``SteadySol``

The reference traces the user's ``diff_operator`` with JAX and differentiates the kernel per matrix
entry.  Here the operator is *lowered*: it is called once with symbolic jets, and must come back as
a linear combination  a0 phi + a1 phi_x + a2 phi_y + a3 phi_xx + a4 phi_yy  whose coefficients may
depend on the row (through ``x`` and ``fields``).  The five coefficient columns drive the CUDA
assembly kernel.  Anything that is not linear in the built-in term set raises ``OperatorLoweringError``.
"""
from __future__ import annotations

import numbers
import warnings
from collections import OrderedDict

import numpy as np

from . import _lib
from . import assembly as _asm
from .linalg import LUFactorization
from .rbf import compute_nb_monomials, identify_rbf


class OperatorLoweringError(TypeError):
    """The user's differential operator is outside the supported (linear, built-in term) set."""


# ==================================================================================================
# Symbolic jets
# ==================================================================================================
def _is_coeff(v):
    return isinstance(v, (numbers.Real, np.ndarray, np.generic))


class Jet:
    """Linear combination of (phi, phi_x, phi_y, phi_xx, phi_yy) with per-row coefficients."""
    __array_ufunc__ = None          # make ndarray <op> Jet defer to the reflected methods below
    __array_priority__ = 1000

    def __init__(self, coef):
        self.coef = list(coef)

    def _lin(self, other, sign):
        if isinstance(other, Jet):
            return Jet([a + sign * b for a, b in zip(self.coef, other.coef)])
        if _is_coeff(other) and np.all(np.asarray(other) == 0):
            return Jet(self.coef)
        raise OperatorLoweringError(
            "the differential operator must be linear in the nodal terms: cannot add %r to a nodal term "
            "(affine parts belong in the rhs operator)" % (other,))

    def __add__(self, o): return self._lin(o, 1.0)
    __radd__ = __add__
    def __sub__(self, o): return self._lin(o, -1.0)
    def __rsub__(self, o): return (-self)._lin(o, 1.0)
    def __neg__(self): return Jet([-a for a in self.coef])
    def __pos__(self): return self

    def __mul__(self, o):
        if isinstance(o, (Jet, JetVector)):
            raise OperatorLoweringError(
                "product of two nodal terms: the operator is non-linear in the basis function and cannot be "
                "assembled (only nodal_value / nodal_gradient / nodal_laplacian / nodal_div_grad, combined "
                "linearly with coefficients that depend on x and fields, are supported)")
        if not _is_coeff(o):
            raise OperatorLoweringError("cannot scale a nodal term by %r" % (o,))
        return Jet([a * o for a in self.coef])
    __rmul__ = __mul__

    def __truediv__(self, o):
        if not _is_coeff(o):
            raise OperatorLoweringError("cannot divide a nodal term by %r" % (o,))
        return Jet([a / o for a in self.coef])

    def __rtruediv__(self, o):
        raise OperatorLoweringError("division by a nodal term is non-linear in the basis function")

    def __pow__(self, o):
        raise OperatorLoweringError("powers of a nodal term are non-linear in the basis function")

    def __float__(self):
        raise OperatorLoweringError("nodal terms are symbolic during assembly and have no numeric value")

    def __array__(self, *a, **k):
        raise OperatorLoweringError(
            "a nodal term was passed to a numpy function; only + - * / with coefficients, indexing of "
            "nodal_gradient and dot products are supported")


class JetVector:
    """Gradient-like pair of jets; supports indexing, iteration and dot products with 2-vectors."""
    __array_ufunc__ = None
    __array_priority__ = 1000

    def __init__(self, comps):
        self.comps = list(comps)

    def __getitem__(self, i): return self.comps[i]
    def __len__(self): return len(self.comps)
    def __iter__(self): return iter(self.comps)
    def __neg__(self): return JetVector([-c for c in self.comps])
    def __mul__(self, o):
        if isinstance(o, (Jet, JetVector)):
            raise OperatorLoweringError("product of two nodal terms is non-linear in the basis function")
        o = np.asarray(o)
        if o.ndim >= 1 and o.shape[0] == len(self.comps):
            return JetVector([c * o[i] for i, c in enumerate(self.comps)])
        return JetVector([c * o for c in self.comps])
    __rmul__ = __mul__
    def __truediv__(self, o): return JetVector([c / o for c in self.comps])
    def __add__(self, o):
        if isinstance(o, JetVector):
            return JetVector([a + b for a, b in zip(self.comps, o.comps)])
        raise OperatorLoweringError("cannot add %r to a nodal gradient" % (o,))
    def __sub__(self, o): return self + (-o)

    def dot(self, v):
        if isinstance(v, (Jet, JetVector)):
            raise OperatorLoweringError("dot product of two nodal terms is non-linear in the basis function")
        v = list(v) if not isinstance(v, np.ndarray) else v
        if len(v) != len(self.comps):
            raise OperatorLoweringError("dot product with a vector of length %d (expected %d)" % (len(v), len(self.comps)))
        out = self.comps[0] * v[0]
        for i in range(1, len(self.comps)):
            out = out + self.comps[i] * v[i]
        return out

    def __matmul__(self, v): return self.dot(v)
    def __rmatmul__(self, v): return self.dot(v)

    def __array_function__(self, func, types, args, kwargs):
        if func in (np.dot, np.inner, np.vdot, np.matmul) and len(args) == 2:
            a, b = args
            return a.dot(b) if isinstance(a, JetVector) else b.dot(a)
        if func is np.sum and len(args) == 1:
            out = self.comps[0]
            for c in self.comps[1:]:
                out = out + c
            return out
        raise OperatorLoweringError("numpy function %s is not supported on a nodal gradient" % getattr(func, "__name__", func))


def dot(a, b):
    """Dot product usable inside operators (stands in for ``jnp.dot``)."""
    if isinstance(a, JetVector):
        return a.dot(b)
    if isinstance(b, JetVector):
        return b.dot(a)
    return np.dot(a, b)


class _Basis:
    """Sentinel standing for 'the RBF centre' / 'the monomial' while an operator is being lowered."""
    def __init__(self, name): self.name = name
    def __repr__(self): return "<symbolic %s>" % self.name
    def __call__(self, *a, **k):
        raise OperatorLoweringError("the %s is symbolic during assembly; use the nodal_* term set" % self.name)


_CENTER = _Basis("rbf centre")
_MONOMIAL = _Basis("monomial")


def _check_symbolic(center, monomial, what):
    if center is _CENTER or monomial is _MONOMIAL:
        return
    raise OperatorLoweringError(
        "%s is only available inside a diff_operator passed to pde_solver (symbolic lowering); for numeric "
        "field values use value / gradient / laplacian" % what)


def nodal_value(x, center=None, rbf=None, monomial=None):
    """rbf or monomial value at x (operators.py:15-32)."""
    _check_symbolic(center, monomial, "nodal_value")
    return Jet([1.0, 0.0, 0.0, 0.0, 0.0])


def nodal_gradient(x, center=None, rbf=None, monomial=None):
    """gradient w.r.t. x, NaN/inf at r = 0 replaced by 0 (operators.py:42-60)."""
    _check_symbolic(center, monomial, "nodal_gradient")
    return JetVector([Jet([0.0, 1.0, 0.0, 0.0, 0.0]), Jet([0.0, 0.0, 1.0, 0.0, 0.0])])


def nodal_laplacian(x, center=None, rbf=None, monomial=None):
    """trace of the Hessian w.r.t. x (operators.py:70-85)."""
    _check_symbolic(center, monomial, "nodal_laplacian")
    return Jet([0.0, 0.0, 0.0, 1.0, 1.0])


def nodal_div_grad(x, center=None, rbf=None, monomial=None, args=None):
    """args[0] phi_xx + args[1] phi_yy (operators.py:88-111)."""
    _check_symbolic(center, monomial, "nodal_div_grad")
    if args is None or len(args) != 2:
        raise OperatorLoweringError("nodal_div_grad needs args=(a, b)")
    return Jet([0.0, 0.0, 0.0, args[0], args[1]])


class BatchPoints(np.ndarray):
    """Coordinates of all rows at once, shape (2, rows): ``x[0]`` / ``x[1]`` are per-row vectors, so
    an operator written for one point ``x`` evaluates for every row in a single call."""
    def __new__(cls, arr):
        return np.asarray(arr, dtype=np.float64).view(cls)

    def __getitem__(self, idx):
        return np.asarray(super().__getitem__(idx))


def lower_diff_operator(diff_operator, cloud, rbf, diff_args=None):
    """Call the user's operator once per basis family with symbolic jets; return the (Ni, 5)
    coefficient tables for RBF columns and monomial columns (reference assembly.py:93-137)."""
    Ni, N = cloud.Ni, cloud.N
    x = BatchPoints(cloud.sorted_nodes[:Ni].T)
    if diff_args:
        F = np.stack([np.asarray(a, dtype=np.float64) for a in diff_args], axis=-1)    # (N, ..., nf)
        if F.shape[0] != N:
            raise ValueError("diff_args fields must have one value per node (N = %d)" % N)
        fields = np.moveaxis(F[:Ni], 0, -1)                                            # (..., nf, Ni)
    else:
        fields = np.ones((1, Ni))                                                      # assembly.py:114

    def run(center, monomial):
        out = diff_operator(x, center, rbf, monomial, fields)
        if isinstance(out, JetVector):
            raise OperatorLoweringError("the differential operator must return a scalar, got a nodal gradient")
        if not isinstance(out, Jet):
            raise OperatorLoweringError(
                "the differential operator returned %r, which does not involve the basis function; it must be a "
                "linear combination of nodal_value / nodal_gradient / nodal_laplacian / nodal_div_grad" % (out,))
        tab = np.empty((Ni, 5))
        for k in range(5):
            c = np.asarray(out.coef[k], dtype=np.float64)
            if c.ndim > 1 or (c.ndim == 1 and c.shape[0] != Ni):
                raise OperatorLoweringError("operator coefficient %d has shape %s; expected a scalar or one value per row" % (k, c.shape))
            tab[:, k] = c
        if not np.all(np.isfinite(tab)):
            raise OperatorLoweringError("operator coefficients are not finite")
        return tab

    return run(_CENTER, None), run(None, _MONOMIAL)


# ==================================================================================================
# Field evaluators (matrix-free, GPU)
# ==================================================================================================
def _points(x):
    """-> (pts (R,2) ndarray, layout) with layout in {'single', 'batch', 'rows'}."""
    if isinstance(x, BatchPoints):
        return np.ascontiguousarray(np.asarray(x).T), "batch"
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        return x.reshape(1, 2), "single"
    return np.ascontiguousarray(x.reshape(-1, 2)), "rows"


def _jets(x, field, centers, rbf):
    torch = _lib.require_cuda()
    kind, param = identify_rbf(rbf)
    pts, layout = _points(x)
    field = np.asarray(field, dtype=np.float64)
    coeffs = field.reshape(field.shape[0], -1).T                   # (nf, N+M)
    dev = "cuda"
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(dev)
    jphi, jpol = _asm.eval_jets(kind, param, t(centers), t(coeffs), t(pts))
    return (jphi + jpol).cpu().numpy(), layout                      # (nf, R, 5)


def _clip(v, clip_val):
    return np.clip(v, -clip_val, clip_val) if clip_val else v


def value(x, field, centers, rbf=None, clip_val=None):
    """Field value at x from its coefficients (operators.py:118-147); all centres, self included."""
    J, layout = _jets(x, field, centers, rbf)
    v = J[0, :, 0]
    return _clip(float(v[0]) if layout == "single" else v, clip_val)


def gradient(x, field, centers, rbf=None, clip_val=None):
    """Field gradient at x (operators.py:156-184).  Shape (2,), (2, rows) inside operators, (R, 2) for rows."""
    J, layout = _jets(x, field, centers, rbf)
    g = J[0, :, 1:3]
    if layout == "single":
        g = g[0]
    elif layout == "batch":
        g = g.T
    return _clip(g, clip_val)


def laplacian(x, field, centers, rbf=None, clip_val=None):
    """Field Laplacian at x (operators.py:294-330)."""
    J, layout = _jets(x, field, centers, rbf)
    v = J[0, :, 3] + J[0, :, 4]
    return _clip(float(v[0]) if layout == "single" else v, clip_val)


def divergence(x, field, centers, rbf=None, clip_val=None):
    """Divergence of a vector field given as (N+M, 2) coefficients (operators.py:263-284)."""
    J, layout = _jets(x, field, centers, rbf)
    v = J[0, :, 1] + J[1, :, 2]
    return _clip(float(v[0]) if layout == "single" else v, clip_val)


value_vec = value
gradient_vec = gradient
laplacian_vec = laplacian
divergence_vec = divergence


# ==================================================================================================
# Boundary-condition preparation (host)
# ==================================================================================================
def _eval_on_nodes(fn, nodes):
    return np.array([float(fn(nodes[k])) for k in range(nodes.shape[0])], dtype=np.float64)


def duplicate_robin_coeffs(boundary_conditions, cloud):
    """operators.py:512-541: per-node beta for Robin facets; strips the (value, beta) tuples."""
    robin_coeffs, new_bc = {}, {}
    for f_id, f_type in cloud.facet_types.items():
        if f_type == "r":
            node_ids = cloud.facet_nodes[f_id]
            bc = boundary_conditions[f_id]
            if type(bc) == tuple:
                new_bc[f_id] = bc[0]
                betas = bc[1]
                if callable(betas):
                    betas = _eval_on_nodes(betas, cloud.sorted_nodes[np.asarray(node_ids)])
                betas = np.broadcast_to(np.asarray(betas, dtype=np.float64), (len(node_ids),))
            else:
                # the reference calls warning.warn on an un-imported name here and dies with
                # AttributeError (operators.py:1,:530); keep it an explicit error
                raise ValueError("Robin facet %r needs a (value, beta) tuple" % f_id)
            for i in node_ids:
                robin_coeffs[i] = betas[i - node_ids[0]]          # operators.py:535-536 (contiguous ids)
        else:
            new_bc[f_id] = boundary_conditions[f_id]
    return robin_coeffs, new_bc


def zerofy_periodic_cond(boundary_conditions, cloud):
    """operators.py:546-555"""
    for f_id, f_type in cloud.facet_types.items():
        if f_type[0] == "p":
            boundary_conditions[f_id] = np.zeros(len(cloud.facet_nodes[f_id]))
    return boundary_conditions


def boundary_conditions_func_to_arr(boundary_conditions, cloud):
    """operators.py:628-647: callables -> arrays over the facet's nodes."""
    out = {}
    for f_id, f_bc in boundary_conditions.items():
        nodes = cloud.sorted_nodes[np.asarray(cloud.facet_nodes[f_id], dtype=int)]
        if callable(f_bc):
            out[f_id] = _eval_on_nodes(f_bc, nodes)
        elif type(f_bc) == tuple:
            v, b = f_bc
            if callable(v):
                v = _eval_on_nodes(v, nodes)
            if callable(b):
                b = _eval_on_nodes(b, nodes)
            out[f_id] = (v, b)
        else:
            out[f_id] = f_bc
    return out


# ==================================================================================================
# Solver
# ==================================================================================================
class SteadySol:
    """(vals, coeffs, mat) -- reference ``SteadySol`` namedtuple (utils.py:148).  ``mat`` (the
    reference's B = diffMat inv(A)[:, :N]) needs inv(A) and is only computed when read."""
    _fields = ("vals", "coeffs", "mat")

    def __init__(self, vals, coeffs, mat_fn=None):
        self.vals, self.coeffs, self._mat_fn, self._mat = vals, coeffs, mat_fn, None

    @property
    def mat(self):
        if self._mat is None and self._mat_fn is not None:
            self._mat = self._mat_fn()
        return self._mat

    def __iter__(self):
        return iter((self.vals, self.coeffs, self.mat))

    def __repr__(self):
        return "PDESolution(vals=%r, coeffs=%r, mat=<lazy>)" % (self.vals, self.coeffs)


class _System:
    """An assembled + factored collocation system resident on the GPU."""

    def __init__(self, cloud, kind, param, M, table):
        self.kind, self.param, self.M = kind, param, M
        self.cloud = cloud                   # the cache key uses id(cloud): keep it alive while cached
        self.rows = _asm.DeviceRows(cloud, table)
        self.K = _asm.assemble_system(self.rows, kind, param, M)
        self.n = cloud.N + M
        self.lu = LUFactorization(self.K, self.n).factor()

    def solve(self, rhs, refine=1):
        """rhs: (n,) numpy -> coefficients (n,) numpy.  `refine` steps of iterative refinement with
        a matrix-free residual (no second copy of K)."""
        torch = self.rows.torch
        b = torch.as_tensor(rhs, dtype=torch.float64).to(self.K.device)
        c = self.lu.solve(b.clone())
        for _ in range(refine):
            r = b - _asm.apply_rows(self.rows, self.kind, self.param, self.M, c.view(1, -1))[0]
            c = c + self.lu.solve(r)
        return c

    def nbytes(self):
        return self.K.numel() * 8


def _dist_world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
    except Exception:
        pass
    return 1, 0


def default_block_width(n, world):
    """Column-block width of the multi-GPU layout: 2048 for large systems (the update GEMM reaches
    33 TFLOP/s at k = 2048 vs 29 at k = 512), smaller when the matrix would otherwise have fewer than
    ~8 blocks per rank."""
    nb = 2048
    while nb > 64 and n // (nb * world) < 8:
        nb //= 2
    return nb


class _DistSystem:
    """The same system sharded column-block-cyclically over all ranks of the default process group
    (updes_b200/distributed.py).  Every rank calls pde_solver with identical arguments."""

    def __init__(self, cloud, kind, param, M, table, world, rank):
        from .distributed import ColumnBlockCyclic, CudaBackend, DistributedLU
        self.kind, self.param, self.M = kind, param, M
        self.cloud = cloud                   # the cache key uses id(cloud): keep it alive while cached
        self.rows = _asm.DeviceRows(cloud, table)
        self.n = cloud.N + M
        self.layout = ColumnBlockCyclic(self.n, default_block_width(self.n, world), world)
        self.be = CudaBackend(self.layout, rank)
        self.be.assemble(self.rows, kind, param, M)
        self.dlu = DistributedLU(self.layout, rank, self.be).factor()
        self.lu = self                       # zero_pivot() interface of LUFactorization
        self.K = self.be.local

    def zero_pivot(self):
        import torch.distributed as dist
        t = self.be.info.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item())

    def solve(self, rhs, refine=1):
        torch = self.rows.torch
        b = torch.as_tensor(rhs, dtype=torch.float64).to(self.be.device)
        c = self.dlu.solve(rhs)
        for _ in range(refine):
            r = b - _asm.apply_rows(self.rows, self.kind, self.param, self.M, c.view(1, -1))[0]
            c = c + self.dlu.solve(r)
        return c

    def nbytes(self):
        return self.be.local.numel() * 8


_CACHE: "OrderedDict[tuple, _System]" = OrderedDict()
_CACHE_BYTES = 80 << 30


def clear_cache():
    """Drop every cached factorisation (frees the HBM they hold)."""
    _CACHE.clear()


def _cached_system(key, build):
    sys_ = _CACHE.get(key)
    if sys_ is not None:
        _CACHE.move_to_end(key)
        return sys_
    sys_ = build()
    _CACHE[key] = sys_
    while len(_CACHE) > 1 and sum(s.nbytes() for s in _CACHE.values()) > _CACHE_BYTES:
        _CACHE.popitem(last=False)
    return sys_


def _interp_system(cloud, kind, param, M):
    """Factorisation of A = [[Phi P], [P^T 0]] (assembly.py:62-90), cached per (cloud, rbf, M)."""
    key = ("A", id(cloud), kind, param, M)
    return _cached_system(key, lambda: _System(cloud, kind, param, M, _asm.build_interpolation_rows(cloud)))


def core_compute_coefficients(field, cloud, rbf, nb_monomials):
    """inv(A) @ [field; 0] (assembly.py:404-410) through the cached LU of A."""
    kind, param = identify_rbf(rbf)
    rhs = np.concatenate([np.asarray(field, dtype=np.float64), np.zeros(nb_monomials)])
    return _interp_system(cloud, kind, param, nb_monomials).solve(rhs).cpu().numpy()


def compute_coefficients(field, cloud, rbf, max_degree):
    """assembly.py:413-418"""
    return core_compute_coefficients(field, cloud, rbf, compute_nb_monomials(max_degree, cloud.dim))


get_field_coefficients = compute_coefficients      # assembly.py:423-430


def assemble_q(rhs_operator, boundary_conditions, cloud, rbf, nb_monomials, rhs_args):
    """Right-hand side (assembly.py:434-485): rhs operator on internal nodes, BC arrays on facets."""
    N, Ni = cloud.N, cloud.Ni
    if rhs_args is not None:
        cols = []
        for f in rhs_args:
            f = np.asarray(f, dtype=np.float64)
            cols.append(core_compute_coefficients(f, cloud, rbf, nb_monomials) if f.shape[0] == N else f)
        fields = np.stack(cols, axis=-1)
    else:
        fields = None
    x = BatchPoints(cloud.sorted_nodes[:Ni].T)
    q = np.zeros(N)
    q[:Ni] = np.broadcast_to(np.asarray(rhs_operator(x, cloud.sorted_nodes, rbf, fields), dtype=np.float64), (Ni,))
    for f_id in cloud.facet_types.keys():
        assert f_id in boundary_conditions.keys(), "facets and boundary functions don't match ids"
        bd = boundary_conditions[f_id]
        ids = np.asarray(cloud.facet_nodes[f_id], dtype=int)
        q[ids] = _eval_on_nodes(bd, cloud.sorted_nodes[ids]) if callable(bd) else np.asarray(bd, dtype=np.float64)
    return q


def pde_solver(diff_operator, rhs_operator, cloud, boundary_conditions, rbf, max_degree, diff_args=None,
               rhs_args=None, refine=1):
    """Solve a linear PDE by global RBF collocation (reference operators.py:559-618).

    Same arguments and result as the reference.  The factorisation is cached on (cloud, rbf,
    max_degree, lowered operator, Robin betas): repeated calls with an unchanged left-hand side --
    the time loops of demos/Advection -- only pay for the right-hand side and two triangular sweeps.
    """
    kind, param = identify_rbf(rbf)
    robin_coeffs, boundary_conditions = duplicate_robin_coeffs(dict(boundary_conditions), cloud)
    boundary_conditions = zerofy_periodic_cond(boundary_conditions, cloud)
    M = compute_nb_monomials(max_degree, cloud.dim)
    coef_phi, coef_pol = lower_diff_operator(diff_operator, cloud, rbf, diff_args)
    betas = np.array([robin_coeffs[k] for k in sorted(robin_coeffs)], dtype=np.float64) if robin_coeffs else None

    key = ("K", id(cloud), kind, param, M, coef_phi.tobytes(), coef_pol.tobytes(), None if betas is None else betas.tobytes())
    world, rank = _dist_world()
    table_fn = lambda: _asm.build_operator_rows(cloud, coef_phi, coef_pol, betas)
    if world > 1:
        system = _cached_system(key + (world,), lambda: _DistSystem(cloud, kind, param, M, table_fn(), world, rank))
    else:
        system = _cached_system(key, lambda: _System(cloud, kind, param, M, table_fn()))
    q = assemble_q(rhs_operator, boundary_conditions, cloud, rbf, M, rhs_args)
    coeffs_dev = system.solve(np.concatenate([q, np.zeros(M)]), refine=refine)
    if system.lu.zero_pivot():
        warnings.warn("collocation matrix is exactly singular (zero pivot at column %d)" % system.lu.zero_pivot())
    # vals = [Phi P] c with the reference's zero-diagonal Phi (assembly.py:31-32, :404-410)
    torch = system.rows.torch
    own = torch.arange(cloud.N, dtype=torch.int32, device=coeffs_dev.device)
    jphi, jpol = _asm.eval_jets(kind, param, system.rows.centres, coeffs_dev.view(1, -1), system.rows.centres, own)
    vals = (jphi[0, :, 0] + jpol[0, :, 0]).cpu().numpy()
    coeffs = coeffs_dev.cpu().numpy()

    def mat_fn():
        return _reference_mat(cloud, kind, param, M, system)

    return SteadySol(vals, coeffs, mat_fn)


def _reference_mat(cloud, kind, param, M, system):
    """B = (diffMat @ inv(A))[:, :N] (assembly.py:396-401), for callers that read ``SteadySol.mat``.
    A is symmetric, so B[r, :] = (inv(A) diffMat[r, :]^T)[:N]: N right-hand sides against the LU of A."""
    torch = system.rows.torch
    N = cloud.N
    if N > 20000:
        raise MemoryError("SteadySol.mat (the reference's B = diffMat inv(A)[:, :N]) needs N = %d right-hand sides "
                          "against inv(A) and a second N x N matrix; it is only provided for N <= 20000" % N)
    if isinstance(system, _DistSystem):
        raise NotImplementedError("SteadySol.mat is not available on the multi-GPU path")
    D = _asm.assemble_system(system.rows, kind, param, M)[:N].contiguous()     # fresh copy: K holds LU now
    X = _interp_system(cloud, kind, param, M).lu.solve(D)
    return X[:, :N].cpu().numpy()


def pde_solver_jit_with_bc(diff_operator, rhs_operator, cloud, boundary_conditions, rbf, max_degree, diff_args=None,
                           rhs_args=None):
    """operators.py:621-625.  There is no tracing compiler here: identical to ``pde_solver``."""
    return pde_solver(diff_operator, rhs_operator, cloud, boundary_conditions, rbf, max_degree, diff_args, rhs_args)


def pde_solver_jit(diff_operator, rhs_operator, cloud, boundary_conditions, rbf, max_degree, diff_args=None,
                   rhs_args=None):
    """operators.py:650-683: boundary callables are turned into arrays first."""
    bc = boundary_conditions_func_to_arr(boundary_conditions, cloud)
    return pde_solver_jit_with_bc(diff_operator, rhs_operator, cloud, bc, rbf, max_degree, diff_args, rhs_args)


def pde_multi_solver(diff_operators, rhs_operators, cloud, boundary_conditions, rbf, max_degree, nb_iters=10, tol=1e-6,
                     diff_args=None, rhs_args=None):
    """Fixed-count Picard iteration over coupled scalar PDEs (reference operators.py:696-771): each sweep
    solves every equation with the latest values of all unknowns as leading ``diff_args``.  Pure host
    loop around ``pde_solver_jit_with_bc``; ``tol`` is accepted and ignored, as in the reference."""
    n = len(diff_operators)
    assert n == len(rhs_operators) == len(boundary_conditions), \
        "The number of differential operators must match the number of right-hand side operators"
    bcs = []
    for bc in boundary_conditions:
        bcs.append(boundary_conditions_func_to_arr(bc, cloud))
    sols_vals = [np.asarray(v, dtype=np.float64) for v in diff_args[0][:n]]
    sols = None
    for _ in range(nb_iters):
        sols = [pde_solver_jit_with_bc(diff_operators[i], rhs_operators[i], cloud, bcs[i], rbf, max_degree,
                                       diff_args=sols_vals + list(diff_args[i][n:]),
                                       rhs_args=None if rhs_args is None else rhs_args[i]) for i in range(n)]
        sols_vals = [s_.vals for s_ in sols]
    return sols
