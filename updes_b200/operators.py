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

import hashlib
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


class _SymbolicEscape(OperatorLoweringError):
    """A symbolic nodal term reached code that needs numbers (``jnp.dot``, ``np.asarray``, ``float`` ...).  The operator
    may still be linear: ``lower_diff_operator`` retries it with numeric probes before giving up."""


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
        raise _SymbolicEscape("nodal terms are symbolic during assembly and have no numeric value")

    def __array__(self, *a, **k):
        raise _SymbolicEscape(
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
        raise _SymbolicEscape("numpy function %s is not supported on a nodal gradient" % getattr(func, "__name__", func))


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


class _Probe:
    """Numeric stand-in for 'the basis function': its jet (phi, phi_x, phi_y, phi_xx, phi_yy) at the evaluation point is
    a fixed vector.  An operator that is linear in the term set maps the five unit jets to its five coefficients, whatever
    numeric library its body uses (``jnp.dot``, ``jnp.array([...])``, torch ...): the fallback of ``lower_diff_operator``
    for operators written against the reference (updes/operators.py:15-111 differentiates a real kernel there)."""

    def __init__(self, name, jet):
        self.name, self.jet = name, jet

    def __repr__(self):
        return "<numeric probe of the %s>" % self.name

    def __call__(self, *a, **k):
        raise OperatorLoweringError("the %s cannot be called during assembly; use the nodal_* term set" % self.name)


def _probe_of(center, monomial):
    if isinstance(center, _Probe):
        return center
    if isinstance(monomial, _Probe):
        return monomial
    return None


def _check_symbolic(center, monomial, what):
    if center is _CENTER or monomial is _MONOMIAL:
        return
    raise OperatorLoweringError(
        "%s is only available inside a diff_operator passed to pde_solver (symbolic lowering); for numeric "
        "field values use value / gradient / laplacian" % what)


def nodal_value(x, center=None, rbf=None, monomial=None):
    """rbf or monomial value at x (operators.py:15-32)."""
    p = _probe_of(center, monomial)
    if p is not None:
        return p.jet[0]
    _check_symbolic(center, monomial, "nodal_value")
    return Jet([1.0, 0.0, 0.0, 0.0, 0.0])


def nodal_gradient(x, center=None, rbf=None, monomial=None):
    """gradient w.r.t. x, NaN/inf at r = 0 replaced by 0 (operators.py:42-60)."""
    p = _probe_of(center, monomial)
    if p is not None:
        return np.array([p.jet[1], p.jet[2]])
    _check_symbolic(center, monomial, "nodal_gradient")
    return JetVector([Jet([0.0, 1.0, 0.0, 0.0, 0.0]), Jet([0.0, 0.0, 1.0, 0.0, 0.0])])


def nodal_laplacian(x, center=None, rbf=None, monomial=None):
    """trace of the Hessian w.r.t. x (operators.py:70-85)."""
    p = _probe_of(center, monomial)
    if p is not None:
        return p.jet[3] + p.jet[4]
    _check_symbolic(center, monomial, "nodal_laplacian")
    return Jet([0.0, 0.0, 0.0, 1.0, 1.0])


def nodal_div_grad(x, center=None, rbf=None, monomial=None, args=None):
    """args[0] phi_xx + args[1] phi_yy (operators.py:88-111)."""
    if args is None or len(args) != 2:
        raise OperatorLoweringError("nodal_div_grad needs args=(a, b)")
    p = _probe_of(center, monomial)
    if p is not None:
        return args[0] * p.jet[3] + args[1] * p.jet[4]
    _check_symbolic(center, monomial, "nodal_div_grad")
    return Jet([0.0, 0.0, 0.0, args[0], args[1]])


class BatchPoints(np.ndarray):
    """Coordinates of all rows at once, shape (2, rows): ``x[0]`` / ``x[1]`` are per-row vectors, so
    an operator written for one point ``x`` evaluates for every row in a single call."""
    def __new__(cls, arr):
        return np.asarray(arr, dtype=np.float64).view(cls)

    def __getitem__(self, idx):
        return np.asarray(super().__getitem__(idx))


# Number of rows of the batched operator call in flight (None outside one).  Lets the field evaluators
# recognise a (2, rows) coordinate array that the user's operator rebuilt from ``x`` (``np.stack([x[0],
# x[1]])`` -- the usual port of ``jnp.array([x[0], x[1]])``) and therefore lost the BatchPoints type.
_BATCH_ROWS = None


class _batch_rows:
    def __init__(self, rows): self.rows = rows
    def __enter__(self):
        global _BATCH_ROWS
        self.prev, _BATCH_ROWS = _BATCH_ROWS, self.rows
    def __exit__(self, *a):
        global _BATCH_ROWS
        _BATCH_ROWS = self.prev


def _fields_table(diff_args, N, Ni):
    """``jnp.stack(diff_args, axis=-1)`` restricted to the internal rows (assembly.py:114, :128); arrays of
    length N+M (coefficient vectors) are accepted and cut to the nodes, as ``fields[i]`` does in the reference."""
    if not diff_args:
        return None
    cols = []
    for a in diff_args:
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 0 or a.shape[0] < Ni:
            raise ValueError("diff_args fields must have one value per node (N = %d), got shape %s" % (N, a.shape))
        cols.append(a[:Ni])
    try:
        return np.stack(cols, axis=-1)                                                 # (Ni, ..., nf)
    except ValueError as e:
        raise ValueError("diff_args fields must share one shape: %s" % e)


def _coef_table(out, Ni):
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
    return tab


def lower_diff_operator(diff_operator, cloud, rbf, diff_args=None):
    """Return the (Ni, 5) coefficient tables of the user's operator on RBF columns and on monomial columns
    (reference assembly.py:93-137).

    Fast path: ONE call per basis family with symbolic jets and the coordinates / fields of all internal
    rows at once (``x`` of shape (2, Ni), ``fields`` of shape (nf, Ni)).  Operators that only make sense
    for one node at a time -- a Python ``if`` on a coordinate, ``float(x[0])``, shape-dependent code:
    the reference vmaps the operator over nodes, so ``x`` is a (2,) point there (assembly.py:126-130) --
    make that call raise; the operator is then evaluated row by row with exactly the reference's
    per-node arguments (``x`` (2,), ``fields[i]`` (nf,)).  Operators whose body hands the nodal terms to a numeric
    library (``jnp.dot(U, nodal_gradient(...))`` as in the reference's demos) cannot be traced symbolically; they are
    evaluated row by row with NUMERIC probes instead (unit jets in, coefficients out, linearity checked per row).
    Genuinely unsupported operators raise ``OperatorLoweringError`` from every path."""
    Ni, N = cloud.Ni, cloud.N
    F = _fields_table(diff_args, N, Ni)
    x = BatchPoints(cloud.sorted_nodes[:Ni].T)
    fields = np.ones((1, Ni)) if F is None else np.moveaxis(F, 0, -1)                 # (..., nf, Ni); assembly.py:114

    def run_batch(center, monomial):
        with _batch_rows(Ni):
            return _coef_table(diff_operator(x, center, rbf, monomial, fields), Ni)

    def symbolic_row(i, center, monomial):
        out = diff_operator(np.array(cloud.sorted_nodes[i]), center, rbf, monomial, np.ones(1) if F is None else F[i])
        return _coef_table(out, 1)[0]

    def run_rows(center, monomial):
        tab = np.empty((Ni, 5))
        for i in range(Ni):
            tab[i] = symbolic_row(i, center, monomial)
        return tab

    def batch_agrees_with_rows(tab, center, monomial):
        """The batched call is an optimisation of the reference's one-node-per-call semantics (assembly.py:126-130).  A body
        written for one node can run on batched arrays and mean something else (a reduction over 'all' axes of x, say):
        sample rows are re-evaluated node by node, and a batched table that does not reproduce them is discarded."""
        for i in sorted({0, Ni // 3, Ni // 2, Ni - 1}):
            try:
                row = symbolic_row(i, center, monomial)
            except Exception:
                return True            # only the batched form runs: nothing to compare with
            if not np.allclose(tab[i], row, rtol=1e-12, atol=1e-300):
                return False
        return True

    def run_numeric(family):
        """Numeric probes: the operator sees plain numbers, so its body may use any array library.  The zero jet must
        give 0 (no affine part), the five unit jets give the coefficients, and one mixed jet must give the same mixture
        (linearity).  First all rows at once (jets as (Ni,) vectors beside the batched x), then -- if the body only
        makes sense for one node, the reference's vmap semantics -- one node per call."""
        mix = (0.7, -1.3, 0.45, 1.9, -0.6)
        jets = [(0.0,) * 5] + [tuple(float(j == k) for j in range(5)) for k in range(5)] + [mix]
        name = "rbf centre" if family == 0 else "monomial"

        def check(vals, where):
            """vals: (7, R) outputs for the seven jets -> (R, 5) coefficients"""
            if np.any(vals[0] != 0.0):
                raise OperatorLoweringError(
                    "the differential operator returns %g for a vanishing basis function%s: it has an affine part or does "
                    "not involve the basis function (affine parts belong in the rhs operator)" % (vals[0].flat[np.argmax(vals[0] != 0)], where))
            tab = vals[1:6].T.copy()
            want = tab @ np.asarray(mix)
            tol = 1e-9 * np.maximum(1.0, np.maximum(np.abs(want), np.max(np.abs(tab), axis=1)))
            if not np.all(np.abs(vals[6] - want) <= tol):
                raise OperatorLoweringError(
                    "the differential operator is not linear in nodal_value / nodal_gradient / nodal_laplacian / "
                    "nodal_div_grad%s (a mixed probe does not give the mixture of the unit probes)" % where)
            return tab

        def evaluate(xarg, farg, jet, rows):
            pr = _Probe(name, jet)
            out = diff_operator(xarg, pr if family == 0 else None, rbf, None if family == 0 else pr, farg)
            out = np.asarray(out, dtype=np.float64)
            if rows == 1 and out.size != 1:
                raise OperatorLoweringError("the differential operator must return a scalar per node, got shape %s" % (out.shape,))
            return np.broadcast_to(out.reshape(-1) if out.ndim else out, (rows,))

        ones = np.ones(1)

        def one_row(i):
            xi, fi = np.array(cloud.sorted_nodes[i]), (ones if F is None else F[i])
            return check(np.stack([evaluate(xi, fi, jet, 1) for jet in jets]), " at node %d" % i)[0]

        try:
            with _batch_rows(Ni):
                vals = np.stack([evaluate(x, fields, tuple(np.full(Ni, v) for v in jet), Ni) for jet in jets])
            tab = check(vals, "")
            # a body written for ONE node can run on the batched arrays and mean something else (jnp.sum over both
            # axes, say): the batched table is only kept if it reproduces per-node evaluation on sample rows
            for i in sorted({0, Ni // 3, Ni // 2, Ni - 1}):
                if not np.array_equal(tab[i], one_row(i)) and not np.allclose(tab[i], one_row(i), rtol=1e-14, atol=0):
                    raise ValueError("batched evaluation differs from per-node evaluation")
            return tab
        except OperatorLoweringError:
            raise
        except Exception:
            pass
        tab = np.empty((Ni, 5))
        for i in range(Ni):
            tab[i] = one_row(i)
        return tab

    tabs = []
    for family, (center, monomial) in enumerate(((_CENTER, None), (None, _MONOMIAL))):
        try:
            try:
                tab = run_batch(center, monomial)
                if Ni > 0 and not batch_agrees_with_rows(tab, center, monomial):
                    tab = run_rows(center, monomial)
            except _SymbolicEscape:
                raise
            except OperatorLoweringError:
                raise
            except Exception:
                tab = run_rows(center, monomial)      # reference semantics: one node per call
        except _SymbolicEscape as escape:
            # a symbolic term reached numeric code (jnp.dot, np.asarray, float ...): ask the operator with numbers
            try:
                tab = run_numeric(family)
            except OperatorLoweringError:
                raise
            except Exception as e:
                raise OperatorLoweringError("the differential operator could not be evaluated numerically either (%s: %s); "
                                            "symbolic lowering had failed with: %s" % (type(e).__name__, e, escape)) from e
        except OperatorLoweringError:
            raise
        except Exception as first:
            # neither batched nor per-node symbolic evaluation ran (e.g. a foreign array library rejected the symbolic
            # terms with its own exception type): numeric probes, then give up with the explicit error
            try:
                tab = run_numeric(family)
            except OperatorLoweringError:
                raise
            except Exception as e:
                raise OperatorLoweringError("the differential operator could not be lowered: symbolic evaluation raised %s: %s; "
                                            "numeric probing raised %s: %s" % (type(first).__name__, first, type(e).__name__, e)) from e
        if not np.all(np.isfinite(tab)):
            raise OperatorLoweringError("operator coefficients are not finite")
        tabs.append(tab)
    return tabs[0], tabs[1]


# ==================================================================================================
# Field evaluators (matrix-free, GPU)
# ==================================================================================================
def _points(x):
    """-> (pts (R,2) ndarray, layout) with layout in {'single', 'batch', 'rows'}."""
    if isinstance(x, BatchPoints):
        return np.ascontiguousarray(np.asarray(x).T), "batch"
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        if x.shape[0] != 2:
            raise ValueError("a single evaluation point must have 2 coordinates, got shape %s" % (x.shape,))
        return x.reshape(1, 2), "single"
    if _BATCH_ROWS is not None and x.ndim == 2 and x.shape == (2, _BATCH_ROWS):
        # inside a batched operator call: coordinates rebuilt from x[0], x[1] (lost the BatchPoints type)
        return np.ascontiguousarray(x.T), "batch"
    if x.ndim != 2 or x.shape[1] != 2:
        raise ValueError("evaluation points must be (2,), (R, 2) or the operator's own x; got shape %s" % (x.shape,))
    return np.ascontiguousarray(x), "rows"


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


value_vec = value_vec_ = value                     # operators.py:149-150 (the evaluators take one point or many)
gradient_vec = gradient_vec_ = gradient            # operators.py:183-184
laplacian_vec = laplacian
divergence_vec = divergence


def integrate_field(field, cloud, rbf, max_degree):
    """Integral over the unit square of a field given by its COEFFICIENTS, by the reference's weighted sum of evaluator
    values (operators.py:381-451): weight 1 on an (Nx-1) x (Ny-1) set of interior points, 1/2 on edge points, 1/4 on
    corners, times the cell area.  Restated as written, including how the reference builds its point sets -- the interior
    points come from ``array(meshgrid(qx, qy)).reshape(nb_squares, 2)`` (a reshape of the stacked coordinate grids, not a
    pairing of them) and the top-left corner re-uses the bottom-right point -- because its own known-answer test
    (updes/tests/test_integrals.py:83, pi/12 to 1e-1) is defined on exactly this sum.  One batched evaluator call."""
    if not (hasattr(cloud, "Nx") and hasattr(cloud, "Ny")):
        raise AssertionError("The cloud must be a SquareCloud instance")
    Nx, Ny = cloud.Nx, cloud.Ny
    nb_squares = (Nx - 1) * (Ny - 1)
    area = (1 / (Nx - 1)) * (1 / (Ny - 1))
    qx = np.linspace(1 / (Nx - 1), 1 - 1 / (Nx - 1), Nx - 1)
    qy = np.linspace(1 / (Ny - 1), 1 - 1 / (Ny - 1), Ny - 1)
    inner = np.array(np.meshgrid(qx, qy)).reshape(nb_squares, 2)
    ey = np.linspace(1 / (Ny - 1), 1 - 1 / (Ny - 1), Ny - 2)
    ex = np.linspace(1 / (Nx - 1), 1 - 1 / (Nx - 1), Nx - 2)
    left, right = np.stack((np.zeros(Ny - 2), ey), axis=-1), np.stack((np.ones(Ny - 2), ey), axis=-1)
    bottom, top = np.stack((ex, np.zeros(Nx - 2)), axis=-1), np.stack((ex, np.ones(Nx - 2)), axis=-1)
    corners = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 0.0], [1.0, 1.0]])        # bl, br, "tl" (= br in the reference), tr
    pts = np.concatenate([inner, left, right, bottom, top, corners], axis=0)
    w = np.concatenate([np.ones(nb_squares), np.full(2 * (Ny - 2) + 2 * (Nx - 2), 0.5), np.full(4, 0.25)])
    vals = np.asarray(value(pts, field, cloud.sorted_nodes, rbf), dtype=np.float64)
    return float(np.dot(w, vals) * area)


def interpolate_field(field, cloud1, cloud2):
    """Carry a nodal field from ``cloud1`` to ``cloud2``: the same nodes, numbered differently because the
    boundary types differ (reference operators.py:457-480; a permutation, no arithmetic).  Used by the
    projection loop of demos/NavierStokes/30_channel_flow_blowing_suction.py:161-213."""
    assert cloud1.N == cloud2.N, "the two clouds do not contain the same number of nodes"
    field = np.asarray(field)
    field_orig = field[np.asarray(cloud1._new_of_old)]          # original numbering
    return field_orig[np.asarray(cloud2._old_of_new)]


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
            # operators.py:535-536 indexes betas with i - node_ids[0] ("TODO: this assumes consistent ordering"): when two
            # Robin facets interleave in the renumbering (North / South own the corners, so their last node comes after
            # the East / West nodes) the offset runs past the end, and a jax array then returns its LAST element
            # (out-of-bounds gathers clamp).  Kept as the reference computes it (quirk Q8).
            last = len(node_ids) - 1
            if node_ids and node_ids[-1] - node_ids[0] > last:
                warnings.warn("Robin facet %r: its nodes are not numbered contiguously (another Robin facet interleaves); "
                              "betas are assigned by offset from the first node, clamped, as the reference does" % f_id)
            for i in node_ids:
                robin_coeffs[i] = betas[min(i - node_ids[0], last)]
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
    """``SteadySol(vals, coeffs, mat)`` -- stands in for the reference's namedtuple
    ``PDESolution`` (utils.py:148): attribute access, indexing, unpacking, ``len``, ``_fields``,
    ``_replace`` and ``_asdict`` behave as on the namedtuple.  The one difference is that ``mat`` (the
    reference's B = diffMat inv(A)[:, :N], assembly.py:396-401, which needs inv(A) and a second N x N
    matrix) is computed the first time it is read; the thunk holds host-side descriptors only, never the
    factored system, so keeping solutions around does not pin HBM."""
    __slots__ = ("vals", "coeffs", "_mat", "_mat_fn")
    _fields = ("vals", "coeffs", "mat")

    def __init__(self, vals, coeffs, mat=None, mat_fn=None):
        self.vals, self.coeffs, self._mat, self._mat_fn = vals, coeffs, mat, mat_fn

    @property
    def mat(self):
        if self._mat is None and self._mat_fn is not None:
            self._mat = self._mat_fn()
            self._mat_fn = None
        return self._mat

    def __len__(self):
        return 3

    def __getitem__(self, i):
        if isinstance(i, slice):
            return tuple(self[k] for k in range(3)[i])
        i = i + 3 if i < 0 else i
        if i == 0: return self.vals
        if i == 1: return self.coeffs
        if i == 2: return self.mat
        raise IndexError("tuple index out of range")

    def __iter__(self):
        yield self.vals
        yield self.coeffs
        yield self.mat

    def _replace(self, **kw):
        bad = set(kw) - set(self._fields)
        if bad:
            raise ValueError("Got unexpected field names: %r" % sorted(bad))
        lazy = "mat" not in kw and self._mat is None
        return SteadySol(kw.get("vals", self.vals), kw.get("coeffs", self.coeffs),
                         None if lazy else kw.get("mat", self._mat), self._mat_fn if lazy else None)

    def _asdict(self):
        return {"vals": self.vals, "coeffs": self.coeffs, "mat": self.mat}

    def __repr__(self):
        return "PDESolution(vals=%r, coeffs=%r, mat=%s)" % (self.vals, self.coeffs, "<lazy>" if self._mat is None else repr(self._mat))


FactorizationError = RuntimeError     # internal device-side failures (never a numerical status) raise RuntimeError

# Optional phase timing of pde_solver (tools/e2e_phases.py): set to a list to collect (phase, seconds since the previous
# mark) with a device synchronize at every mark; None (default) costs one comparison per mark.
TRACE = None
_trace_t = [0.0]


def _mark(phase):
    if TRACE is None:
        return
    import time
    import torch
    torch.cuda.synchronize()
    now = time.perf_counter()
    TRACE.append((phase, now - _trace_t[0]))
    _trace_t[0] = now


class _System:
    """An assembled + factored collocation system resident on the GPU."""

    def __init__(self, cloud, kind, param, M, table, equilibrate=True):
        self.kind, self.param, self.M = kind, param, M
        self.cloud = cloud                   # the cache key uses id(cloud): keep it alive while cached
        _mark("  digest + row descriptors (host)")
        self.rows = _asm.DeviceRows(cloud, table)
        _mark("  row descriptors -> device")
        self.K = _asm.assemble_system(self.rows, kind, param, M)
        _mark("  allocate K + assembly")
        self.n = cloud.N + M
        self.lu = LUFactorization(self.K, self.n).factor(equilibrate=equilibrate)
        _mark("  equilibration + LU")

    def check(self):
        """Raise on internal failures; return the LAPACK-style zero-pivot status (0 = none)."""
        return self.lu.check()

    def solve(self, rhs, refine=1):
        """rhs: (n,) numpy or CUDA tensor -> coefficients (n,) CUDA tensor.  `refine` steps of iterative
        refinement with a matrix-free residual (no second copy of K)."""
        torch = self.rows.torch
        b = torch.as_tensor(rhs, dtype=torch.float64).to(self.K.device)
        c = self.lu.solve(b.clone())
        for _ in range(refine):
            r = b - _asm.apply_rows(self.rows, self.kind, self.param, self.M, c.view(1, -1))[0]
            c = c + self.lu.solve(r)
        return c

    def nbytes(self):
        return self.K.numel() * 8

    @staticmethod
    def predict_nbytes(n, world=1):
        return n * _asm.padded_ld(n) * 8


# ---- multi-GPU opt-in ---------------------------------------------------------------------------------
_DIST = {"enabled": False, "group": None, "grid": None}


def enable_distributed(group=None, grid=None):
    """Shard every subsequent ``pde_solver*`` call over the ranks of ``group`` (default: the default
    ``torch.distributed`` process group, which must already be initialised with the NCCL backend, one
    process per GPU).  EVERY rank of the group must then call the solver with identical arguments.
    Opt-in on purpose: an initialised process group alone (e.g. inside an unrelated data-parallel job)
    does not change how ``pde_solver`` runs.

    ``grid=(P, Q)`` with P > 1 selects the 2-D block-cyclic layout (updes_b200/grid2d.py; P * Q must equal the
    group size); the default is the measured 1 x Q column-block-cyclic layout (updes_b200/distributed.py)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("enable_distributed() needs an initialised torch.distributed process group (backend nccl)")
    if grid is not None:
        P, Q = int(grid[0]), int(grid[1])
        if P < 1 or Q < 1 or P * Q != dist.get_world_size(group):
            raise ValueError("grid=(%d, %d) does not match the %d ranks of the process group" % (P, Q, dist.get_world_size(group)))
        grid = (P, Q) if P > 1 else None           # 1 x Q is the default path
    _DIST["enabled"], _DIST["group"], _DIST["grid"] = True, group, grid
    clear_cache()


def disable_distributed():
    _DIST["enabled"], _DIST["group"], _DIST["grid"] = False, None, None
    clear_cache()


def _dist_world(distributed=None):
    """(world, rank, group) of the sharded path, or (1, 0, None) when it is off."""
    on = _DIST["enabled"] if distributed is None else bool(distributed)
    if not on:
        return 1, 0, None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("distributed=True needs an initialised torch.distributed process group")
    g = _DIST["group"]
    return dist.get_world_size(g), dist.get_rank(g), g


def default_block_width(n, world):
    """Column-block width of the multi-GPU layout: 2048 for large systems, halved until every rank owns at least
    ~12 blocks.  With the fire-and-forget GEMM epilogue the update loses little at a smaller inner dimension
    (35.5 / 35.1 / 34.4 TFLOP/s at k = 2048 / 1024 / 512), while few blocks per rank cost balance: at 4 GPUs and
    n = 90 003, nb = 2048 (11 blocks per rank) left the last rank 250 ms (6 %) more trailing update than the first."""
    nb = 2048
    while nb > 64 and n // (nb * world) < 12:
        nb //= 2
    return nb


class _DistSystem:
    """The same system sharded column-block-cyclically over the ranks of a process group
    (updes_b200/distributed.py).  Every rank calls pde_solver with identical arguments."""

    def __init__(self, cloud, kind, param, M, table, world, rank, group=None):
        from .distributed import ColumnBlockCyclic, CudaBackend, DistributedLU
        self.kind, self.param, self.M = kind, param, M
        self.cloud = cloud                   # the cache key uses id(cloud): keep it alive while cached
        self.group = group
        self.rows = _asm.DeviceRows(cloud, table)
        self.n = cloud.N + M
        self.layout = ColumnBlockCyclic(self.n, default_block_width(self.n, world), world)
        self.be = CudaBackend(self.layout, rank)
        self.be.assemble(self.rows, kind, param, M)
        self.dlu = DistributedLU(self.layout, rank, self.be, group=group).equilibrate().factor()
        self.K = self.be.local

    def check(self):
        import torch.distributed as dist
        t = self.be.info.clone()
        lo = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        if int(lo.item()) < 0:
            raise FactorizationError("the panel kernel's grid barrier timed out on some rank (info = %d)" % int(lo.item()))
        self.be.check_sweeps()
        return int(t.item())

    def solve(self, rhs, refine=1):
        torch = self.rows.torch
        b = torch.as_tensor(rhs, dtype=torch.float64).to(self.be.device)
        c = self.dlu.solve(b)
        for _ in range(refine):
            r = b - _asm.apply_rows(self.rows, self.kind, self.param, self.M, c.view(1, -1))[0]
            c = c + self.dlu.solve(r)
        return c

    def nbytes(self):
        return self.be.nbytes()

    @staticmethod
    def predict_nbytes(n, world):
        nb = default_block_width(n, world)
        cols = -(-n // world) + nb
        return n * _asm.padded_ld(cols) * 8 + 2 * (n * nb + nb) * 8


class _DistSystem2D:
    """The same system on a P x Q block-cyclic process grid (updes_b200/grid2d.py); selected by
    ``enable_distributed(grid=(P, Q))`` with P > 1.  Same interface as ``_DistSystem``."""

    def __init__(self, cloud, kind, param, M, table, grid, rank, group=None):
        from .grid2d import BlockCyclic2D, CudaKernels2D, DistributedLU2D
        self.kind, self.param, self.M = kind, param, M
        self.cloud = cloud
        self.group = group
        self.rows = _asm.DeviceRows(cloud, table)
        self.n = cloud.N + M
        P, Q = grid
        self.layout = BlockCyclic2D(self.n, default_block_width(self.n, max(P, Q)), P, Q)
        self.dlu = DistributedLU2D(self.layout, rank, CudaKernels2D(), group=group)
        self.dlu.assemble(self.rows, kind, param, M)
        self.dlu.equilibrate().factor()
        self.K = self.dlu.local

    def check(self):
        status = self.dlu.zero_pivot()
        if status < 0:
            raise FactorizationError("the panel kernel's grid barrier timed out on some rank (info = %d)" % status)
        self.dlu.K.check_sweeps()
        return status

    def solve(self, rhs, refine=1):
        torch = self.rows.torch
        b = torch.as_tensor(rhs, dtype=torch.float64).to(self.dlu.device)
        c = self.dlu.solve(b)
        for _ in range(refine):
            r = b - _asm.apply_rows(self.rows, self.kind, self.param, self.M, c.view(1, -1))[0]
            c = c + self.dlu.solve(r)
        return c

    def nbytes(self):
        return self.dlu.nbytes()

    @staticmethod
    def predict_nbytes(n, grid):
        P, Q = grid
        nb = default_block_width(n, max(P, Q))
        mloc, ld = -(-n // (nb * P)) * nb + nb, -(-n // (nb * Q)) * nb + nb
        return 8 * (mloc * ld + (n + nb) * nb + 2 * (mloc * nb + nb) + 3 * nb * ld)


def _make_dist_system(cloud, kind, param, M, table, world, rank, group):
    """(builder, predicted bytes) of the sharded system in the layout enable_distributed() selected."""
    grid = _DIST["grid"]
    if grid is not None and grid[0] * grid[1] == world:
        return (lambda: _DistSystem2D(cloud, kind, param, M, table(), grid, rank, group)), _DistSystem2D.predict_nbytes(cloud.N + M, grid)
    return (lambda: _DistSystem(cloud, kind, param, M, table(), world, rank, group)), _DistSystem.predict_nbytes(cloud.N + M, world)


# ---- cache of factored systems -----------------------------------------------------------------------------
_CACHE: "OrderedDict[tuple, object]" = OrderedDict()
_CACHE_FRACTION = 0.88        # of the device's total memory, shared by all cached systems


def clear_cache():
    """Drop every cached factorisation (frees the HBM they hold, whatever solutions are still referenced)."""
    _CACHE.clear()
    try:
        import torch
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
    except Exception:
        pass


def cache_budget_bytes():
    torch = _lib.require_cuda()
    free, total = torch.cuda.mem_get_info()
    return int(_CACHE_FRACTION * total)


def _cached_system(key, build, need_bytes):
    """LRU cache keyed on (cloud, rbf, lowered operator, ...).  Space for the new system is made BEFORE it
    is built (evicting least-recently-used systems until the prediction fits the device budget), so the
    peak is never more than the budget plus transients."""
    sys_ = _CACHE.get(key)
    if sys_ is not None:
        _CACHE.move_to_end(key)
        return sys_
    torch = _lib.require_cuda()
    budget = cache_budget_bytes()
    evicted = False
    while _CACHE and sum(s.nbytes() for s in _CACHE.values()) + need_bytes > budget:
        _CACHE.popitem(last=False)
        evicted = True
    if evicted or torch.cuda.mem_get_info()[0] < need_bytes * 1.02:
        torch.cuda.empty_cache()
    sys_ = build()
    _CACHE[key] = sys_
    return sys_


def _digest(*arrays):
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        if a is None:
            h.update(b"-")
        else:
            a = np.ascontiguousarray(a)
            h.update(str(a.shape).encode()); h.update(a.tobytes())
    return h.hexdigest()


def _interp_system(cloud, kind, param, M, distributed=None):
    """Factorisation of A = [[Phi P], [P^T 0]] (assembly.py:62-90), cached per (cloud, rbf, M).  On the
    multi-GPU path A is sharded like K (an (N+M)^2 matrix per rank would not fit at the sizes that path is for)."""
    _asm.require_global_support(cloud)
    world, rank, group = _dist_world(distributed)
    key = ("A", id(cloud), kind, param, M, world, _DIST["grid"] if world > 1 else None)
    n = cloud.N + M
    if world > 1:
        build, need = _make_dist_system(cloud, kind, param, M, lambda: _asm.build_interpolation_rows(cloud), world, rank, group)
        return _cached_system(key, build, need)
    return _cached_system(key, lambda: _System(cloud, kind, param, M, _asm.build_interpolation_rows(cloud)),
                          _System.predict_nbytes(n))


def core_compute_coefficients(field, cloud, rbf, nb_monomials):
    """inv(A) @ [field; 0] (assembly.py:404-410) through the cached LU of A."""
    kind, param = identify_rbf(rbf)
    rhs = np.concatenate([np.asarray(field, dtype=np.float64), np.zeros(nb_monomials)])
    return _interp_system(cloud, kind, param, nb_monomials).solve(rhs).cpu().numpy()


def compute_coefficients(field, cloud, rbf, max_degree):
    """assembly.py:413-418"""
    return core_compute_coefficients(field, cloud, rbf, compute_nb_monomials(max_degree, cloud.dim))


get_field_coefficients = compute_coefficients      # assembly.py:423-430


def gradient_vals(x, field, cloud, rbf, max_degree):
    """Gradient at x of a field given by its nodal VALUES (operators.py:186-205): coefficients through the cached LU of
    A, then the matrix-free evaluator.  ``x``: one point (2,) or points (R, 2)."""
    return gradient(x, compute_coefficients(field, cloud, rbf, max_degree), cloud.sorted_nodes, rbf)


def laplacian_vals(x, field, cloud, rbf, max_degree):
    """Laplacian at x of a field given by its nodal values (operators.py:354-368)."""
    return laplacian(x, compute_coefficients(field, cloud, rbf, max_degree), cloud.sorted_nodes, rbf)


gradient_vals_vec = gradient_vals_vec_ = gradient_vals          # operators.py:207-208
laplacian_vals_vec = laplacian_vals_vec_ = laplacian_vals       # operators.py:370-371


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
    try:
        with _batch_rows(Ni):
            q_int = np.asarray(rhs_operator(x, cloud.sorted_nodes, rbf, fields), dtype=np.float64)
        q[:Ni] = np.broadcast_to(q_int, (Ni,))
    except OperatorLoweringError:
        raise
    except Exception:
        # per-node evaluation, the reference's vmap semantics (assembly.py:466-469): x is one (2,) point
        for i in range(Ni):
            q[i] = float(rhs_operator(np.array(cloud.sorted_nodes[i]), cloud.sorted_nodes, rbf, fields))
    for f_id in cloud.facet_types.keys():
        assert f_id in boundary_conditions.keys(), "facets and boundary functions don't match ids"
        bd = boundary_conditions[f_id]
        ids = np.asarray(cloud.facet_nodes[f_id], dtype=int)
        q[ids] = _eval_on_nodes(bd, cloud.sorted_nodes[ids]) if callable(bd) else np.asarray(bd, dtype=np.float64)
    return q


def pde_solver(diff_operator, rhs_operator, cloud, boundary_conditions, rbf, max_degree, diff_args=None,
               rhs_args=None, *, refine=1, distributed=None):
    """Solve a linear PDE by global RBF collocation (reference operators.py:559-618).

    Same positional arguments and result as the reference.  The factorisation is cached on (cloud, rbf,
    max_degree, lowered operator, Robin betas): repeated calls with an unchanged left-hand side --
    the time loops of demos/Advection -- only pay for the right-hand side and two triangular sweeps.
    Keyword-only extras: ``refine`` (iterative-refinement steps, default 1), ``distributed`` (None = follow
    ``enable_distributed()``; True/False forces the sharded / single-GPU path for this call)."""
    _asm.require_global_support(cloud)
    _mark("enter")
    kind, param = identify_rbf(rbf)
    robin_coeffs, boundary_conditions = duplicate_robin_coeffs(dict(boundary_conditions), cloud)
    boundary_conditions = zerofy_periodic_cond(boundary_conditions, cloud)
    M = compute_nb_monomials(max_degree, cloud.dim)
    coef_phi, coef_pol = lower_diff_operator(diff_operator, cloud, rbf, diff_args)
    _mark("bc preparation + operator lowering")
    betas = np.array([robin_coeffs[k] for k in sorted(robin_coeffs)], dtype=np.float64) if robin_coeffs else None

    world, rank, group = _dist_world(distributed)
    key = ("K", id(cloud), kind, param, M, _digest(coef_phi, coef_pol, betas), world, _DIST["grid"] if world > 1 else None)
    n = cloud.N + M
    table_fn = lambda: _asm.build_operator_rows(cloud, coef_phi, coef_pol, betas)
    if world > 1:
        build, need = _make_dist_system(cloud, kind, param, M, table_fn, world, rank, group)
        system = _cached_system(key, build, need)
    else:
        system = _cached_system(key, lambda: _System(cloud, kind, param, M, table_fn()), _System.predict_nbytes(n))
    _mark("digest + row descriptors + assembly + equilibration + LU (or cache hit)")
    q = assemble_q(rhs_operator, boundary_conditions, cloud, rbf, M, rhs_args)
    _mark("right-hand side")
    coeffs_dev = system.solve(np.concatenate([q, np.zeros(M)]), refine=refine)
    zero_pivot = system.check()              # raises FactorizationError on internal failures
    _mark("solve + refinement + status")
    if zero_pivot:
        warnings.warn("collocation matrix is exactly singular (zero pivot at column %d)" % zero_pivot)
    # vals = [Phi P] c with the reference's zero-diagonal Phi (assembly.py:31-32, :404-410)
    torch = system.rows.torch
    own = torch.arange(cloud.N, dtype=torch.int32, device=coeffs_dev.device)
    jphi, jpol = _asm.eval_jets(kind, param, system.rows.centres, coeffs_dev.view(1, -1), system.rows.centres, own)
    vals = (jphi[0, :, 0] + jpol[0, :, 0]).cpu().numpy()
    coeffs = coeffs_dev.cpu().numpy()
    _mark("vals = [Phi P] c + download")
    # the lazy ``mat`` holds host descriptors only (never `system`): solutions must not pin the factors in HBM
    table = system.rows.table
    if world > 1:
        def mat_fn():
            raise NotImplementedError("SteadySol.mat is not available on the multi-GPU path")
    else:
        def mat_fn():
            return _reference_mat(cloud, kind, param, M, table)
    return SteadySol(vals, coeffs, None, mat_fn)


def _reference_mat(cloud, kind, param, M, table):
    """B = (diffMat @ inv(A))[:, :N] (assembly.py:396-401), for callers that read ``SteadySol.mat``.
    A is symmetric, so B[r, :] = (inv(A) diffMat[r, :]^T)[:N]: N right-hand sides against the LU of A.
    diffMat is re-assembled from the row descriptors (the factored K no longer holds it)."""
    N = cloud.N
    if N > 20000:
        raise MemoryError("SteadySol.mat (the reference's B = diffMat inv(A)[:, :N]) needs N = %d right-hand sides "
                          "against inv(A) and a second N x N matrix; it is only provided for N <= 20000" % N)
    rows = _asm.DeviceRows(cloud, table)
    D = _asm.assemble_system(rows, kind, param, M)[:N].contiguous()
    X = _interp_system(cloud, kind, param, M, distributed=False).lu.solve(D)
    return X[:, :N].cpu().numpy()


def pde_solver_jit_with_bc(diff_operator, rhs_operator, cloud, boundary_conditions, rbf, max_degree, diff_args=None,
                           rhs_args=None, **kw):
    """operators.py:621-625.  There is no tracing compiler here: identical to ``pde_solver``."""
    return pde_solver(diff_operator, rhs_operator, cloud, boundary_conditions, rbf, max_degree, diff_args, rhs_args, **kw)


def pde_solver_jit(diff_operator, rhs_operator, cloud, boundary_conditions, rbf, max_degree, diff_args=None,
                   rhs_args=None, **kw):
    """operators.py:650-683: boundary callables are turned into arrays first."""
    bc = boundary_conditions_func_to_arr(boundary_conditions, cloud)
    return pde_solver_jit_with_bc(diff_operator, rhs_operator, cloud, bc, rbf, max_degree, diff_args, rhs_args, **kw)


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


# ==================================================================================================
# Finite-difference helpers of the reference surface (host-only; operators.py:211-291, :483-509)
# ==================================================================================================
def _closest_opposite(cloud, ids, directions, chunk=None):
    """For every node ``ids[k]``: among the other nodes of its support whose unit offset u satisfies
    ``dot(directions[k], u) + 1 <= 0.1`` (a cone around MINUS the direction), the closest one -- the primitive shared by
    cartesian_gradient, enforce_cartesian_gradient_neumann and apply_neumann_conditions of the reference.  Returns
    (index or -1, its distance, distance of the LAST node of the support list = the farthest node).  Equidistant
    candidates: the reference keeps the last one in its support order (``<=``); here the one with the largest id."""
    xy = np.asarray(cloud.sorted_nodes, dtype=np.float64)
    ids = np.asarray(ids, dtype=np.int64)
    directions = np.asarray(directions, dtype=np.float64).reshape(len(ids), 2)
    keep = getattr(cloud, "support_size", cloud.N)
    keep = cloud.N if keep in ("max", None) else int(keep)
    chunk = chunk or max(1, 4_000_000 // max(cloud.N, 1))                     # a few (chunk, N) work arrays at a time
    close = np.full(len(ids), -1, dtype=np.int64)
    cdist = np.full(len(ids), 1e20)
    far = np.zeros(len(ids))
    for lo in range(0, len(ids), chunk):
        sl = slice(lo, min(lo + chunk, len(ids)))
        off = xy[None, :, :] - xy[ids[sl], None, :]
        nrm = np.sqrt(off[..., 0] ** 2 + off[..., 1] ** 2)
        nrm[np.arange(sl.stop - sl.start), ids[sl]] = np.inf                          # the node itself is not in its support
        if keep < cloud.N:                                                              # local supports: the keep - 1 nearest only
            kth = np.partition(nrm, keep - 2, axis=1)[:, keep - 2]
            nrm = np.where(nrm <= kth[:, None], nrm, np.inf)
        with np.errstate(invalid="ignore", divide="ignore"):
            cos = (off[..., 0] * directions[sl, 0:1] + off[..., 1] * directions[sl, 1:2]) / nrm
        cand = np.where((cos + 1.0 <= 1e-1) & np.isfinite(nrm), nrm, np.inf)
        best = cand.min(axis=1)
        has = np.isfinite(best)
        tie = cand == best[:, None]
        pick = cloud.N - 1 - np.argmax(tie[:, ::-1], axis=1)                           # largest id among the closest
        close[sl] = np.where(has, pick, -1)
        cdist[sl] = np.where(has, best, 1e20)
        far[sl] = np.where(np.isfinite(nrm), nrm, -np.inf).max(axis=1)
    return close, cdist, far


def cartesian_gradient_vec(node_ids, field, cloud):
    """Backward differences towards the nearest node "behind" each axis direction (operators.py:211-259), as the reference
    computes them: component d of node i is ``(field[i] - field[closest]) / vec_norm`` where ``vec_norm`` is left over from
    the LAST iteration of the reference's loop over the support list -- the distance to the farthest node of the support,
    not to the closest -- and a node with no neighbour behind direction d returns early with the remaining components 0.
    Kept as written (demos/NavierStokes/11_...:187 and 16_...:181 call it).  Rows are indexed by node id, like there."""
    node_ids = [int(i) for i in node_ids]
    field = np.asarray(field, dtype=np.float64)
    grad = np.zeros((len(node_ids), 2))
    ids = np.asarray(node_ids, dtype=np.int64)
    alive = np.ones(len(ids), dtype=bool)
    for d, direction in ((0, (1.0, 0.0)), (1, (0.0, 1.0))):
        close, _, far = _closest_opposite(cloud, ids, np.tile(direction, (len(ids), 1)))
        alive &= close >= 0                                       # `return final_grad` as soon as a direction has no neighbour
        vals = (field[ids] - field[np.where(close >= 0, close, 0)]) / far
        grad[ids[alive], d] = vals[alive]
    return grad


def cartesian_gradient(node_id, field, cloud, clip_val=None):
    """operators.py:211-252 for one node."""
    return _clip(_cartesian_one(node_id, field, cloud), clip_val)


def _cartesian_one(node_id, field, cloud):
    field = np.asarray(field, dtype=np.float64)
    i = int(node_id)
    out = np.zeros(2)
    for d, direction in ((0, (1.0, 0.0)), (1, (0.0, 1.0))):
        close, _, far = _closest_opposite(cloud, [i], [direction])
        if close[0] < 0:
            return out
        out[d] = (field[i] - field[close[0]]) / far[0]
    return out


def _neumann_nodes_and_normals(cloud):
    ids = [i for f, t in cloud.facet_types.items() if t == "n" for i in cloud.facet_nodes[f]]
    normals = np.array([np.asarray(cloud.outward_normals[i], dtype=np.float64) for i in ids]).reshape(len(ids), 2)
    return np.asarray(ids, dtype=np.int64), normals


def enforce_cartesian_gradient_neumann(field, grads, boundary_conditions, cloud, clip_val=None):
    """operators.py:262-291: on every Neumann node, BOTH gradient components are overwritten by the one-sided difference
    ``(field[i] - field[closest]) / distance`` towards the closest node opposite to the outward normal."""
    field = np.asarray(field, dtype=np.float64)
    grads = np.array(grads, dtype=np.float64, copy=True)
    ids, normals = _neumann_nodes_and_normals(cloud)
    if len(ids):
        close, cdist, _ = _closest_opposite(cloud, ids, normals)
        grads[ids] = ((field[ids] - field[close]) / cdist)[:, None] if grads.ndim == 2 else (field[ids] - field[close]) / cdist
    return _clip(grads, clip_val)


def apply_neumann_conditions(field, boundary_conditions, cloud):
    """operators.py:483-509: every Neumann node takes the value of the closest node opposite to its outward normal (a
    zero-flux condition imposed by copying; the boundary values themselves are not read, as in the reference)."""
    field = np.array(field, dtype=np.float64, copy=True)
    ids, normals = _neumann_nodes_and_normals(cloud)
    if len(ids):
        close, _, _ = _closest_opposite(cloud, ids, normals)
        # the reference updates node after node, so a Neumann node whose closest neighbour is an EARLIER Neumann node
        # reads the already updated value
        for i, c in zip(ids, close):
            field[i] = field[c]
    return field
