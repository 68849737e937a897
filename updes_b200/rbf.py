"""Built-in radial basis functions and monomials (reference ``updes/utils.py:19-144``).

The callables keep the reference's signature ``rbf(x, center, a=... | eps=...)`` and evaluate with
numpy, so user code can still call them; the solver never calls them -- it recognises them
(``identify_rbf``) and runs the closed forms inside the CUDA kernels.  Anything that is not one of
the five built-in kernels (possibly wrapped in ``functools.partial``) is rejected explicitly.
"""
from __future__ import annotations

import functools
import math

import numpy as np

RBF_CODES = {"polyharmonic": 0, "thin_plate": 1, "gaussian": 2, "multiquadric": 3, "inverse_multiquadric": 4}


def distance(node1, node2):
    """utils.py:19-22"""
    diff = np.asarray(node1, dtype=np.float64) - np.asarray(node2, dtype=np.float64)
    return np.sqrt(np.sum(diff * diff, axis=-1))


def multiquadric(x, center, eps=1.0):
    """Hardy's multiquadric, utils.py:30-35"""
    return np.sqrt(1 + (eps * distance(x, center)) ** 2)


def inverse_multiquadric(x, center, eps=1.0):
    """utils.py:37-42"""
    return 1.0 / np.sqrt(1 + (eps * distance(x, center)) ** 2)


def gaussian(x, center, eps=1.0):
    """utils.py:44-48"""
    return np.exp(-(eps * distance(x, center)) ** 2)


def polyharmonic(x, center, a=1):
    """Polyharmonic spline r^(2a+1), utils.py:50-55"""
    return distance(x, center) ** (2 * a + 1)


def thin_plate(x, center, a=1):
    """Thin-plate spline r^(2a) log r with the r = 0 value forced to 0, utils.py:63-69"""
    r = distance(x, center)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.nan_to_num(np.log(r) * r ** (2 * a), nan=0.0, posinf=0.0, neginf=0.0)


# radial profiles as functions of r (utils.py:30-66): what the kernels above apply to distance(x, center)
def multiquadric_func(r, eps):
    return np.sqrt(1 + (eps * np.asarray(r, dtype=np.float64)) ** 2)


def inv_multiquadric_func(r, eps):
    return 1.0 / np.sqrt(1 + (eps * np.asarray(r, dtype=np.float64)) ** 2)


def gaussian_func(r, eps):
    return np.exp(-(eps * np.asarray(r, dtype=np.float64)) ** 2)


def polyharmonic_func(r, a):
    return np.asarray(r, dtype=np.float64) ** (2 * a + 1)


def thin_plate_func(r, a):
    r = np.asarray(r, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.nan_to_num(np.log(r) * r ** (2 * a), nan=0.0, posinf=0.0, neginf=0.0)


def make_nodal_rbf(x, node, rbf):
    """utils.py:72-89: ``rbf`` given as a function of the distance, evaluated at |x - node| (polyharmonic when None -- the
    reference then calls its two-point kernel with one argument and fails; here r^3)."""
    r = distance(x, node)
    return polyharmonic_func(r, 1) if rbf is None else rbf(r)


for _f, _name in ((multiquadric, "multiquadric"), (inverse_multiquadric, "inverse_multiquadric"),
                  (gaussian, "gaussian"), (polyharmonic, "polyharmonic"), (thin_plate, "thin_plate")):
    _f.updes_kind = _name


def identify_rbf(rbf):
    """Return ``(kind, param)`` for a built-in kernel, possibly wrapped in ``functools.partial``
    (``partial(polyharmonic, a=1)``, ``partial(gaussian, eps=10.)``; reference usage:
    demos/Laplace/00_laplace_with_rbf.py:35, updes/tests/test_operators.py:40)."""
    kwargs = {}
    f = rbf
    while isinstance(f, functools.partial):
        if f.args:
            raise TypeError("rbf partials may only bind the keyword parameter (a= or eps=)")
        kwargs = {**f.keywords, **kwargs}
        f = f.func
    kind = getattr(f, "updes_kind", None)
    if kind is None:
        raise TypeError(
            "unsupported rbf %r: updes_b200 evaluates kernels in closed form on the GPU and supports only the "
            "built-in polyharmonic, thin_plate, gaussian, multiquadric and inverse_multiquadric" % (rbf,))
    if kind in ("polyharmonic", "thin_plate"):
        extra = set(kwargs) - {"a"}
        a = kwargs.get("a", 1)
        if extra or int(a) != a or a < 0 or (kind == "thin_plate" and a < 1):
            raise TypeError("%s takes an integer a >= %d" % (kind, 1 if kind == "thin_plate" else 0))
        return kind, float(int(a))
    extra = set(kwargs) - {"eps"}
    if extra:
        raise TypeError("%s takes eps= only" % kind)
    return kind, float(kwargs.get("eps", 1.0))


# ---- monomials (utils.py:92-144) ---------------------------------------------------------------
MONOMIAL_EXPONENTS = [(0, 0), (1, 0), (0, 1), (2, 0), (1, 1), (0, 2), (3, 0), (2, 1), (1, 2), (0, 3),
                      (4, 0), (3, 1), (2, 2), (1, 3), (0, 4)]


def make_monomial(x, id):
    """utils.py:92-134"""
    if id >= len(MONOMIAL_EXPONENTS):
        raise NotImplementedError("monomials of degree > 4 are not supported (reference utils.py:133-134)")
    a, b = MONOMIAL_EXPONENTS[id]
    x = np.asarray(x, dtype=np.float64)
    return x[..., 0] ** a * x[..., 1] ** b


def make_all_monomials(nb_monomials):
    """utils.py:136-139"""
    return [functools.partial(make_monomial, id=j) for j in range(nb_monomials)]


def compute_nb_monomials(max_degree, problem_dimension=2):
    """utils.py:142-144"""
    nb = math.comb(max_degree + problem_dimension, max_degree)
    if problem_dimension != 2 or nb > 15:
        raise NotImplementedError("2-D monomials up to degree 4 only (reference utils.py:92-134)")
    return nb
