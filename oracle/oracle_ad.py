"""Autodiff-structured restatement of the Updes term set -- TEST INFRASTRUCTURE ONLY.

The reference obtains every derivative by JAX autodiff of the kernel callable
(``jax.grad`` / ``jax.jacfwd(jax.grad)``, operators.py:30-111) and cleans the r = 0 singularity with
``nan_to_num``.  JAX is not installed in this image, so this module rebuilds the same structure with
``torch.func`` in float64 on the CPU: same kernels written as functions of (x, center), same
transforms, same nan_to_num placement.  It pins the closed forms of ``updes_oracle.c`` (and through
them the CUDA kernels) and lets tests run *arbitrary* user operators written against the reference's
``diff_operator(x, center, rbf, monomial, fields)`` signature.  Small N only.
"""
from __future__ import annotations

from functools import partial

import torch
from torch.func import grad, jacfwd, vmap

torch.set_default_dtype(torch.float64)


def _nan_to_num(t):
    return torch.nan_to_num(t, nan=0.0, posinf=0.0, neginf=0.0)


# ---- utils.py:19-69 ---------------------------------------------------------------------------
def distance(node1, node2):
    diff = node1 - node2
    return torch.sqrt(diff @ diff)


def multiquadric(x, center, eps=1.0):
    return torch.sqrt(1 + (eps * distance(x, center)) ** 2)


def inverse_multiquadric(x, center, eps=1.0):
    return 1.0 / torch.sqrt(1 + (eps * distance(x, center)) ** 2)


def gaussian(x, center, eps=1.0):
    return torch.exp(-(eps * distance(x, center)) ** 2)


def polyharmonic(x, center, a=1):
    return distance(x, center) ** (2 * a + 1)


def thin_plate(x, center, a=1):
    r = distance(x, center)
    return _nan_to_num(torch.log(r) * r ** (2 * a))


RBFS = {"polyharmonic": polyharmonic, "thin_plate": thin_plate, "gaussian": gaussian,
        "multiquadric": multiquadric, "inverse_multiquadric": inverse_multiquadric}


def make_rbf(kind, param):
    if kind in ("polyharmonic", "thin_plate"):
        return partial(RBFS[kind], a=int(param))
    return partial(RBFS[kind], eps=float(param))


# ---- utils.py:92-139 ---------------------------------------------------------------------------
_MON = [(0, 0), (1, 0), (0, 1), (2, 0), (1, 1), (0, 2), (3, 0), (2, 1), (1, 2), (0, 3),
        (4, 0), (3, 1), (2, 2), (1, 3), (0, 4)]


def make_monomial(x, id):
    a, b = _MON[id]
    return (x[0] ** a) * (x[1] ** b) + 0.0 * x[0]     # keep a graph even for the constant monomial


def make_all_monomials(nb):
    return [partial(make_monomial, id=j) for j in range(nb)]


# ---- operators.py:15-111: the term set ----------------------------------------------------------
def nodal_value(x, center=None, rbf=None, monomial=None):
    if center is not None:
        return rbf(x, center)
    return monomial(x)


def nodal_gradient(x, center=None, rbf=None, monomial=None):
    if center is not None:
        return _nan_to_num(grad(rbf)(x, center))
    return grad(monomial)(x)


def nodal_laplacian(x, center=None, rbf=None, monomial=None):
    if center is not None:
        return _nan_to_num(torch.trace(jacfwd(grad(rbf))(x, center)))
    return torch.trace(jacfwd(grad(monomial))(x))


def nodal_div_grad(x, center=None, rbf=None, monomial=None, args=None):
    a = torch.as_tensor(args, dtype=torch.float64)
    matrix = torch.stack((a, a), dim=-1) if a.ndim == 1 else a
    if center is not None:
        return _nan_to_num(torch.trace(matrix * jacfwd(grad(rbf))(x, center)))
    return torch.trace(matrix * jacfwd(grad(monomial))(x))


# ---- assembly.py:93-137 with a real callable operator ------------------------------------------
def assemble_op_Phi_P(operator, cloud, rbf, nb_monomials, args=None):
    """Internal rows through the user's operator, by autodiff (assembly.py:93-137)."""
    N, Ni, M = cloud.N, cloud.Ni, nb_monomials
    nodes = torch.as_tensor(cloud.sorted_nodes, dtype=torch.float64)
    fields = torch.stack([torch.as_tensor(a, dtype=torch.float64) for a in args], dim=-1) if args else torch.ones((N, 1))
    opPhi = torch.zeros((Ni, N))
    opP = torch.zeros((Ni, M))

    def operator_rbf(x, center, f):
        return operator(x, center, rbf, None, f)

    op_vec = vmap(operator_rbf, in_dims=(None, 0, None))
    monomials = make_all_monomials(M)
    for i in range(Ni):
        support = torch.tensor([j for j in range(N) if j != i])          # cloud.py:110-112
        opPhi[i, support] = op_vec(nodes[i], nodes[support], fields[i])
        for j in range(M):
            opP[i, j] = operator(nodes[i], None, rbf, monomials[j], fields[i])
    return opPhi.numpy(), opP.numpy()


def assemble_Phi(cloud, rbf):
    """assembly.py:10-36"""
    N = cloud.N
    nodes = torch.as_tensor(cloud.sorted_nodes, dtype=torch.float64)
    Phi = torch.zeros((N, N))
    rbf_vec = vmap(rbf, in_dims=(None, 0))
    for i in range(N):
        support = torch.tensor([j for j in range(N) if j != i])
        Phi[i, support] = rbf_vec(nodes[i], nodes[support])
    return Phi.numpy()


def rbf_jet(rbf, x, center):
    """(phi, phi_x, phi_y, phi_xx, phi_yy) by autodiff with the reference's nan_to_num."""
    x = torch.as_tensor(x, dtype=torch.float64)
    c = torch.as_tensor(center, dtype=torch.float64)
    g = _nan_to_num(grad(rbf)(x, c))
    H = _nan_to_num(jacfwd(grad(rbf))(x, c))
    return torch.stack([rbf(x, c), g[0], g[1], H[0, 0], H[1, 1]]).numpy()
