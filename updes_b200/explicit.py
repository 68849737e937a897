"""The reference's explicit-matrix functions under their own names (``updes/assembly.py:10-401``):
``assemble_Phi / assemble_P / assemble_A / assemble_invert_A / assemble_op_Phi_P / assemble_bd_Phi_P / assemble_B``.

``pde_solver`` never forms these matrices separately (one system K is assembled in HBM and factored in place,
DESIGN.md section 1), and the reference's demos never call them directly.  They are provided for scripts
and tests that do: every function assembles on the GPU with the same kernels as the solver, through the C-ABI, and
returns host ``numpy`` arrays with the reference's shapes.  Explicit (N+M)^2 host matrices only make sense for small
clouds: N > 20 000 raises ``MemoryError`` (as ``SteadySol.mat`` does).
"""
from __future__ import annotations

import numpy as np

from . import assembly as _asm
from .rbf import identify_rbf

_LIMIT = 20000


def _check_size(cloud, what):
    if cloud.N > _LIMIT:
        raise MemoryError("%s returns an explicit host matrix; it is only provided for N <= %d (N = %d). "
                          "pde_solver / pde_solver_jit work on the device-resident system instead" % (what, _LIMIT, cloud.N))


def _system_matrix(cloud, table, kind, param, M):
    rows = _asm.DeviceRows(cloud, table)
    K = _asm.assemble_system(rows, kind, param, M)
    return K[:, :cloud.N + M].cpu().numpy()


def assemble_A(cloud, rbf, nb_monomials=2):
    """[[Phi, P], [P^T, 0]] (assembly.py:62-85); Phi has a zero diagonal (the node is dropped from its own support)."""
    _check_size(cloud, "assemble_A")
    kind, param = identify_rbf(rbf)
    return _system_matrix(cloud, _asm.build_interpolation_rows(cloud), kind, param, int(nb_monomials))


def assemble_Phi(cloud, rbf):
    """Phi[i, j] = rbf(x_i, x_j), j != i (assembly.py:10-36)."""
    return np.ascontiguousarray(assemble_A(cloud, rbf, 1)[:cloud.N, :cloud.N])


def assemble_P(cloud, nb_monomials):
    """P[i, j] = monomial_j(x_i) (assembly.py:39-59)."""
    from .rbf import polyharmonic
    M = int(nb_monomials)
    return np.ascontiguousarray(assemble_A(cloud, polyharmonic, M)[:cloud.N, cloud.N:cloud.N + M])


def assemble_invert_A(cloud, rbf, nb_monomials):
    """inv(A) (assembly.py:87-90), from the cached LU of A: n right-hand sides (the columns of the identity)."""
    _check_size(cloud, "assemble_invert_A")
    from .operators import _interp_system
    kind, param = identify_rbf(rbf)
    M = int(nb_monomials)
    n = cloud.N + M
    system = _interp_system(cloud, kind, param, M, distributed=False)
    torch = system.rows.torch
    eye = torch.zeros((n, system.K.shape[1]), dtype=torch.float64, device=system.K.device)
    eye[:, :n] = torch.eye(n, dtype=torch.float64, device=system.K.device)
    X = system.lu.solve(eye)                       # row i = inv(A) e_i = column i of the inverse
    return np.ascontiguousarray(X[:, :n].cpu().numpy().T)


def assemble_op_Phi_P(operator, cloud, rbf, nb_monomials, args):
    """op(Phi) (Ni, N) and op(P) (Ni, M): the differential operator on internal rows (assembly.py:93-137)."""
    _check_size(cloud, "assemble_op_Phi_P")
    from .operators import lower_diff_operator
    kind, param = identify_rbf(rbf)
    M = int(nb_monomials)
    coef_phi, coef_pol = lower_diff_operator(operator, cloud, rbf, args)
    K = _system_matrix(cloud, _asm.build_operator_rows(cloud, coef_phi, coef_pol), kind, param, M)
    return np.ascontiguousarray(K[:cloud.Ni, :cloud.N]), np.ascontiguousarray(K[:cloud.Ni, cloud.N:])


def _betas(cloud, robin_coeffs):
    if cloud.Nr == 0:
        return None
    if not robin_coeffs:
        return np.zeros(cloud.Nr)                  # assembly.py:199-202: no coefficients -> zeros
    betas = np.array([float(robin_coeffs[k]) for k in sorted(robin_coeffs)], dtype=np.float64)
    if betas.shape[0] != cloud.Nr:
        raise ValueError("robin_coeffs must hold one coefficient per Robin node (%d), got %d" % (cloud.Nr, betas.shape[0]))
    return betas


def assemble_bd_Phi_P(cloud, rbf, nb_monomials, robin_coeffs=None):
    """bd(Phi) (N - Ni, N) and bd(P) (N - Ni, M): Dirichlet, Neumann, Robin and periodic rows (assembly.py:141-362).
    ``robin_coeffs``: {sorted node id: beta} as ``duplicate_robin_coeffs`` returns it."""
    _check_size(cloud, "assemble_bd_Phi_P")
    kind, param = identify_rbf(rbf)
    M = int(nb_monomials)
    K = _system_matrix(cloud, _asm.build_operator_rows(cloud, np.zeros((cloud.Ni, 5)), None, _betas(cloud, robin_coeffs)), kind, param, M)
    return np.ascontiguousarray(K[cloud.Ni:cloud.N, :cloud.N]), np.ascontiguousarray(K[cloud.Ni:cloud.N, cloud.N:])


def assemble_B(operator, cloud, rbf, nb_monomials, diff_args, robin_coeffs):
    """B = ([[op(Phi) op(P)], [bd(Phi) bd(P)]] inv(A))[:, :N] (assembly.py:366-401): N right-hand sides against the LU of A."""
    _check_size(cloud, "assemble_B")
    from .operators import _reference_mat, lower_diff_operator
    kind, param = identify_rbf(rbf)
    coef_phi, coef_pol = lower_diff_operator(operator, cloud, rbf, diff_args)
    table = _asm.build_operator_rows(cloud, coef_phi, coef_pol, _betas(cloud, robin_coeffs))
    return _reference_mat(cloud, kind, param, int(nb_monomials), table)
