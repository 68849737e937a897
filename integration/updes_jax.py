"""jax.ffi wrappers over integration/updes_jax_ffi.cc.  JAX is not installed in this environment: the module is
exercised by tests/run_jax_adapter.py over a jax.ffi stand-in and a mock of XLA's FFI binding API (tests/mock_xla),
not against a real jaxlib.

Drop into the reference as ``updes/b200.py``.  ``pde_solver`` keeps the reference's signature and result
(updes/operators.py:559-618: ``SteadySol(vals, coeffs, mat)``) while assembly, the dense solve and the field
evaluation run in libupdes_b200.so on XLA-owned device buffers and the XLA stream:

    reference step (file:line)                               here
    assemble_B  (assembly.py:366-401; inv + GEMM)            UpdesAssemble  -> K  (one matrix)
    assemble_q  (assembly.py:434-485)                        updes_b200.operators.assemble_q (host, O(N))
    lx.linear_solve(B, q, QR)  (operators.py:612-613)        UpdesFactorSolve (K donated, LU in place)
    core_compute_coefficients  (assembly.py:404-410)         -- c IS the solution of K c = [q; 0]
    sol_vals                                                 UpdesEvalJets: vals = [Phi P] c, matrix-free
"""
import ctypes
import os

import numpy as np

import jax
import jax.numpy as jnp

from updes_b200.assembly import build_operator_rows, padded_ld          # array-library agnostic host logic
from updes_b200.operators import (SteadySol, assemble_q, boundary_conditions_func_to_arr, duplicate_robin_coeffs,
                                  lower_diff_operator, zerofy_periodic_cond)
from updes_b200.rbf import RBF_CODES, compute_nb_monomials, identify_rbf

jax.config.update("jax_enable_x64", True)                                 # updes/config.py:15-16

_so = ctypes.cdll.LoadLibrary(os.environ.get("UPDES_JAX_FFI_LIB", "libupdes_jax_ffi.so"))
for _name in ("UpdesAssemble", "UpdesFactorSolve", "UpdesEvalJets"):
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_so, _name)), platform="CUDA")
_so.UpdesEvalJetsWorkspaceBytes.restype = ctypes.c_size_t
_so.UpdesEvalJetsWorkspaceBytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]


def assemble_K(cloud, table, kind, param, M):
    """K = [[op(Phi) op(P)], [bd(Phi) bd(P)], [P^T 0]] (replaces assembly.py:93-362 + :62-85) as one XLA buffer."""
    n = cloud.N + M
    mi, mb = table.masks()
    call = jax.ffi.ffi_call("UpdesAssemble", jax.ShapeDtypeStruct((n, padded_ld(n)), jnp.float64))
    i32 = lambda a: jnp.asarray(a, dtype=jnp.int32)
    return call(jnp.asarray(cloud.sorted_nodes), i32(table.p1), i32(table.p2), jnp.asarray(table.cphi1),
                jnp.asarray(table.cphi2), jnp.asarray(table.cpol1), jnp.asarray(table.cpol2), i32(table.skip),
                kind=np.int32(RBF_CODES[kind]), param=np.float64(param), M=np.int32(M), mask_internal=np.int32(mi),
                mask_boundary=np.int32(mb), Ni=np.int32(cloud.Ni))


def factor_solve(K, rhs):
    """In-place LU with partial pivoting + one solve.  K is donated so XLA does not copy 65 GB at n = 90k.
    Returns (x, info): info = 0, or the 1-based column of an exactly zero pivot (LAPACK convention)."""
    n = K.shape[0]
    out = (jax.ShapeDtypeStruct(K.shape, K.dtype), jax.ShapeDtypeStruct((n,), K.dtype),
           jax.ShapeDtypeStruct((n,), jnp.int32), jax.ShapeDtypeStruct((1,), jnp.int32))
    _, x, _, info = jax.ffi.ffi_call("UpdesFactorSolve", out, input_output_aliases={0: 0})(K, rhs)
    return x, info


def eval_jets(centres, coeffs, pts, kind, param, M):
    """(nf, R, 5) RBF and polynomial parts of (value, dx, dy, dxx, dyy) of the fields given by `coeffs` (nf, N+M)
    at `pts` (R, 2): the reference's value / gradient / laplacian evaluators (operators.py:118-351)."""
    nf, R, N = coeffs.shape[0], pts.shape[0], centres.shape[0]
    ws = max(int(_so.UpdesEvalJetsWorkspaceBytes(N, R, nf)), 8)
    out = (jax.ShapeDtypeStruct((nf, R, 5), jnp.float64), jax.ShapeDtypeStruct((nf, R, 5), jnp.float64),
           jax.ShapeDtypeStruct(((ws + 7) // 8,), jnp.float64))
    jphi, jpol, _ = jax.ffi.ffi_call("UpdesEvalJets", out)(centres, coeffs, pts, kind=np.int32(RBF_CODES[kind]),
                                                          param=np.float64(param), M=np.int32(M))
    return jphi, jpol


def pde_solver(diff_operator, rhs_operator, cloud, boundary_conditions, rbf, max_degree, diff_args=None, rhs_args=None):
    """operators.py:559-618 with the dense work on the GPU.  Operators outside the nodal_* term set raise
    updes_b200.OperatorLoweringError (lower_diff_operator), as the north_star requires."""
    kind, param = identify_rbf(rbf)
    M = compute_nb_monomials(max_degree, cloud.dim)
    robin, bcs = duplicate_robin_coeffs(boundary_conditions_func_to_arr(dict(boundary_conditions), cloud), cloud)
    bcs = zerofy_periodic_cond(bcs, cloud)
    coef_phi, coef_pol = lower_diff_operator(diff_operator, cloud, rbf, diff_args)
    betas = np.array([robin[k] for k in sorted(robin)], dtype=np.float64) if robin else None
    K = assemble_K(cloud, build_operator_rows(cloud, coef_phi, coef_pol, betas), kind, param, M)
    q = assemble_q(rhs_operator, bcs, cloud, rbf, M, rhs_args)            # host, O(N) (+ one A-solve per nodal field)
    coeffs, info = factor_solve(K, jnp.concatenate([jnp.asarray(q), jnp.zeros(M)]))
    if int(info[0]) < 0:
        raise RuntimeError("updes_b200: device-side wait timed out inside the LU (info = %d)" % int(info[0]))
    centres = jnp.asarray(cloud.sorted_nodes)
    jphi, jpol = eval_jets(centres, coeffs[None, :], centres, kind, param, M)
    # the reference's Phi has a zero diagonal (Q1, assembly.py:31-32); the evaluator sums every centre, so take
    # the self term phi(0) c_i out again (0 for polyharmonic / thin-plate, 1 * c_i for the other three kernels)
    phi0 = 0.0 if kind in ("polyharmonic", "thin_plate") else 1.0
    vals = jphi[0, :, 0] + jpol[0, :, 0] - phi0 * coeffs[: cloud.N]
    return SteadySol(vals, coeffs, None)


pde_solver_jit = pde_solver      # operators.py:650-683: nothing to trace, the custom calls are already compiled
