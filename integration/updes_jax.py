"""jax.ffi wrappers over integration/updes_jax_ffi.cc -- SOURCE ONLY (JAX is not installed here).

Drop into the reference as ``updes/b200.py``; ``pde_solver`` below keeps the reference's signature
(updes/operators.py:559-618) while the assembly and the dense solve run in libupdes_b200.so."""
import ctypes

import numpy as np

import jax
import jax.numpy as jnp

from updes_b200.assembly import build_operator_rows, padded_ld          # array-library agnostic host logic
from updes_b200.operators import (boundary_conditions_func_to_arr, duplicate_robin_coeffs, lower_diff_operator,
                                  zerofy_periodic_cond)
from updes_b200.rbf import RBF_CODES, compute_nb_monomials, identify_rbf

_so = ctypes.cdll.LoadLibrary("libupdes_jax_ffi.so")
for _name in ("UpdesAssemble", "UpdesFactorSolve", "UpdesEvalJets"):
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_so, _name)), platform="CUDA")


def assemble_K(cloud, table, kind, param, M):
    n = cloud.N + M
    mi, mb = table.masks()
    call = jax.ffi.ffi_call("UpdesAssemble", jax.ShapeDtypeStruct((n, padded_ld(n)), jnp.float64))
    return call(jnp.asarray(cloud.sorted_nodes), jnp.asarray(table.p1), jnp.asarray(table.p2), jnp.asarray(table.cphi1),
                jnp.asarray(table.cphi2), jnp.asarray(table.cpol1), jnp.asarray(table.cpol2), jnp.asarray(table.skip),
                kind=np.int32(RBF_CODES[kind]), param=np.float64(param), M=np.int32(M), mask_internal=np.int32(mi),
                mask_boundary=np.int32(mb), Ni=np.int32(cloud.Ni))


def factor_solve(K, rhs):
    n = K.shape[0]
    out = (jax.ShapeDtypeStruct(K.shape, K.dtype), jax.ShapeDtypeStruct((n,), K.dtype),
           jax.ShapeDtypeStruct((n,), jnp.int32), jax.ShapeDtypeStruct((1,), jnp.int32))
    _, x, _, info = jax.ffi.ffi_call("UpdesFactorSolve", out, input_output_aliases={0: 0})(K, rhs)
    return x, info


def pde_solver(diff_operator, rhs_operator, cloud, boundary_conditions, rbf, max_degree, diff_args=None, rhs_args=None):
    kind, param = identify_rbf(rbf)
    M = compute_nb_monomials(max_degree, cloud.dim)
    robin, bcs = duplicate_robin_coeffs(boundary_conditions_func_to_arr(boundary_conditions, cloud), cloud)
    bcs = zerofy_periodic_cond(bcs, cloud)
    coef_phi, coef_pol = lower_diff_operator(diff_operator, cloud, rbf, diff_args)
    betas = np.array([robin[k] for k in sorted(robin)]) if robin else None
    K = assemble_K(cloud, build_operator_rows(cloud, coef_phi, coef_pol, betas), kind, param, M)
    q = np.zeros(cloud.N + M)
    # ... right-hand side as in updes_b200.operators.assemble_q, then:
    coeffs, info = factor_solve(K, jnp.asarray(q))
    return coeffs
