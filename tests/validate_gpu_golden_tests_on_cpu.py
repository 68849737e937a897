#!/usr/bin/env python
"""How tests/test_gpu_zy_reference_golden.py was validated without a GPU (round 2: the file was written after the GPU
budget was spent).  Every CUDA-backed call the tests make is replaced HERE, in this script only, by the CPU oracle
(assembly -> closed forms of oracle/updes_oracle.c, the solve -> LAPACK LU with the product's row equilibration and one
refinement step, evaluators -> oracle.eval_field), and each test function is then run as written.  This checks the TEST
LOGIC -- shapes, golden keys, tolerances with their margins, the host orchestration around the device calls -- not the
kernels (those are checked against the same oracle on a GPU by the rest of the `-m gpu` suite).

    python tests/validate_gpu_golden_tests_on_cpu.py
"""
import importlib
import os
import sys

import numpy as np
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))   # (this file lives in tests/: only tests/ may use the oracle)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import oracle as O  # noqa: E402
import updes_b200 as u  # noqa: E402
from updes_b200 import operators as ops  # noqa: E402
from updes_b200.rbf import identify_rbf  # noqa: E402

O.build()
T = importlib.import_module("test_gpu_zy_reference_golden")
SKIP = {"test_explicit_assembly_functions_against_the_reference",        # wrappers: see the block at the end
        "test_config2_time_steps_against_the_reference_advection_demo"}  # asserts kernel launch counts


def betas_of(cloud, table):
    Ni = cloud.Ni
    lo = Ni + cloud.Nd + cloud.Nn
    return table.cphi1[lo:lo + cloud.Nr, 0] if cloud.Nr else None


def fake_assemble(cloud, table, kind, param, M):
    if table.Ni == cloud.N:
        return O.assemble_A(cloud, kind, param, M)
    return O.assemble_K(cloud, kind, param, M, table.cphi1[:cloud.Ni], betas_of(cloud, table))


def lu_solve_like_the_product(K, rhs):
    m = np.abs(K).max(1)
    sc = np.ldexp(1.0, 1 - np.frexp(m)[1])
    lu = sla.lu_factor(K * sc[:, None])
    c = sla.lu_solve(lu, rhs * sc)
    return c + sla.lu_solve(lu, (rhs - K @ c) * sc)


class Sol:
    pass


def fake_coefficients(field, cloud, rbf, M):
    kind, param = identify_rbf(rbf)
    return lu_solve_like_the_product(O.assemble_A(cloud, kind, param, M), np.concatenate([np.asarray(field, dtype=float), np.zeros(M)]))


def fake_solver(diff_operator, rhs_operator, cloud, boundary_conditions, rbf, max_degree, diff_args=None, rhs_args=None):
    kind, param = identify_rbf(rbf)
    robin, bcs = u.duplicate_robin_coeffs(dict(u.boundary_conditions_func_to_arr(boundary_conditions, cloud)), cloud)
    bcs = u.zerofy_periodic_cond(bcs, cloud)
    M = u.compute_nb_monomials(max_degree, 2)
    coef, _ = u.lower_diff_operator(diff_operator, cloud, rbf, diff_args)
    betas = np.array([robin[k] for k in sorted(robin)]) if robin else None
    K = O.assemble_K(cloud, kind, param, M, coef, betas)
    q = ops.assemble_q(rhs_operator, bcs, cloud, rbf, M, rhs_args)
    A = O.assemble_A(cloud, kind, param, M)
    s = Sol()
    s.coeffs = lu_solve_like_the_product(K, np.concatenate([q, np.zeros(M)]))
    s.vals = A[:cloud.N] @ s.coeffs
    s.mat = (K[:cloud.N] @ np.linalg.inv(A))[:, :cloud.N]
    return s


def evaluator(which):
    def f(x, cf, centers, rbf, clip_val=None):
        kind, param = identify_rbf(rbf)
        pts = np.asarray(x).T if isinstance(x, u.BatchPoints) else np.asarray(x, dtype=float).reshape(-1, 2)
        cf = np.asarray(cf, dtype=float)
        ev = lambda c, w: O.eval_field(pts, np.asarray(centers), np.ascontiguousarray(c), kind, param, w)
        if which == "gradient":
            out = np.stack([ev(cf, "dx"), ev(cf, "dy")], axis=1)
            return out[0] if np.ndim(x) == 1 else out
        if which == "divergence":
            return ev(cf[:, 0], "dx") + ev(cf[:, 1], "dy")
        return ev(cf, which)
    return f


T._assemble = fake_assemble
ops.core_compute_coefficients = fake_coefficients
ops.compute_coefficients = u.compute_coefficients = u.get_field_coefficients = \
    lambda field, cloud, rbf, max_degree: fake_coefficients(field, cloud, rbf, u.compute_nb_monomials(max_degree, 2))
for name in ("value", "gradient", "laplacian", "divergence"):
    for alias in (name, name + "_vec"):
        setattr(u, alias, evaluator(name))
        setattr(ops, alias, evaluator(name))
u.assemble_A = lambda cloud, rbf, M=2: O.assemble_A(cloud, *identify_rbf(rbf), int(M))
u.pde_solver_jit = u.pde_solver_jit_with_bc = u.pde_solver = fake_solver
ops.pde_solver_jit_with_bc = fake_solver                                  # pde_multi_solver calls it through the module

ran = 0
for name in sorted(n for n in dir(T) if n.startswith("test_") and n not in SKIP):
    fn = getattr(T, name)
    params = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
    for vals in (params[0].args[1] if params else [()]):
        fn(*vals)
        ran += 1
        print("ok   ", name, vals if vals else "")
print("%d test functions of tests/test_gpu_zy_reference_golden.py hold on the CPU stand-ins" % ran)
