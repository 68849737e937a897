"""GPU parity of the whole path: pde_solver vs the reference formulation (inv + GEMM + QR) on the CPU."""
from functools import partial

import numpy as np
import pytest

import updes_b200 as u
import os
import sys

from helpers import (CONFIG1_FACETS, CONFIG2_FACETS, advdiff_op, backward_error, cloud_from_golden, exact_solution,
                     laplace_op)

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import configs  # noqa: E402  (tools/configs.py: BASELINE.json configs 1-3 as the reference's demos define them)

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


def _assert_solution_parity(got_vals, got_coeffs, ref_vals, K, rhs, A_rows, what):
    """north_star: 'solutions within 1e-8 relative (or within cond-scaled backward error 1e-13 for ill-conditioned
    clouds)'.  Both the reference formulation (inv + GEMM + QR) and the product (LU + refinement) approximate the same
    discrete solution; that solution is computed here with extended-precision refinement, and the test asserts
      (1) product vs exact <= 1e-8                              -- the 1e-8 clause against the thing being approximated
      (2) backward error of the product's coefficients <= 1e-13 -- the ill-conditioned clause, always
      (3) product vs reference formulation <= 1e-8, or <= 4x the reference formulation's OWN distance from the exact
          solution when that distance exceeds 1e-8 (inv(A) loses cond(A) ~ 1e9 digits on these clouds)."""
    exact_vals, _ = exact_solution(K, rhs, A_rows)
    e_prod, e_ref, d = _rel(got_vals, exact_vals), _rel(ref_vals, exact_vals), _rel(got_vals, ref_vals)
    berr = backward_error(K, got_coeffs, rhs)
    print("%s: product-vs-exact %.2e  reference-vs-exact %.2e  product-vs-reference %.2e  backward error %.2e"
          % (what, e_prod, e_ref, d, berr))
    assert e_prod <= 1e-8, (what, e_prod)
    assert berr <= 1e-13, (what, berr)
    assert d <= max(1e-8, 4.0 * e_ref), (what, d, e_ref)


def test_readme_laplace_matches_reference_formulation(oracle):
    """README example (config 1): 30x20, polyharmonic a=1, degree 1, Dirichlet sine / Neumann."""
    cloud = u.SquareCloud(Nx=30, Ny=20, facet_types=CONFIG1_FACETS)
    sine = lambda c: np.sin(np.pi * c[0])
    zero = lambda c: 0.0
    bcs = {"South": zero, "West": zero, "North": sine, "East": zero}
    sol = u.pde_solver_jit(diff_operator=laplace_op(u), rhs_operator=lambda x, centers, rbf, fields: 0.0, cloud=cloud,
                           boundary_conditions=bcs, rbf=u.polyharmonic, max_degree=1)
    bc_arr = u.boundary_conditions_func_to_arr(bcs, cloud)
    q = oracle.assemble_q(cloud, np.zeros(cloud.Ni), bc_arr)
    coef = np.tile([0.0, 0.0, 0.0, 1.0, 1.0], (cloud.Ni, 1))
    vals, coeffs, B = oracle.reference_solve(cloud, "polyharmonic", 1, 1, coef, q)
    assert _rel(sol.vals, vals) <= 1e-8                   # north_star: solutions within 1e-8 relative
    K = oracle.assemble_K(cloud, "polyharmonic", 1, 3, coef)
    A = oracle.assemble_A(cloud, "polyharmonic", 1, 3)
    _assert_solution_parity(sol.vals, sol.coeffs, vals, K, np.concatenate([q, np.zeros(3)]), A[:cloud.N], "config 1")
    # namedtuple surface (utils.py:148)
    v, c, m = sol
    assert v is sol.vals and c is sol.coeffs and sol[0] is sol.vals and sol._replace(vals=None).coeffs is sol.coeffs
    # analytic solution of demos/Laplace/00_laplace_with_rbf.py:109-110
    xy = cloud.sorted_nodes
    exact = np.sin(np.pi * xy[:, 0]) * np.cosh(np.pi * xy[:, 1]) / np.cosh(np.pi)
    assert np.max(np.abs(sol.vals - exact)) <= 2e-2
    # SteadySol.mat is the reference's B (assembly.py:396-401)
    assert _rel(sol.mat, B) <= 1e-6


def test_advection_periodic_time_steps(oracle):
    """Config 2: periodic adv-diff, factor once, several solves with rhs = value(u)/DT."""
    DT = 1e-4
    cloud = u.SquareCloud(Nx=35, Ny=35, facet_types=CONFIG2_FACETS, noise_key=11)
    rbf = partial(u.polyharmonic, a=1)
    xy = cloud.sorted_nodes
    u0 = np.exp(-((xy[:, 0] - 0.35) ** 2 + (xy[:, 1] - 0.5) ** 2) / (2 * 0.1 ** 2))
    zero = lambda c: 0.0
    bcs = {k: zero for k in CONFIG2_FACETS}
    rhs = lambda x, centers, rbf, fields: u.value(x, fields[:, 0], centers, rbf) / DT
    coef = np.tile([1 / DT, 100.0, 0.0, -0.08, -0.08], (cloud.Ni, 1))
    uu = u0.copy()
    from updes_b200 import _lib
    u.clear_cache()
    for step in range(3):
        _lib.profile_enable(True)        # clears the per-class records
        l0 = _lib.launch_count()
        prev = uu
        sol = u.pde_solver_jit(diff_operator=advdiff_op(u, DT), rhs_operator=rhs, rhs_args=[prev], cloud=cloud,
                               boundary_conditions=bcs, rbf=rbf, max_degree=0)
        uu = sol.vals
        # reference formulation on the CPU, one step from the SAME previous field (parity is per step on identical
        # inputs; two trajectories fed by their own rounding drift apart by the propagator's growth, which is a
        # property of the discrete problem): coefficients of the previous field, value(x)/DT on internal nodes
        A = oracle.assemble_A(cloud, "polyharmonic", 1, 1)
        cprev = np.linalg.solve(A, np.concatenate([prev, np.zeros(1)]))
        q_int = oracle.eval_field(cloud.sorted_nodes[:cloud.Ni], cloud.sorted_nodes, cprev, "polyharmonic", 1, "value") / DT
        q = oracle.assemble_q(cloud, q_int, {k: np.zeros(len(cloud.facet_nodes[k])) for k in cloud.facet_types})
        ref, _, _ = oracle.reference_solve(cloud, "polyharmonic", 1, 0, coef, q)
        Kc = oracle.assemble_K(cloud, "polyharmonic", 1, 1, coef)
        _assert_solution_parity(uu, sol.coeffs, ref, Kc, np.concatenate([q, np.zeros(1)]), A[:cloud.N], "config 2 step %d" % step)
        if step == 0:
            launches_first = _lib.launch_count() - l0
        else:
            # factor-once / solve-many: later steps launch no assembly, panel or GEMM kernels at all
            assert _lib.launch_count() - l0 < 0.2 * launches_first, (launches_first, _lib.launch_count() - l0)
            assert _lib.profile_read("gemm")[2] == 0 and _lib.profile_read("panel")[2] == 0 and _lib.profile_read("assemble")[2] == 0
    _lib.profile_enable(False)


def test_gaussian_constant_field_gmsh_like(oracle):
    """Reference test_operators.py restated on a square cloud: all-Neumann, diff = nodal_value,
    rhs = 12 -> constant field; its gradient and divergence vanish (atol 1e-2 in the reference)."""
    cloud = u.SquareCloud(Nx=22, Ny=22, facet_types={k: "n" for k in ("South", "West", "North", "East")}, noise_key=12)
    rbf = partial(u.gaussian, eps=10.0)
    op = lambda x, c, r, m, f: u.nodal_value(x, c, r, m)
    sol = u.pde_solver(op, lambda x, centers, rbf, fields: 12.0, cloud, {k: (lambda c: 0.0) for k in cloud.facet_types},
                       rbf, 1)
    g = u.gradient_vec(cloud.sorted_nodes, sol.coeffs, cloud.sorted_nodes, rbf)
    d = u.divergence_vec(cloud.sorted_nodes, np.stack([sol.coeffs, sol.coeffs], -1), cloud.sorted_nodes, rbf)
    assert np.allclose(np.linalg.norm(g, axis=-1)[:cloud.Ni], 0, atol=1e-2)
    assert np.allclose(d[:cloud.Ni], 0, atol=1e-2)
    assert np.allclose(sol.vals[:cloud.Ni], 12.0, atol=1e-6)


def test_config3_pressure_poisson_on_gmsh_cloud(oracle):
    """Config 3 (phi solve of demos/NavierStokes/30_...:89-94): Laplacian with Neumann walls/inflow and a
    Dirichlet outflow on the mesh.msh cloud; solution vs the reference formulation on the CPU."""
    cloud, _ = cloud_from_golden("mesh_msh_cloud_phi.npz")
    rng = np.random.default_rng(2)
    src = rng.normal(size=cloud.Ni)
    rhs = lambda x, centers, rbf, fields: src
    bcs = {k: np.zeros(len(v)) for k, v in cloud.facet_nodes.items()}
    sol = u.pde_solver_jit(laplace_op(u), rhs, cloud, bcs, partial(u.polyharmonic, a=1), 1)
    coef = np.tile([0.0, 0.0, 0.0, 1.0, 1.0], (cloud.Ni, 1))
    q = oracle.assemble_q(cloud, src, bcs)
    vals, coeffs, _ = oracle.reference_solve(cloud, "polyharmonic", 1, 1, coef, q)
    K = oracle.assemble_K(cloud, "polyharmonic", 1, 3, coef)
    A = oracle.assemble_A(cloud, "polyharmonic", 1, 3)
    _assert_solution_parity(sol.vals, sol.coeffs, vals, K, np.concatenate([q, np.zeros(3)]), A[:cloud.N], "config 3 phi solve")


def test_pde_multi_solver_genuinely_coupled():
    """pde_multi_solver (operators.py:696-771): two equations coupled through each other's latest values,
        lap(u0) - (1 + u1^2) u0 = 0,     lap(u1) + (x + u0) d(u1)/dx = -1,
    against a hand-rolled Jacobi-Picard loop (every sweep solves all equations with the PREVIOUS sweep's values:
    the reference builds the list of solutions before updating sols_vals, operators.py:752-768)."""
    cloud = u.SquareCloud(Nx=16, Ny=14, facet_types=CONFIG1_FACETS, noise_key=2)
    rbf = partial(u.polyharmonic, a=1)
    zero, one = (lambda c: 0.0), (lambda c: 1.0)
    bcs = [{"South": zero, "West": zero, "North": one, "East": zero}, {"South": zero, "West": one, "North": zero, "East": zero}]
    op0 = lambda x, c, r, m, f: u.nodal_laplacian(x, c, r, m) - (1.0 + f[1] ** 2) * u.nodal_value(x, c, r, m)
    op1 = lambda x, c, r, m, f: u.nodal_laplacian(x, c, r, m) + (x[0] + f[0]) * u.nodal_gradient(x, c, r, m)[0]
    rhs0 = lambda x, centers, rbf, fields: 0.0
    rhs1 = lambda x, centers, rbf, fields: -1.0
    u0 = [np.zeros(cloud.N), np.zeros(cloud.N)]
    nb_iters = 3
    sols = u.pde_multi_solver([op0, op1], [rhs0, rhs1], cloud, bcs, rbf, 1, nb_iters=nb_iters, diff_args=[u0, u0], rhs_args=None)
    # hand-rolled loop with explicit coefficient tables through the plain single-equation solver
    prev = [u0[0].copy(), u0[1].copy()]
    Ni, xs = cloud.Ni, cloud.sorted_nodes[:cloud.Ni, 0]
    for _ in range(nb_iters):
        f0, f1 = prev[0], prev[1]
        opA = lambda x, c, r, m, f, f1=f1: u.nodal_laplacian(x, c, r, m) - (1.0 + f1[:Ni] ** 2) * u.nodal_value(x, c, r, m)
        opB = lambda x, c, r, m, f, f0=f0: u.nodal_laplacian(x, c, r, m) + (xs + f0[:Ni]) * u.nodal_gradient(x, c, r, m)[0]
        sA = u.pde_solver_jit(opA, rhs0, cloud, bcs[0], rbf, 1)
        sB = u.pde_solver_jit(opB, rhs1, cloud, bcs[1], rbf, 1)
        prev = [sA.vals, sB.vals]
    assert _rel(sols[0].vals, prev[0]) <= 1e-10 and _rel(sols[1].vals, prev[1]) <= 1e-10
    # the coupling is real: the decoupled problems have visibly different solutions
    dec0 = u.pde_solver_jit(lambda x, c, r, m, f: u.nodal_laplacian(x, c, r, m) - u.nodal_value(x, c, r, m), rhs0, cloud, bcs[0], rbf, 1)
    assert _rel(sols[0].vals, dec0.vals) >= 1e-4
    # and swapping the order of the unknowns in diff_args changes the answer (the plumbing is positional)
    swapped = u.pde_multi_solver([op0, op1], [rhs0, rhs1], cloud, bcs, rbf, 1, nb_iters=1, diff_args=[[np.ones(cloud.N), np.zeros(cloud.N)]] * 2)
    straight = u.pde_multi_solver([op0, op1], [rhs0, rhs1], cloud, bcs, rbf, 1, nb_iters=1, diff_args=[[np.zeros(cloud.N), np.ones(cloud.N)]] * 2)
    assert _rel(swapped[0].vals, straight[0].vals) >= 1e-4


def test_config3_projection_loop_matches_the_reference_formulation(oracle):
    """Config 3 as the demo runs it (demos/NavierStokes/30_channel_flow_blowing_suction.py:161-213): two iterations of the
    u / v / phi projection loop on the two complementary clouds of the reference's mesh.msh, fields carried across
    with interpolate_field -- against the same loop restated with the oracle's reference formulation
    (inv(A) coefficients for rhs fields, B = D inv(A), QR)."""
    cv, _ = cloud_from_golden("mesh_msh_cloud_vel.npz")
    cp, _ = cloud_from_golden("mesh_msh_cloud_phi.npz")
    Re, nb_iter, M = 100.0, 2, 3
    uu, vv, p_, hist = configs.config3_projection_loop(u, cv, cp, nb_iter=nb_iter, Re=Re)

    # ---- oracle restatement of the same loop -------------------------------------------------------------
    bc_u, bc_v, bc_phi = configs.config3_boundary_arrays(cv, cp)
    Av, Ap = oracle.assemble_A(cv, "polyharmonic", 1, M), oracle.assemble_A(cp, "polyharmonic", 1, M)
    coefs_of = lambda A, f: np.linalg.solve(A, np.concatenate([f, np.zeros(M)]))            # inv(A) [f; 0], assembly.py:404-410
    ev = lambda cloud, c, which, pts=None: oracle.eval_field(cloud.sorted_nodes if pts is None else pts, cloud.sorted_nodes, c, "polyharmonic", 1, which)
    ru, rv = np.zeros(cv.N), np.zeros(cv.N)
    rp_ = np.zeros(cp.N)
    lap_coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cp.Ni, 1))
    worst = 0.0
    for it in range(nb_iter):
        p = u.interpolate_field(rp_, cp, cv)
        cpv = coefs_of(Av, p)
        coef = np.stack([np.zeros(cv.Ni), ru[:cv.Ni], rv[:cv.Ni], np.full(cv.Ni, -1 / Re), np.full(cv.Ni, -1 / Re)], axis=1)
        qu = oracle.assemble_q(cv, -ev(cv, cpv, "dx", cv.sorted_nodes[:cv.Ni]), bc_u)
        qv = oracle.assemble_q(cv, -ev(cv, cpv, "dy", cv.sorted_nodes[:cv.Ni]), bc_v)
        ustar, _, _ = oracle.reference_solve(cv, "polyharmonic", 1, 1, coef, qu)
        vstar, _, _ = oracle.reference_solve(cv, "polyharmonic", 1, 1, coef, qv)
        u_, v_ = u.interpolate_field(ustar, cv, cp), u.interpolate_field(vstar, cv, cp)
        cu_, cv_ = coefs_of(Ap, u_), coefs_of(Ap, v_)
        div = ev(cp, cu_, "dx", cp.sorted_nodes[:cp.Ni]) + ev(cp, cv_, "dy", cp.sorted_nodes[:cp.Ni])
        qphi = oracle.assemble_q(cp, div, bc_phi)
        phi, cphi, _ = oracle.reference_solve(cp, "polyharmonic", 1, 1, lap_coef, qphi)
        rp_ = rp_ + phi
        gradphi_ = np.stack([ev(cp, cphi, "dx"), ev(cp, cphi, "dy")], axis=-1)
        gradphi = u.interpolate_field(gradphi_, cp, cv)
        ru, rv = ustar - gradphi[:, 0], vstar - gradphi[:, 1]
        h = hist[it]
        for name, got, want in (("u*", h[0], ustar), ("v*", h[1], vstar), ("phi", h[2], phi), ("u", h[3], ru), ("v", h[4], rv)):
            d = _rel(got, want)
            worst = max(worst, d)
            print("config 3 iteration %d %-3s product-vs-reference-formulation %.2e" % (it, name, d))
    # Every solve here inherits the reference formulation's own error (inv(A) at cond(A) ~ 1e9 puts it 2.5e-8 from the
    # exact discrete solution on this cloud, see test_config3_pressure_poisson_on_gmsh_cloud), and the loop feeds results
    # back in, so the loop-level bound is the cond-scaled one; single solves are held to 1e-8 against the exact solution above.
    # (measured on the CPU with LAPACK LU in place of the product: up to 2.6e-6 on v after the gradient of phi is subtracted)
    assert worst <= 2e-5, worst
    assert np.all(np.isfinite(uu)) and np.all(np.isfinite(vv)) and np.abs(uu).max() > 0.5       # a developed channel flow


def test_cached_factors_are_not_pinned_by_solutions():
    """a SteadySol must not keep the factored system alive -- clear_cache() returns the HBM even while
    solutions are still referenced, and reading .mat afterwards still works (it re-assembles from host descriptors)."""
    import torch
    cloud, solve = configs.config1(u, 40, 30)
    u.clear_cache(); torch.cuda.synchronize()
    base = torch.cuda.memory_allocated()
    sol = solve()
    held = torch.cuda.memory_allocated() - base
    n = cloud.N + 3
    assert held >= 8 * n * n                       # the factored K is cached
    u.clear_cache(); torch.cuda.synchronize()
    assert torch.cuda.memory_allocated() - base < 0.05 * held, "a live solution still pins the factors"
    B = sol.mat                                    # lazily: D inv(A)[:, :N]
    assert B.shape == (cloud.N, cloud.N) and np.all(np.isfinite(B))
    assert np.max(np.abs(B @ sol.vals - B @ sol.vals)) == 0.0
    u.clear_cache()


def test_user_rhs_that_rebuilds_coordinates():
    """an rhs operator that rebuilds x from its components (the port of jnp.array([x[0], x[1]])) must see the
    batched layout, not be reshaped silently into wrong evaluation points."""
    cloud = u.SquareCloud(Nx=12, Ny=10, facet_types=CONFIG1_FACETS, noise_key=1)
    rbf = partial(u.polyharmonic, a=1)
    fld = np.sin(2 * cloud.sorted_nodes[:, 0]) * cloud.sorted_nodes[:, 1]
    bcs = {k: (lambda c: 0.0) for k in cloud.facet_types}
    direct = lambda x, centers, rbf, fields: u.value(x, fields[:, 0], centers, rbf)
    rebuilt = lambda x, centers, rbf, fields: u.value(np.stack([x[0], x[1]]), fields[:, 0], centers, rbf)
    per_node = lambda x, centers, rbf, fields: (u.value(x, fields[:, 0], centers, rbf) if float(x[0]) >= -1.0 else 0.0)   # float(x[0]): rows only
    s1 = u.pde_solver_jit(laplace_op(u), direct, cloud, bcs, rbf, 1, rhs_args=[fld])
    s2 = u.pde_solver_jit(laplace_op(u), rebuilt, cloud, bcs, rbf, 1, rhs_args=[fld])
    s3 = u.pde_solver_jit(laplace_op(u), per_node, cloud, bcs, rbf, 1, rhs_args=[fld])
    assert np.array_equal(s1.vals, s2.vals)
    assert _rel(s3.vals, s1.vals) <= 1e-12


def test_reference_test_operators_on_its_own_mesh_gpu():
    """The reference's own test (updes/tests/test_operators.py:102-103) through the CUDA path, on the
    committed parse of its mesh.msh fixture: constant field, vanishing gradient and divergence."""
    cloud, _ = cloud_from_golden("mesh_msh_cloud_alln.npz")
    rbf = partial(u.gaussian, eps=10.0)
    op = lambda x, c, r, m, f: u.nodal_value(x, c, r, m)
    sol = u.pde_solver(op, lambda x, centers, rbf, fields: 12.0, cloud, {k: (lambda c: 0.0) for k in cloud.facet_types}, rbf, 1)
    grads = u.gradient_vec(cloud.sorted_nodes, sol.coeffs, cloud.sorted_nodes, rbf)
    divs = u.divergence_vec(cloud.sorted_nodes, np.stack([sol.coeffs, sol.coeffs], -1), cloud.sorted_nodes, rbf)
    assert np.allclose(np.linalg.norm(grads, axis=-1), 0, atol=1e-2) and np.allclose(divs, 0, atol=1e-2)


def test_differentiable_solve_wrt_boundary_data():
    """d(loss)/d(boundary array) through the factored system: K^-T via the transposed solve, checked
    against central finite differences (the DP demos' use case, demos/Laplace/10_...:90-104)."""
    import torch
    from updes_b200 import assembly as asm
    from updes_b200.autodiff import linear_solve
    from updes_b200.linalg import LUFactorization
    cloud = u.SquareCloud(Nx=14, Ny=12, facet_types=CONFIG1_FACETS, noise_key=4)
    M, n = 3, cloud.N + 3
    coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1))
    rows = asm.DeviceRows(cloud, asm.build_operator_rows(cloud, coef))
    K = asm.assemble_system(rows, "polyharmonic", 1.0, M)
    Kd = K[:, :n].clone()
    lu = LUFactorization(K, n).factor()
    north = torch.as_tensor(np.asarray(cloud.facet_nodes["North"])).cuda()
    w = torch.linspace(0.5, 1.5, n, dtype=torch.float64, device="cuda")

    def loss_of(bc):
        b = torch.zeros(n, dtype=torch.float64, device="cuda").index_put((north,), bc)
        c = linear_solve(lu, b)
        return (w * c * c).sum()

    bc = torch.sin(torch.linspace(0, 3.0, len(north), dtype=torch.float64, device="cuda")).requires_grad_(True)
    loss = loss_of(bc)
    loss.backward()
    g = bc.grad.clone()
    # reference gradient with dense algebra: dL/db = K^-T (2 w c), restricted to the North rows
    b = torch.zeros(n, dtype=torch.float64, device="cuda").index_put((north,), bc.detach())
    c = torch.linalg.solve(Kd, b)
    gref = torch.linalg.solve(Kd.T, 2 * w * c)[north]
    assert float((g - gref).abs().max() / gref.abs().max()) <= 1e-8
    # and one finite-difference probe
    e = torch.zeros_like(bc); e[3] = 1.0
    h = 1e-6
    fd = (loss_of(bc.detach() + h * e) - loss_of(bc.detach() - h * e)) / (2 * h)
    assert abs(float(fd) - float(g[3])) <= 1e-5 * max(1.0, abs(float(g[3])))
