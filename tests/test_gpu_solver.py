"""GPU parity of the whole path: pde_solver vs the reference formulation (inv + GEMM + QR) on the CPU."""
from functools import partial

import numpy as np
import pytest

import updes_b200 as u
from helpers import CONFIG1_FACETS, CONFIG2_FACETS, advdiff_op, cloud_from_golden, laplace_op

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


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
    uu, ref = u0.copy(), u0.copy()
    launches0 = None
    for step in range(3):
        sol = u.pde_solver_jit(diff_operator=advdiff_op(u, DT), rhs_operator=rhs, rhs_args=[uu], cloud=cloud,
                               boundary_conditions=bcs, rbf=rbf, max_degree=0)
        uu = sol.vals
        # reference formulation on the CPU: coefficients of the previous field, value(x)/DT on internal nodes
        A = oracle.assemble_A(cloud, "polyharmonic", 1, 1)
        cprev = np.linalg.solve(A, np.concatenate([ref, np.zeros(1)]))
        q_int = oracle.eval_field(cloud.sorted_nodes[:cloud.Ni], cloud.sorted_nodes, cprev, "polyharmonic", 1, "value") / DT
        q = oracle.assemble_q(cloud, q_int, {k: np.zeros(len(cloud.facet_nodes[k])) for k in cloud.facet_types})
        ref, _, _ = oracle.reference_solve(cloud, "polyharmonic", 1, 0, coef, q)
        assert _rel(uu, ref) <= 1e-6, (step, _rel(uu, ref))


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
    assert _rel(sol.vals, vals) <= 1e-7, _rel(sol.vals, vals)
    # and the system is actually solved: residual of K c = [q;0] at round-off level
    K = oracle.assemble_K(cloud, "polyharmonic", 1, 3, coef)
    r = K @ sol.coeffs - np.concatenate([q, np.zeros(3)])
    assert np.max(np.abs(r)) <= 1e-13 * (np.abs(K).sum(1).max() * np.abs(sol.coeffs).max() + np.abs(q).max())


def test_pde_multi_solver_picard_two_fields():
    """pde_multi_solver (operators.py:696-771): two decoupled-in-the-limit equations converge to the
    single-equation solutions after a few sweeps."""
    cloud = u.SquareCloud(Nx=16, Ny=14, facet_types=CONFIG1_FACETS)
    rbf = partial(u.polyharmonic, a=1)
    zero = lambda c: 0.0
    one = lambda c: 1.0
    bcs = [{"South": zero, "West": zero, "North": one, "East": zero}, {"South": zero, "West": one, "North": zero, "East": zero}]
    # eq. i:  lap(u_i) + 0 * (other field) u_i = 0  -- the coefficient depends on the other unknown through fields
    op0 = lambda x, c, r, m, f: u.nodal_laplacian(x, c, r, m) + (0.0 * f[1]) * u.nodal_value(x, c, r, m)
    op1 = lambda x, c, r, m, f: u.nodal_laplacian(x, c, r, m) + (0.0 * f[0]) * u.nodal_value(x, c, r, m)
    rhs = lambda x, centers, rbf, fields: 0.0
    u0 = [np.zeros(cloud.N), np.zeros(cloud.N)]
    sols = u.pde_multi_solver([op0, op1], [rhs, rhs], cloud, bcs, rbf, 1, nb_iters=2, diff_args=[u0, u0], rhs_args=None)
    ref0 = u.pde_solver_jit(laplace_op(u), rhs, cloud, bcs[0], rbf, 1)
    ref1 = u.pde_solver_jit(laplace_op(u), rhs, cloud, bcs[1], rbf, 1)
    assert _rel(sols[0].vals, ref0.vals) <= 1e-10 and _rel(sols[1].vals, ref1.vals) <= 1e-10


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
