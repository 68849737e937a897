"""BASELINE.json configs 1-3 written against the updes_b200 call surface, the way the reference's own
scripts define them.  Shared by tests/ (parity against the oracle) and bench.py (`small_configs`).

  config 1  README Laplace example, SquareCloud 30x20, polyharmonic a=1, degree 1        README.md:37-68
  config 2  periodic advection-diffusion time stepping, factor once / solve per step       demos/Advection/01_adv_diff_periodic.py:34-113
  config 3  Navier-Stokes channel projection loop on two complementary GMSH clouds         demos/NavierStokes/30_channel_flow_blowing_suction.py:40-213
"""
from functools import partial

import numpy as np

CONFIG1_FACETS = {"South": "n", "West": "d", "North": "d", "East": "d"}
CONFIG2_FACETS = {"South": "p1", "North": "p1", "West": "p2", "East": "p2"}
FACETS_VEL = {"Wall": "d", "Inflow": "d", "Outflow": "n", "Blowing": "d", "Suction": "d"}
FACETS_PHI = {"Wall": "n", "Inflow": "n", "Outflow": "d", "Blowing": "n", "Suction": "n"}


# ---- config 1 ------------------------------------------------------------------------------------
def config1(u, Nx=30, Ny=20):
    cloud = u.SquareCloud(Nx=Nx, Ny=Ny, facet_types=CONFIG1_FACETS)
    op = lambda x, center, rbf, monomial, fields: u.nodal_laplacian(x, center, rbf, monomial)
    rhs = lambda x, centers, rbf, fields: 0.0
    bcs = {"South": lambda c: 0.0, "West": lambda c: 0.0, "North": lambda c: np.sin(np.pi * c[0]), "East": lambda c: 0.0}
    solve = lambda: u.pde_solver_jit(diff_operator=op, rhs_operator=rhs, cloud=cloud, boundary_conditions=bcs,
                                     rbf=u.polyharmonic, max_degree=1)
    return cloud, solve


# ---- config 2 ------------------------------------------------------------------------------------
def config2(u, Nx=35, Ny=35, DT=1e-4, VEL=(100.0, 0.0), K=0.08, noise_key=11):
    cloud = u.SquareCloud(Nx=Nx, Ny=Ny, facet_types=CONFIG2_FACETS, noise_key=noise_key)
    rbf = partial(u.polyharmonic, a=1)

    def diff_operator(x, center, rbf, monomial, fields):
        val = u.nodal_value(x, center, rbf, monomial)
        grad = u.nodal_gradient(x, center, rbf, monomial)
        lap = u.nodal_laplacian(x, center, rbf, monomial)
        return (val / DT) + u.dot(np.asarray(VEL), grad) - K * lap

    def rhs_operator(x, centers, rbf, fields):
        return u.value(x, fields[:, 0], centers, rbf) / DT

    xy = cloud.sorted_nodes
    u0 = np.exp(-((xy[:, 0] - 0.35) ** 2 + (xy[:, 1] - 0.5) ** 2) / (2 * 0.1 ** 2))
    bcs = {k: (lambda c: 0.0) for k in CONFIG2_FACETS}
    step = lambda uprev: u.pde_solver_jit(diff_operator=diff_operator, rhs_operator=rhs_operator, rhs_args=[uprev],
                                          cloud=cloud, boundary_conditions=bcs, rbf=rbf, max_degree=0)
    coef = np.tile([1.0 / DT, VEL[0], VEL[1], -K, -K], (cloud.Ni, 1))
    return cloud, u0, step, coef


# ---- config 3 ------------------------------------------------------------------------------------
def config3_operators(u, Re=100.0):
    def diff_operator_u(x, center=None, rbf=None, monomial=None, fields=None):
        U_prev = np.array([fields[0], fields[1]])
        u_grad = u.nodal_gradient(x, center, rbf, monomial)
        u_lap = u.nodal_laplacian(x, center, rbf, monomial)
        return u.dot(U_prev, u_grad) - u_lap / Re

    def rhs_operator_u(x, centers=None, rbf=None, fields=None):
        return -u.gradient(x, fields[:, 0], centers, rbf)[0]

    def rhs_operator_v(x, centers=None, rbf=None, fields=None):
        return -u.gradient(x, fields[:, 0], centers, rbf)[1]

    def diff_operator_phi(x, center=None, rbf=None, monomial=None, fields=None):
        return u.nodal_laplacian(x, center, rbf, monomial)

    def rhs_operator_phi(x, centers=None, rbf=None, fields=None):
        return u.divergence(x, fields[:, :2], centers, rbf)

    return diff_operator_u, rhs_operator_u, diff_operator_u, rhs_operator_v, diff_operator_phi, rhs_operator_phi


def config3_boundary_arrays(cloud_vel, cloud_phi):
    xy = cloud_vel.sorted_nodes
    z = lambda cloud, f: np.zeros(len(cloud.facet_nodes[f]))
    inflow = np.asarray(cloud_vel.facet_nodes["Inflow"], dtype=int)
    bc_u = {f: z(cloud_vel, f) for f in FACETS_VEL}
    bc_u["Inflow"] = 4.0 * xy[inflow, 1] * (1.0 - xy[inflow, 1])                  # parabolic profile, :118
    bc_v = {f: z(cloud_vel, f) for f in FACETS_VEL}
    bc_v["Blowing"] = np.full(len(cloud_vel.facet_nodes["Blowing"]), 0.3)
    bc_v["Suction"] = np.full(len(cloud_vel.facet_nodes["Suction"]), 0.3)
    bc_phi = {f: z(cloud_phi, f) for f in FACETS_PHI}
    return bc_u, bc_v, bc_phi


def config3_projection_loop(u, cloud_vel, cloud_phi, nb_iter=2, Re=100.0, Pa=0.0, max_degree=1):
    """simulate_forward_navier_stokes (demos/NavierStokes/30_...:97-213): per iteration a u-solve and a v-solve on
    cloud_vel (matrix depends on the previous velocity), a pressure-correction Poisson solve on cloud_phi
    (constant matrix: factored once), fields carried between the clouds with interpolate_field."""
    rbf = u.polyharmonic
    du, ru, dv, rv, dphi, rphi = config3_operators(u, Re)
    bc_u, bc_v, bc_phi = config3_boundary_arrays(cloud_vel, cloud_phi)
    uu = np.zeros(cloud_vel.N); vv = np.zeros(cloud_vel.N)
    p_ = np.zeros(cloud_phi.N)
    p_[np.asarray(cloud_phi.facet_nodes["Outflow"], dtype=int)] = Pa
    history = []
    for _ in range(nb_iter):
        p = u.interpolate_field(p_, cloud_phi, cloud_vel)
        usol = u.pde_solver_jit_with_bc(diff_operator=du, diff_args=[uu, vv], rhs_operator=ru, rhs_args=[p], cloud=cloud_vel,
                                        boundary_conditions=bc_u, rbf=rbf, max_degree=max_degree)
        vsol = u.pde_solver_jit_with_bc(diff_operator=dv, diff_args=[uu, vv], rhs_operator=rv, rhs_args=[p], cloud=cloud_vel,
                                        boundary_conditions=bc_v, rbf=rbf, max_degree=max_degree)
        ustar, vstar = usol.vals, vsol.vals
        u_ = u.interpolate_field(ustar, cloud_vel, cloud_phi)
        v_ = u.interpolate_field(vstar, cloud_vel, cloud_phi)
        phisol_ = u.pde_solver_jit_with_bc(diff_operator=dphi, rhs_operator=rphi, rhs_args=[u_, v_], cloud=cloud_phi,
                                           boundary_conditions=bc_phi, rbf=rbf, max_degree=max_degree)
        p_ = p_ + phisol_.vals
        gradphi_ = u.gradient_vec(cloud_phi.sorted_nodes, phisol_.coeffs, cloud_phi.sorted_nodes, rbf)
        gradphi = u.interpolate_field(gradphi_, cloud_phi, cloud_vel)
        U = np.stack([ustar, vstar], axis=-1) - gradphi
        uu, vv = U[:, 0], U[:, 1]
        history.append((ustar, vstar, phisol_.vals, uu, vv, p_))
    return uu, vv, p_, history
