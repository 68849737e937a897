"""The problems of tests/golden/make_reference_golden.py written against the PRODUCT's call surface (``lib`` =
updes_b200), plus what each lowers to.  The golden files hold what the reference's own code returned for the same
problems; CPU tests compare the oracle and the product's host layer with them, GPU tests the CUDA path."""
import os
from functools import partial

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_K(g):
    """[[opPhi opP], [bdPhi bdP], [P^T 0]] composed from the reference's blocks (the P^T rows are A's, assembly.py:80-83)."""
    N, M = g["opPhi"].shape[1], int(g["M"])
    top = np.concatenate([g["opPhi"], g["opP"]], axis=1)
    mid = np.concatenate([g["bdPhi"], g["bdP"]], axis=1)
    return np.concatenate([top, mid, g["A"][N:, :N + M]], axis=0)


def facet_dict(g):
    names = [str(v) for v in g["facet_names"]]
    offs = np.concatenate([[0], np.cumsum(g["facet_sizes"])])
    return {nm: g["facet_nodes"][offs[k]:offs[k + 1]].tolist() for k, nm in enumerate(names)}


def assert_cloud_equals_golden(cloud, g):
    """Every array the assembly reads from the cloud, bit for bit."""
    assert np.array_equal(np.asarray(cloud.sorted_nodes), g["sorted_nodes"])
    son = np.asarray(cloud.sorted_outward_normals, dtype=np.float64).reshape(-1, 2)
    assert np.array_equal(son, g["sorted_outward_normals"].reshape(-1, 2))
    assert [cloud.N, cloud.Ni, cloud.Nd, cloud.Nn, cloud.Nr] == g["counts"].tolist()
    assert list(cloud.Np) == g["Np"].tolist()
    want = facet_dict(g)
    assert list(cloud.facet_nodes.keys()) == list(want.keys())
    for k in want:
        assert list(cloud.facet_nodes[k]) == want[k], k
    assert [cloud.facet_types[k] for k in want] == [str(t) for t in g["facet_types"]]


# ---- the four solved problems -------------------------------------------------------------------------------------
class Case:
    pass


def laplace(lib, nx, ny):
    c = Case()
    c.cloud_args = dict(Nx=nx, Ny=ny, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})
    c.kind, c.param, c.max_degree = "polyharmonic", 1, 1
    c.rbf = lib.polyharmonic
    c.op = lambda x, center, rbf, monomial, fields: lib.nodal_laplacian(x, center, rbf, monomial)
    c.rhs = lambda x, centers, rbf, fields: 0.0
    c.bcs = {"South": lambda p: 0.0, "West": lambda p: 0.0, "North": lambda p: np.sin(np.pi * p[0]), "East": lambda p: 0.0}
    c.coef = lambda cloud: np.tile([0.0, 0.0, 0.0, 1.0, 1.0], (cloud.Ni, 1))
    c.diff_args = c.rhs_args = None
    return c


def robin(lib):
    c = Case()
    c.cloud_args = dict(Nx=11, Ny=8, facet_types={"South": "n", "West": "r", "North": "d", "East": "r"})
    c.kind, c.param, c.max_degree = "gaussian", 3.0, 2
    c.rbf = partial(lib.gaussian, eps=3.0)

    def op(x, center, rbf, monomial, fields):
        val = lib.nodal_value(x, center, rbf, monomial)
        grad = lib.nodal_gradient(x, center, rbf, monomial)
        lap = lib.nodal_laplacian(x, center, rbf, monomial)
        return 2.5 * val + lib.dot(np.array([1.5, -0.5]), grad) - 0.3 * lap
    c.op = op
    c.rhs = lambda x, centers, rbf, fields: np.cos(3.0 * x[0]) * x[1]
    c.bcs = {"South": lambda p: 0.25 * p[0], "West": (lambda p: 1.0 + p[1], lambda p: 2.0 + p[1]),
             "North": lambda p: np.sin(np.pi * p[0]), "East": (lambda p: -0.5, 0.75 * np.ones(6))}
    c.coef = lambda cloud: np.tile([2.5, 1.5, -0.5, -0.3, -0.3], (cloud.Ni, 1))
    c.diff_args = c.rhs_args = None
    return c


def periodic(lib, u0=None):
    DT, VEL, K = 1e-4, (100.0, 0.0), 0.08
    c = Case()
    c.cloud_args = dict(Nx=10, Ny=10, facet_types={"South": "p1", "North": "p1", "West": "p2", "East": "p2"})
    c.kind, c.param, c.max_degree = "polyharmonic", 1, 0
    c.rbf = partial(lib.polyharmonic, a=1)

    def op(x, center, rbf, monomial, fields):
        val = lib.nodal_value(x, center, rbf, monomial)
        grad = lib.nodal_gradient(x, center, rbf, monomial)
        lap = lib.nodal_laplacian(x, center, rbf, monomial)
        return (val / DT) + lib.dot(np.asarray(VEL), grad) - K * lap
    c.op = op
    c.rhs = lambda x, centers, rbf, fields: lib.value(x, fields[:, 0], centers, rbf) / DT
    c.bcs = {k: (lambda p: 0.0) for k in c.cloud_args["facet_types"]}
    c.coef = lambda cloud: np.tile([1.0 / DT, VEL[0], VEL[1], -K, -K], (cloud.Ni, 1))
    c.diff_args, c.rhs_args = None, (None if u0 is None else [u0])
    return c


def kernels_operator(lib):
    """Uses every term of the set with field-dependent coefficients (fields = f0, f1 of the golden file):
    f0 phi + f1 phi_x + 0.3 phi_y + (f0 - 0.7) phi_xx + (f1 - 0.7) phi_yy."""
    def op(x, center, rbf, monomial, fields):
        val = lib.nodal_value(x, center, rbf, monomial)
        grad = lib.nodal_gradient(x, center, rbf, monomial)
        lap = lib.nodal_laplacian(x, center, rbf, monomial)
        dg = lib.nodal_div_grad(x, center, rbf, monomial, (fields[0], fields[1]))
        return fields[0] * val + lib.dot([fields[1], 0.3], grad) - 0.7 * lap + dg
    return op


def kernels_coef(g, Ni):
    f0, f1 = g["f0"][:Ni], g["f1"][:Ni]
    return np.stack([f0, f1, np.full(Ni, 0.3), f0 - 0.7, f1 - 0.7], axis=1)


KERNELS_CLOUD = dict(Nx=7, Ny=6, facet_types={"South": "n", "West": "d", "North": "d", "East": "n"})


def kernel_rbf(lib, name, param):
    f = getattr(lib, name)
    return partial(f, a=int(param)) if name in ("polyharmonic", "thin_plate") else partial(f, eps=float(param))


def oracle_config3_loop(oracle, interpolate_field, bcs, cv, cp, nb_iter=2, Re=100.0, M=3):
    """simulate_forward_navier_stokes (demos/NavierStokes/30_...:97-213) restated with the oracle's reference
    formulation: inv(A) coefficients for the rhs fields, B = D inv(A), QR.  Returns [(u, v, p_)] after every iteration."""
    bc_u, bc_v, bc_phi = bcs
    Av, Ap = oracle.assemble_A(cv, "polyharmonic", 1, M), oracle.assemble_A(cp, "polyharmonic", 1, M)
    coefs_of = lambda A, f: np.linalg.solve(A, np.concatenate([f, np.zeros(M)]))
    ev = lambda cloud, c, which, pts=None: oracle.eval_field(cloud.sorted_nodes if pts is None else pts, cloud.sorted_nodes, c,
                                                            "polyharmonic", 1, which)
    ru, rv, rp_ = np.zeros(cv.N), np.zeros(cv.N), np.zeros(cp.N)
    lap_coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cp.Ni, 1))
    out = []
    for _ in range(nb_iter):
        p = interpolate_field(rp_, cp, cv)
        cpv = coefs_of(Av, p)
        coef = np.stack([np.zeros(cv.Ni), ru[:cv.Ni], rv[:cv.Ni], np.full(cv.Ni, -1 / Re), np.full(cv.Ni, -1 / Re)], axis=1)
        qu = oracle.assemble_q(cv, -ev(cv, cpv, "dx", cv.sorted_nodes[:cv.Ni]), bc_u)
        qv = oracle.assemble_q(cv, -ev(cv, cpv, "dy", cv.sorted_nodes[:cv.Ni]), bc_v)
        ustar, _, _ = oracle.reference_solve(cv, "polyharmonic", 1, 1, coef, qu)
        vstar, _, _ = oracle.reference_solve(cv, "polyharmonic", 1, 1, coef, qv)
        u_, v_ = interpolate_field(ustar, cv, cp), interpolate_field(vstar, cv, cp)
        cu_, cv_ = coefs_of(Ap, u_), coefs_of(Ap, v_)
        div = ev(cp, cu_, "dx", cp.sorted_nodes[:cp.Ni]) + ev(cp, cv_, "dy", cp.sorted_nodes[:cp.Ni])
        phi, cphi, _ = oracle.reference_solve(cp, "polyharmonic", 1, 1, lap_coef, oracle.assemble_q(cp, div, bc_phi))
        rp_ = rp_ + phi
        gradphi = interpolate_field(np.stack([ev(cp, cphi, "dx"), ev(cp, cphi, "dy")], axis=-1), cp, cv)
        ru, rv = ustar - gradphi[:, 0], vstar - gradphi[:, 1]
        out.append((ru, rv, rp_))
    return out


MULTI_CLOUD = dict(Nx=9, Ny=8, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})


def multi_problem(lib):
    """The coupled pair of tests/golden/make_reference_golden.py:case_multi on the product's surface."""
    zero, one = (lambda c: 0.0), (lambda c: 1.0)
    bcs = [{"South": zero, "West": zero, "North": one, "East": zero}, {"South": zero, "West": one, "North": zero, "East": zero}]
    op0 = lambda x, c, r, m, f: lib.nodal_laplacian(x, c, r, m) - (1.0 + f[1] ** 2) * lib.nodal_value(x, c, r, m)
    op1 = lambda x, c, r, m, f: lib.nodal_laplacian(x, c, r, m) + (x[0] + f[0]) * lib.nodal_gradient(x, c, r, m)[0]
    rhs0 = lambda x, centers, rbf, fields: 0.0
    rhs1 = lambda x, centers, rbf, fields: -1.0
    return [op0, op1], [rhs0, rhs1], bcs, partial(lib.polyharmonic, a=1)


CLOUD_KEYS = ("sorted_nodes", "sorted_outward_normals", "counts", "Np", "facet_names", "facet_sizes", "facet_nodes", "facet_types")


def fuzz_cases():
    """(k, Nx, Ny, facet dict in the generator's order, kernel, param, M, fields (5, N), betas or None, golden sub-dict)
    for every random problem of tests/golden/ref_fuzz_16.npz."""
    g = load("ref_fuzz_16")
    for k in range(int(g["ncases"])):
        pre = "c%02d_" % k
        nx, ny, deg = (int(v) for v in g[pre + "config"])
        facets = {str(f): str(t) for f, t in g[pre + "facets_in"]}
        sub = {key: g[pre + key] for key in CLOUD_KEYS}
        sub["diffMat"] = g[pre + "diffMat"]
        betas = g[pre + "betas"] if len(g[pre + "betas"]) else None
        import math
        yield k, nx, ny, facets, str(g[pre + "kernel"][0]), float(g[pre + "kernel"][1]), math.comb(deg + 2, deg), g[pre + "fields"], betas, sub
