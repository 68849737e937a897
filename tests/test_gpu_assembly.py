"""GPU parity: assembled collocation blocks vs the CPU oracle (through the C-ABI)."""
import numpy as np
import pytest

import updes_b200 as u
from updes_b200 import assembly as asm
from helpers import CONFIG1_FACETS, CONFIG2_FACETS, cloud_from_golden, rel_err_rowscaled, true_rel_err

pytestmark = pytest.mark.gpu

KERNELS = [("polyharmonic", 1), ("polyharmonic", 2), ("polyharmonic", 0), ("thin_plate", 1), ("thin_plate", 2),
           ("gaussian", 1.0), ("gaussian", 3.0), ("multiquadric", 1.0), ("multiquadric", 2.5),
           ("inverse_multiquadric", 1.0), ("inverse_multiquadric", 4.0)]


def _assemble_K(cloud, kind, param, M, coef, betas=None):
    table = asm.build_operator_rows(cloud, coef, None, betas)
    rows = asm.DeviceRows(cloud, table)
    K = asm.assemble_system(rows, kind, param, M)
    n = cloud.N + M
    Kh = K.cpu().numpy()
    assert np.all(Kh[:, n:] == 0.0), "padding columns must be zero"
    return Kh[:, :n]


@pytest.mark.parametrize("kind,param", KERNELS)
@pytest.mark.parametrize("degree", [0, 1, 4])
def test_config1_blocks_all_kernels(oracle, kind, param, degree):
    """Laplace-type + full jet operator on the README cloud (30x20, d/n facets), every kernel."""
    cloud = u.SquareCloud(Nx=30, Ny=20, facet_types=CONFIG1_FACETS)
    M = u.compute_nb_monomials(degree, 2)
    rng = np.random.default_rng(0)
    coef = rng.normal(size=(cloud.Ni, 5))                      # general operator, row-dependent coefficients
    got = _assemble_K(cloud, kind, param, M, coef)
    want = oracle.assemble_K(cloud, kind, param, M, coef)
    assert rel_err_rowscaled(got, want) <= 1e-12
    # second number: TRUE per-entry relative error on every entry above 1e-6 of its row's largest magnitude
    # (the row-scaled figure alone would let a tiny entry be 100 % wrong)
    t = true_rel_err(got, want)
    print("%s(%s) degree %d: row-scaled %.1e, true per-entry %.1e" % (kind, param, degree, rel_err_rowscaled(got, want), t))
    assert t <= 1e-10, t


@pytest.mark.parametrize("mask_cols", [[0], [1, 2], [3, 4], [0, 1, 2], [3], [0, 3, 4]])
def test_jet_mask_specialisations(oracle, mask_cols):
    """Each jet-mask specialisation of the kernel (only some coefficient columns non-zero)."""
    cloud = u.SquareCloud(Nx=17, Ny=13, facet_types=CONFIG1_FACETS, noise_key=5)
    coef = np.zeros((cloud.Ni, 5))
    coef[:, mask_cols] = np.random.default_rng(1).normal(size=(cloud.Ni, len(mask_cols)))
    for kind, param in [("polyharmonic", 1), ("gaussian", 2.0), ("thin_plate", 1)]:
        got = _assemble_K(cloud, kind, param, 3, coef)
        want = oracle.assemble_K(cloud, kind, param, 3, coef)
        assert rel_err_rowscaled(got, want) <= 1e-12


@pytest.mark.parametrize("kind,param", KERNELS)
def test_isotropic_laplacian_fast_path(oracle, kind, param):
    """c0*phi + c3*(phi_xx + phi_yy) rows take the closed-form radial-Laplacian kernel (Laplace, Helmholtz)."""
    cloud = u.SquareCloud(Nx=19, Ny=23, facet_types=CONFIG1_FACETS, noise_key=9)
    rng = np.random.default_rng(4)
    for with_val in (False, True):
        coef = np.zeros((cloud.Ni, 5))
        coef[:, 3] = coef[:, 4] = rng.normal(size=cloud.Ni)
        if with_val:
            coef[:, 0] = rng.normal(size=cloud.Ni)
        table = asm.build_operator_rows(cloud, coef)
        assert table.masks()[0] & asm.JET_ISO
        got = _assemble_K(cloud, kind, param, 3, coef)
        want = oracle.assemble_K(cloud, kind, param, 3, coef)
        assert rel_err_rowscaled(got, want) <= 1e-12


def test_config2_periodic_rows(oracle):
    """35x35 periodic cloud of demos/Advection/01 (two periodic groups), adv-diff coefficients."""
    cloud = u.SquareCloud(Nx=35, Ny=35, facet_types=CONFIG2_FACETS, noise_key=7)
    coef = np.tile(np.array([1e4, 100.0, 0.0, -0.08, -0.08]), (cloud.Ni, 1))
    for kind, param, deg in [("polyharmonic", 1, 0), ("gaussian", 5.0, 1)]:
        M = u.compute_nb_monomials(deg, 2)
        got = _assemble_K(cloud, kind, param, M, coef)
        want = oracle.assemble_K(cloud, kind, param, M, coef)
        assert rel_err_rowscaled(got, want) <= 1e-12


def test_robin_rows_with_normal_quirk(oracle):
    """Robin + Neumann facets together: bd(Phi) reads the reference's shifted normal slot (Q3)."""
    cloud = u.SquareCloud(Nx=12, Ny=10, facet_types={"South": "r", "West": "n", "North": "d", "East": "r"})
    betas = np.linspace(0.5, 2.0, cloud.Nr)
    coef = np.tile(np.array([0.0, 0.0, 0.0, 1.0, 1.0]), (cloud.Ni, 1))
    got = _assemble_K(cloud, "multiquadric", 2.0, 3, coef, betas)
    want = oracle.assemble_K(cloud, "multiquadric", 2.0, 3, coef, betas)
    assert rel_err_rowscaled(got, want) <= 1e-12


def test_interpolation_matrix_A(oracle):
    """A = [[Phi P],[P^T 0]] (assembly.py:62-85) incl. the zero diagonal of Phi (Q1)."""
    cloud = u.SquareCloud(Nx=12, Ny=12, facet_types={k: "d" for k in ("South", "West", "North", "East")})
    M = u.compute_nb_monomials(3, 2)
    rows = asm.DeviceRows(cloud, asm.build_interpolation_rows(cloud))
    got = asm.assemble_system(rows, "gaussian", 1.5, M).cpu().numpy()[:, :cloud.N + M]
    want = oracle.assemble_A(cloud, "gaussian", 1.5, M)
    assert np.all(np.diag(got)[:cloud.N] == 0.0)
    assert rel_err_rowscaled(got, want) <= 1e-12


def test_eval_jets_and_apply(oracle):
    """Matrix-free jets vs the oracle's field evaluators, and K @ c vs the assembled matrix."""
    import torch
    cloud = u.SquareCloud(Nx=21, Ny=16, facet_types=CONFIG2_FACETS, noise_key=2)
    N, M = cloud.N, 3
    rng = np.random.default_rng(3)
    coeffs = rng.normal(size=(N + M, 2))
    xs = rng.uniform(0, 1, size=(50, 2))
    xs[:10] = cloud.sorted_nodes[:10]                       # include r = 0 hits
    for kind, param in [("polyharmonic", 1), ("gaussian", 2.0), ("inverse_multiquadric", 1.0), ("thin_plate", 1)]:
        rbf = {"polyharmonic": u.polyharmonic, "gaussian": u.gaussian, "inverse_multiquadric": u.inverse_multiquadric,
               "thin_plate": u.thin_plate}[kind]
        from functools import partial
        rbf = partial(rbf, a=int(param)) if kind in ("polyharmonic", "thin_plate") else partial(rbf, eps=param)
        v = u.value_vec(xs, coeffs[:, 0], cloud.sorted_nodes, rbf)
        g = u.gradient_vec(xs, coeffs[:, 0], cloud.sorted_nodes, rbf)
        lap = u.laplacian_vec(xs, coeffs[:, 0], cloud.sorted_nodes, rbf)
        div = u.divergence_vec(xs, coeffs, cloud.sorted_nodes, rbf)
        ov = oracle.eval_field(xs, cloud.sorted_nodes, coeffs[:, 0], kind, param, "value")
        ox = oracle.eval_field(xs, cloud.sorted_nodes, coeffs[:, 0], kind, param, "dx")
        oy = oracle.eval_field(xs, cloud.sorted_nodes, coeffs[:, 0], kind, param, "dy")
        ol = oracle.eval_field(xs, cloud.sorted_nodes, coeffs[:, 0], kind, param, "laplacian")
        oy1 = oracle.eval_field(xs, cloud.sorted_nodes, coeffs[:, 1], kind, param, "dy")
        tol = lambda ref: 1e-11 * np.max(np.abs(ref))
        assert np.max(np.abs(v - ov)) <= tol(ov)
        assert np.max(np.abs(g[:, 0] - ox)) <= tol(ox) and np.max(np.abs(g[:, 1] - oy)) <= tol(oy)
        assert np.max(np.abs(lap - ol)) <= 1e-10 * np.max(np.abs(ol))
        assert np.max(np.abs(div - (ox + oy1))) <= tol(ox) + tol(oy1)
    # K @ c, matrix-free vs assembled
    coef = rng.normal(size=(cloud.Ni, 5))
    table = asm.build_operator_rows(cloud, coef)
    rows = asm.DeviceRows(cloud, table)
    K = asm.assemble_system(rows, "polyharmonic", 1, M)
    c = torch.as_tensor(coeffs.T.copy()).cuda()
    got = asm.apply_rows(rows, "polyharmonic", 1, M, c).cpu().numpy()
    want = (K[:, :N + M] @ c.T).T.cpu().numpy()
    assert np.max(np.abs(got - want)) <= 1e-11 * np.max(np.abs(want))


@pytest.mark.parametrize("tag", ["vel", "phi"])
def test_config3_gmsh_cloud_blocks(oracle, tag):
    """Config 3: the reference's mesh.msh cloud (N = 1385, approximate GMSH normals) at native size, with
    the Navier-Stokes momentum operator U.grad - lap/Re (fields = u, v; demos/NavierStokes/30_...:61-65)
    and the pressure-correction Laplacian (:89-90).  Cloud arrays come from tests/golden/."""
    cloud, _ = cloud_from_golden("mesh_msh_cloud_%s.npz" % tag)
    assert (cloud.N, cloud.Ni) == (1385, 1227)
    rng = np.random.default_rng(8)
    uu, vv = rng.normal(size=cloud.N), rng.normal(size=cloud.N)

    def ns(x, center, rbf, monomial, fields):
        U = np.array([fields[0], fields[1]])
        return u.dot(U, u.nodal_gradient(x, center, rbf, monomial)) - u.nodal_laplacian(x, center, rbf, monomial) / 100.0

    lap = lambda x, center, rbf, monomial, fields: u.nodal_laplacian(x, center, rbf, monomial)
    for op, args, kind, param, deg in [(ns, [uu, vv], "polyharmonic", 1, 1), (lap, None, "polyharmonic", 1, 1),
                                       (ns, [uu, vv], "thin_plate", 3, 4)]:
        from functools import partial
        rbf = partial(getattr(u, kind), a=param)
        coef, coef_p = u.lower_diff_operator(op, cloud, rbf, args)
        M = u.compute_nb_monomials(deg, 2)
        got = _assemble_K(cloud, kind, param, M, coef)
        want = oracle.assemble_K(cloud, kind, param, M, coef)
        assert rel_err_rowscaled(got, want) <= 1e-12


def test_golden_samples_through_the_cuda_path():
    """Committed golden K samples (configs 1 and 2) reproduced by the CUDA assembly."""
    for name, kind, coefrow, M in [("config1_30x20_phs3.npz", "polyharmonic", [0, 0, 0, 1.0, 1.0], 3),
                                   ("config2_35x35_periodic.npz", "polyharmonic", [1e4, 100.0, 0.0, -0.08, -0.08], 1)]:
        cloud, g = cloud_from_golden(name)
        K = _assemble_K(cloud, kind, 1, M, np.tile(coefrow, (cloud.Ni, 1)))
        sample = K[g["rows"]][:, g["cols"]]
        scale = np.max(np.abs(g["K_sample"]), axis=1, keepdims=True)
        assert np.max(np.abs(sample - g["K_sample"]) / scale) <= 1e-12
