"""Pin the CPU oracle (CPU only, no GPU): closed forms vs autodiff restatement, the reference's own
three known-answer tests restated through the oracle, and the analytic Laplace solution.

These pins predate tests/test_reference_golden.py, which checks the oracle against golden vectors produced by the
reference's own code (run over the JAX-API stand-in of oracle/refshim); they stay as independent second opinions."""
import os
from functools import partial

import numpy as np
import pytest

from helpers import CONFIG1_FACETS, CONFIG2_FACETS, cloud_from_golden, rel_err_rowscaled

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

KERNELS = [("polyharmonic", 1), ("polyharmonic", 2), ("polyharmonic", 0), ("thin_plate", 1), ("thin_plate", 3),
           ("gaussian", 1.0), ("gaussian", 10.0), ("multiquadric", 1.0), ("multiquadric", 3.5),
           ("inverse_multiquadric", 1.0), ("inverse_multiquadric", 2.0)]


@pytest.mark.parametrize("kind,param", KERNELS)
def test_closed_form_jets_match_autodiff(oracle, kind, param):
    """updes_oracle.c closed forms == torch.func autodiff of the kernel (mirrors jax.grad / jacfwd)."""
    from oracle import oracle_ad as AD
    rbf = AD.make_rbf(kind, param)
    rng = np.random.default_rng(0)
    for _ in range(40):
        x, c = rng.uniform(0, 1, 2), rng.uniform(0, 1, 2)
        a, b = oracle.rbf_jet(kind, param, x, c), AD.rbf_jet(rbf, x, c)
        assert np.max(np.abs(a - b)) <= 1e-12 * max(np.max(np.abs(b)), 1e-300)
    x = rng.uniform(0, 1, 2)
    a0, b0 = oracle.rbf_jet(kind, param, x, x), AD.rbf_jet(rbf, x, x)
    assert np.array_equal(a0, b0)            # r = 0: nan_to_num semantics, exact
    assert np.all(a0[1:] == 0.0)


def test_monomial_jets_match_autodiff(oracle):
    import torch
    from torch.func import grad, jacfwd
    from oracle import oracle_ad as AD
    x = np.array([0.37, 0.81])
    for j, mono in enumerate(AD.make_all_monomials(15)):
        xt = torch.as_tensor(x)
        g = grad(mono)(xt); H = jacfwd(grad(mono))(xt)
        want = np.array([float(mono(xt)), float(g[0]), float(g[1]), float(H[0, 0]), float(H[1, 1])])
        assert np.allclose(oracle.monomial_jet(j, x), want, rtol=1e-13, atol=1e-15)


def test_operator_rows_match_autodiff_operator(oracle):
    """assemble_op_Phi_P with lowered coefficients == the same operator run through autodiff."""
    from oracle import oracle_ad as AD
    import torch
    cloud = oracle.RefSquareCloud(7, 6, CONFIG1_FACETS, noise_seed=1)
    fld = np.linspace(1.0, 2.0, cloud.N)

    def op(x, center, rbf, monomial, fields):
        val = AD.nodal_value(x, center, rbf, monomial)
        grad = AD.nodal_gradient(x, center, rbf, monomial)
        lap = AD.nodal_laplacian(x, center, rbf, monomial)
        dg = AD.nodal_div_grad(x, center, rbf, monomial, (fields[0], 2.0 * fields[0]))
        return val / 0.5 + torch.dot(torch.tensor([3.0, -1.0]), grad) - 0.08 * lap + dg

    coef = np.zeros((cloud.Ni, 5))
    coef[:, 0] = 2.0; coef[:, 1] = 3.0; coef[:, 2] = -1.0
    coef[:, 3] = -0.08 + fld[:cloud.Ni]; coef[:, 4] = -0.08 + 2.0 * fld[:cloud.Ni]
    for kind, param in [("polyharmonic", 1), ("gaussian", 2.0), ("multiquadric", 1.0)]:
        a_phi, a_p = AD.assemble_op_Phi_P(op, cloud, AD.make_rbf(kind, param), 6, [fld])
        c_phi, c_p = oracle.assemble_op_Phi_P(cloud, kind, param, 6, coef)
        assert rel_err_rowscaled(c_phi, a_phi) <= 1e-12
        assert rel_err_rowscaled(c_p, a_p) <= 1e-12
        assert rel_err_rowscaled(oracle.assemble_Phi(cloud, kind, param), AD.assemble_Phi(cloud, AD.make_rbf(kind, param))) <= 1e-13


def test_reference_test_integrals_known_answer(oracle):
    """updes/tests/test_integrals.py:83-84 restated: 12x12 all-Dirichlet cloud, polyharmonic a=5,
    degree 3; coefficients = inv(A)[f;0]; integral of x^2/(1+y^2) over the unit square ~ pi/12 (1e-1)."""
    cloud = oracle.RefSquareCloud(12, 12, {k: "d" for k in ("South", "West", "North", "East")})
    M = oracle.compute_nb_monomials(3)
    xy = cloud.sorted_nodes
    f = xy[:, 0] ** 2 / (1 + xy[:, 1] ** 2)
    A = oracle.assemble_A(cloud, "polyharmonic", 5, M)
    coeffs = np.linalg.inv(A) @ np.concatenate([f, np.zeros(M)])
    # interpolant must reproduce the nodal values (A c = [f;0]) ...
    back = oracle.eval_field(xy, xy, coeffs, "polyharmonic", 5, "value")
    assert np.max(np.abs(back - f)) <= 1e-4      # r^11 is ill-conditioned: inv(A) loses ~10 digits
    # ... and integrate to pi/12 (midpoint rule on a fine grid of the interpolant)
    g = (np.arange(60) + 0.5) / 60
    X, Y = np.meshgrid(g, g)
    pts = np.stack([X.ravel(), Y.ravel()], 1)
    integral = oracle.eval_field(pts, xy, coeffs, "polyharmonic", 5, "value").mean()
    assert abs(integral - np.pi / 12) < 1e-1


def test_reference_test_interpolation_known_answer(oracle):
    """updes/tests/test_interpolation.py:60-63 restated: two 8x8 clouds with swapped BC types share
    the same node set; mapping through the renumbering maps is the identity on coordinates."""
    a = oracle.RefSquareCloud(8, 8, {"South": "d", "West": "d", "North": "n", "East": "n"})
    b = oracle.RefSquareCloud(8, 8, {"South": "n", "West": "n", "North": "d", "East": "d"})
    for old in range(64):
        assert np.allclose(a.sorted_nodes[a.renumbering_map[old]], b.sorted_nodes[b.renumbering_map[old]], atol=1e-12)
    assert a.Ni == b.Ni == 36


def test_reference_test_operators_known_answer(oracle):
    """updes/tests/test_operators.py:102-103 restated on a small all-Neumann cloud: gaussian eps=10,
    degree 1, diff = nodal_value, rhs = 12 -> constant field with ~zero gradient (atol 1e-2)."""
    cloud = oracle.RefSquareCloud(14, 14, {k: "n" for k in ("South", "West", "North", "East")}, noise_seed=12)
    coef = np.tile([1.0, 0, 0, 0, 0], (cloud.Ni, 1))
    q = oracle.assemble_q(cloud, np.full(cloud.Ni, 12.0), {k: np.zeros(len(v)) for k, v in cloud.facet_nodes.items()})
    vals, coeffs, _ = oracle.reference_solve(cloud, "gaussian", 10.0, 1, coef, q)
    gx = oracle.eval_field(cloud.sorted_nodes[:cloud.Ni], cloud.sorted_nodes, coeffs, "gaussian", 10.0, "dx")
    gy = oracle.eval_field(cloud.sorted_nodes[:cloud.Ni], cloud.sorted_nodes, coeffs, "gaussian", 10.0, "dy")
    assert np.allclose(vals[:cloud.Ni], 12.0, atol=1e-6)
    assert np.allclose(np.hypot(gx, gy), 0, atol=1e-2)


def test_reference_test_operators_on_its_own_mesh(oracle):
    """updes/tests/test_operators.py verbatim: GmshCloud(mesh.msh) with every facet Neumann, gaussian
    eps=10, max_degree=1, diff = nodal_value, rhs = 12, zero Neumann data; assert |grad u| ~ 0 and
    div(u, u) ~ 0 with atol 1e-2 (lines 102-103).  The cloud is the committed parse of that fixture."""
    cloud, _ = cloud_from_golden("mesh_msh_cloud_alln.npz")
    assert (cloud.N, cloud.Nd, cloud.Nn) == (1385, 0, 158)
    coef = np.tile([1.0, 0, 0, 0, 0], (cloud.Ni, 1))
    q = oracle.assemble_q(cloud, np.full(cloud.Ni, 12.0), {k: np.zeros(len(v)) for k, v in cloud.facet_nodes.items()})
    vals, coeffs, _ = oracle.reference_solve(cloud, "gaussian", 10.0, 1, coef, q)
    xy = cloud.sorted_nodes
    gx = oracle.eval_field(xy, xy, coeffs, "gaussian", 10.0, "dx")
    gy = oracle.eval_field(xy, xy, coeffs, "gaussian", 10.0, "dy")
    assert np.allclose(np.hypot(gx, gy), 0, atol=1e-2)          # test_operators.py:103, first clause
    assert np.allclose(gx + gy, 0, atol=1e-2)                   # divergence of (u, u), second clause
    assert np.allclose(vals[:cloud.Ni], 12.0, atol=1e-6)


def test_laplace_analytic_solution(oracle):
    """Analytic answer of demos/Laplace/00_laplace_with_rbf.py:109-110 through the reference formulation,
    and the LU reformulation K c = [q;0], u = [Phi P] c gives the same vals / coeffs."""
    cloud = oracle.RefSquareCloud(30, 20, CONFIG1_FACETS)
    coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1))
    xy = cloud.sorted_nodes
    bc = {f: (np.sin(np.pi * xy[ids, 0]) if f == "North" else np.zeros(len(ids))) for f, ids in cloud.facet_nodes.items()}
    q = oracle.assemble_q(cloud, np.zeros(cloud.Ni), bc)
    vals, coeffs, _ = oracle.reference_solve(cloud, "polyharmonic", 1, 1, coef, q)
    exact = np.sin(np.pi * xy[:, 0]) * np.cosh(np.pi * xy[:, 1]) / np.cosh(np.pi)
    assert np.max(np.abs(vals - exact)) <= 2e-2
    K = oracle.assemble_K(cloud, "polyharmonic", 1, 3, coef)
    c = np.linalg.solve(K, np.concatenate([q, np.zeros(3)]))
    u = oracle.assemble_A(cloud, "polyharmonic", 1, 3)[:cloud.N] @ c
    assert np.max(np.abs(u - vals)) <= 1e-8 * np.max(np.abs(vals))


def test_golden_fixture_matches_oracle(oracle):
    """Committed golden vectors (tests/golden/make_golden.py) still equal what the oracle computes."""
    g = np.load(os.path.join(GOLDEN, "config1_30x20_phs3.npz"))
    cloud = oracle.RefSquareCloud(30, 20, CONFIG1_FACETS)
    coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1))
    K = oracle.assemble_K(cloud, "polyharmonic", 1, 3, coef)
    assert np.array_equal(cloud.sorted_nodes, g["sorted_nodes"])
    assert np.max(np.abs(K[g["rows"]][:, g["cols"]] - g["K_sample"])) <= 1e-13 * np.max(np.abs(g["K_sample"]))


def test_golden_solution_vectors_match_reference_formulation(oracle):
    """The committed q / vals / coeffs of config 1 are what the reference formulation (inv + GEMM + QR)
    yields today, and they solve the reformulated system K c = [q; 0], vals = [Phi P] c."""
    g = np.load(os.path.join(GOLDEN, "config1_30x20_phs3.npz"))
    cloud = oracle.RefSquareCloud(30, 20, CONFIG1_FACETS)
    coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1))
    vals, coeffs, _ = oracle.reference_solve(cloud, "polyharmonic", 1, 1, coef, g["q"])
    assert np.max(np.abs(vals - g["vals"])) <= 1e-9 * np.max(np.abs(g["vals"]))
    K = oracle.assemble_K(cloud, "polyharmonic", 1, 3, coef)
    c = np.linalg.solve(K, np.concatenate([g["q"], np.zeros(3)]))
    A = oracle.assemble_A(cloud, "polyharmonic", 1, 3)
    assert np.max(np.abs(A[:cloud.N] @ c - g["vals"])) <= 1e-8 * np.max(np.abs(g["vals"]))


# ---- second, independent pin: 50-digit numerical differentiation of the kernels' DEFINITIONS -----
def _mp_kernel(kind, param):
    """The reference's kernel definitions (updes/utils.py:19-69) transcribed for mpmath: functions of (x, y; cx, cy)."""
    import mpmath as mp
    r = lambda x, y, cx, cy: mp.sqrt((x - cx) ** 2 + (y - cy) ** 2)            # utils.py:19-22
    if kind == "polyharmonic":
        return lambda x, y, cx, cy: r(x, y, cx, cy) ** (2 * int(param) + 1)     # utils.py:50-55
    if kind == "thin_plate":
        return lambda x, y, cx, cy: mp.log(r(x, y, cx, cy)) * r(x, y, cx, cy) ** (2 * int(param))   # utils.py:63-69
    if kind == "gaussian":
        return lambda x, y, cx, cy: mp.exp(-(mp.mpf(param) * r(x, y, cx, cy)) ** 2)     # utils.py:44-48
    if kind == "multiquadric":
        return lambda x, y, cx, cy: mp.sqrt(1 + (mp.mpf(param) * r(x, y, cx, cy)) ** 2)   # utils.py:30-35
    return lambda x, y, cx, cy: 1 / mp.sqrt(1 + (mp.mpf(param) * r(x, y, cx, cy)) ** 2)   # utils.py:37-42


def _mp_jet(kind, param, x, c):
    """(phi, phi_x, phi_y, phi_xx, phi_yy) w.r.t. the evaluation point by 50-digit numerical differentiation."""
    import mpmath as mp
    mp.mp.dps = 50
    f = _mp_kernel(kind, param)
    X, Y, CX, CY = (mp.mpf(float(v)) for v in (x[0], x[1], c[0], c[1]))
    fx = lambda a: f(a, Y, CX, CY)
    fy = lambda b: f(X, b, CX, CY)
    return [f(X, Y, CX, CY), mp.diff(fx, X, 1), mp.diff(fy, Y, 1), mp.diff(fx, X, 2), mp.diff(fy, Y, 2)]


@pytest.mark.parametrize("kind,param", KERNELS)
def test_closed_form_jets_match_50_digit_differentiation(oracle, kind, param):
    """The C closed forms against an arithmetic path that shares nothing with them: mpmath (50 digits) numerical
    differentiation of the kernel definition.  TRUE per-entry relative error <= 1e-12 wherever the entry is
    not a near-cancellation; points close to the sign change of the gaussian Laplacian terms (r ~ 1/eps) and very
    small / large r are included on purpose, there the error is measured against the jet's largest component."""
    rng = np.random.default_rng(11)
    pts = [(rng.uniform(0, 1, 2), rng.uniform(0, 1, 2)) for _ in range(12)]
    pts.append((np.array([0.3, 0.4]), np.array([0.3 + 1e-7, 0.4 - 2e-7])))              # tiny r
    pts.append((np.array([0.0, 0.0]), np.array([1.0, 1.0])))                            # largest r on the unit square
    if kind in ("gaussian", "multiquadric", "inverse_multiquadric"):
        rr = 1.0 / float(param)                                                           # phi_xx of the gaussian changes sign near dx = 1/(sqrt(2) eps)
        pts.append((np.array([0.2, 0.2]), np.array([0.2 + rr / np.sqrt(2.0), 0.2])))
        pts.append((np.array([0.2, 0.2]), np.array([0.2 + rr / np.sqrt(2.0) * (1 + 1e-6), 0.2 + 1e-9])))
    worst_true, worst_scaled = 0.0, 0.0
    for x, c in pts:
        got = oracle.rbf_jet(kind, param, x, c)
        want = [float(v) for v in _mp_jet(kind, param, x, c)]
        scale = max(abs(v) for v in want)
        for g, w in zip(got, want):
            worst_scaled = max(worst_scaled, abs(g - w) / scale)
            if abs(w) > 1e-6 * scale:
                worst_true = max(worst_true, abs(g - w) / abs(w))
    assert worst_scaled <= 1e-13, (kind, param, worst_scaled)
    assert worst_true <= 2e-10, (kind, param, worst_true)      # entries down to 1e-6 of the jet scale keep >= 4 extra digits


def test_op_rows_equals_the_rows_of_the_full_block(oracle):
    """uo_op_rows (row-list form used by the full-size GPU parity test) == the same rows of uo_assemble_op_Phi_P."""
    import updes_b200 as u
    cloud = u.SquareCloud(Nx=9, Ny=7, facet_types={"South": "n", "West": "d", "North": "r", "East": "d"})
    rng = np.random.default_rng(5)
    coef = rng.normal(size=(cloud.Ni, 5))
    for kind, param in (("polyharmonic", 1.0), ("gaussian", 2.0), ("thin_plate", 2.0)):
        full_phi, full_p = oracle.assemble_op_Phi_P(cloud, kind, param, 6, coef)
        pick = np.array([0, 3, cloud.Ni - 1, 17], dtype=np.int32)
        phi, p = oracle.op_rows(cloud, kind, param, 6, pick, coef[pick])
        assert np.array_equal(phi, full_phi[pick]) and np.array_equal(p, full_p[pick])
