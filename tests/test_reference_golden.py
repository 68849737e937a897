"""CPU tests against golden vectors produced by the REFERENCE'S OWN CODE (tests/golden/ref_*.npz, written by
tests/golden/make_reference_golden.py: the unmodified /root/reference package executed over the JAX-API stand-in of
oracle/refshim/).  They pin (1) the oracle -- clouds bit for bit, every block of diffMat and A to 1e-12 per entry, the
solve -- and (2) the product's host layer: clouds, operator lowering, boundary-condition preparation, right-hand side.
The GPU counterpart (the CUDA path against the same files) is tests/test_gpu_zy_reference_golden.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

import updes_b200 as u
import reference_cases as rc
from helpers import exact_solution, rel_err_rowscaled, true_rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"


def _check_blocks(oracle, cloud, g, kind, param, coef, betas=None, prefix=""):
    M = g[prefix + "opP"].shape[1]
    A = oracle.assemble_A(cloud, kind, param, M)
    opPhi, opP = oracle.assemble_op_Phi_P(cloud, kind, param, M, coef)
    bdPhi, bdP = oracle.assemble_bd_Phi_P(cloud, kind, param, M, betas)
    worst = 0.0
    for name, got in (("A", A), ("opPhi", opPhi), ("opP", opP), ("bdPhi", bdPhi), ("bdP", bdP)):
        want = g[prefix + name]
        assert got.shape == want.shape, name
        e, t = rel_err_rowscaled(got, want), true_rel_err(got, want)
        assert e <= 1e-12, (name, e)           # north_star: 1e-12 relative per entry (row-scale floor)
        assert t <= 1e-10, (name, t)           # true per-entry relative error where no cancellation
        worst = max(worst, e)
    return worst


def _solution_close(vals, g, K, rhs, A_rows, what):
    """Two inverse-based pipelines (the reference's over torch LAPACK, the oracle's over scipy LAPACK) agree to the
    accuracy either has: 1e-8, or 4x the golden solution's own distance from the exactly solved discrete system."""
    exact, _ = exact_solution(K, rhs, A_rows)
    scale = np.max(np.abs(exact))
    e_gold = np.max(np.abs(g["vals"] - exact)) / scale
    d = np.max(np.abs(vals - g["vals"])) / scale
    print("%s: golden-vs-exact %.2e, ours-vs-golden %.2e" % (what, e_gold, d))
    assert d <= max(1e-8, 4.0 * e_gold), (what, d, e_gold)


@pytest.mark.parametrize("name,nx,ny", [("ref_laplace_12x9", 12, 9)])
def test_oracle_laplace_blocks_and_solution(oracle, name, nx, ny):
    g = rc.load(name)
    case = rc.laplace(u, nx, ny)
    cloud = oracle.RefSquareCloud(nx, ny, case.cloud_args["facet_types"])
    rc.assert_cloud_equals_golden(cloud, g)
    coef = case.coef(cloud)
    print("worst block error %.1e" % _check_blocks(oracle, cloud, g, case.kind, case.param, coef))
    xy = cloud.sorted_nodes
    bc = {f: (np.sin(np.pi * xy[ids, 0]) if f == "North" else np.zeros(len(ids))) for f, ids in cloud.facet_nodes.items()}
    q = oracle.assemble_q(cloud, np.zeros(cloud.Ni), bc)
    assert np.array_equal(q, g["q"])
    vals, coeffs, B = oracle.reference_solve(cloud, case.kind, case.param, case.max_degree, coef, q)
    assert np.max(np.abs(B - g["B"])) <= 1e-9 * np.max(np.abs(g["B"]))            # B = diffMat inv(A)[:, :N] (assembly.py:396-401)
    K = rc.golden_K(g)
    _solution_close(vals, g, K, np.concatenate([q, np.zeros(3)]), g["A"][:cloud.N], name)


def test_oracle_robin_with_neumann_quirk_q3(oracle):
    """Neumann and Robin facets together: the reference's bd(Phi) reads normals[i-Ni-Nd-Nn] (assembly.py:206) -- the
    golden rows carry that quirk because the reference's code produced them."""
    g = rc.load("ref_robin_11x8")
    case = rc.robin(u)
    cloud = oracle.RefSquareCloud(11, 8, case.cloud_args["facet_types"])
    rc.assert_cloud_equals_golden(cloud, g)
    assert cloud.Nn > 0 and cloud.Nr > 0
    _check_blocks(oracle, cloud, g, case.kind, case.param, case.coef(cloud), betas=g["betas"])
    vals, _, B = oracle.reference_solve(cloud, case.kind, case.param, case.max_degree, case.coef(cloud), g["q"], betas=g["betas"])
    assert np.max(np.abs(B - g["B"])) <= 1e-9 * np.max(np.abs(g["B"]))
    M = int(g["M"])
    _solution_close(vals, g, rc.golden_K(g), np.concatenate([g["q"], np.zeros(M)]), g["A"][:cloud.N], "robin")


def test_oracle_periodic_advection_diffusion_step(oracle):
    g = rc.load("ref_periodic_10x10")
    case = rc.periodic(u)
    cloud = oracle.RefSquareCloud(10, 10, case.cloud_args["facet_types"])
    rc.assert_cloud_equals_golden(cloud, g)
    assert list(cloud.Np) == [20, 16]
    coef = case.coef(cloud)
    _check_blocks(oracle, cloud, g, case.kind, case.param, coef)
    # right-hand side value(u0)/DT on internal nodes (operators.py:118-147 through assembly.py:455-468)
    A = oracle.assemble_A(cloud, case.kind, case.param, 1)
    cprev = np.linalg.solve(A, np.concatenate([g["u0"], np.zeros(1)]))
    q_int = oracle.eval_field(cloud.sorted_nodes[:cloud.Ni], cloud.sorted_nodes, cprev, case.kind, case.param, "value") / 1e-4
    q = oracle.assemble_q(cloud, q_int, {k: np.zeros(len(cloud.facet_nodes[k])) for k in cloud.facet_types})
    assert np.max(np.abs(q - g["q"])) <= 1e-10 * np.max(np.abs(g["q"]))
    vals, _, _ = oracle.reference_solve(cloud, case.kind, case.param, case.max_degree, coef, g["q"])
    _solution_close(vals, g, rc.golden_K(g), np.concatenate([g["q"], np.zeros(1)]), g["A"][:cloud.N], "periodic")


def test_oracle_all_kernels_degree4_and_field_evaluators(oracle):
    """All five kernels, 15 monomials, an operator with every term of the set (incl. nodal_div_grad) and
    field-dependent coefficients; value / gradient / laplacian of a random coefficient vector at nodes (r = 0 terms,
    nan_to_num) and free points."""
    g = rc.load("ref_kernels_7x6")
    cloud = oracle.RefSquareCloud(7, 6, rc.KERNELS_CLOUD["facet_types"])
    rc.assert_cloud_equals_golden(cloud, g)
    coef = rc.kernels_coef(g, cloud.Ni)
    for name, param in zip(g["kernel_names"], g["kernel_params"]):
        name = str(name)
        _check_blocks(oracle, cloud, g, name, param, coef, prefix=name + "_")
        cf, pts = g[name + "_coeffs"], g["eval_pts"]
        for which, want in (("value", g[name + "_value"]), ("dx", g[name + "_gradient"][:, 0]), ("dy", g[name + "_gradient"][:, 1]),
                            ("laplacian", g[name + "_laplacian"])):
            got = oracle.eval_field(pts, cloud.sorted_nodes, cf, name, param, which)
            assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want)), (name, which)
        # divergence of the vector field (coeffs, coeffs2) (operators.py:294-312) and inv(A) [f1; 0] (assembly.py:404-430)
        div = (oracle.eval_field(pts, cloud.sorted_nodes, cf, name, param, "dx")
               + oracle.eval_field(pts, cloud.sorted_nodes, g[name + "_coeffs2"], name, param, "dy"))
        assert np.max(np.abs(div - g[name + "_divergence"])) <= 1e-12 * np.max(np.abs(g[name + "_divergence"])), name
        fc = np.linalg.solve(oracle.assemble_A(cloud, name, param, 6), np.concatenate([g["f1"], np.zeros(6)]))
        assert np.max(np.abs(fc - g[name + "_field_coeffs"])) <= 1e-8 * np.max(np.abs(g[name + "_field_coeffs"])), name


def test_oracle_config1_full_size_solution(oracle):
    """Config 1 at full size (30x20): the reference's own pde_solver_jit result."""
    g = rc.load("ref_config1_30x20")
    old = np.load(os.path.join(rc.GOLDEN, "config1_30x20_phs3.npz"))
    case = rc.laplace(u, 30, 20)
    cloud = oracle.RefSquareCloud(30, 20, case.cloud_args["facet_types"])
    rc.assert_cloud_equals_golden(cloud, g)
    assert np.array_equal(g["q"], old["q"])
    vals, coeffs, B = oracle.reference_solve(cloud, case.kind, case.param, 1, case.coef(cloud), g["q"])
    assert np.max(np.abs(B[g["B_rows"]] - g["B_sample"])) <= 1e-8 * np.max(np.abs(g["B_sample"]))
    assert np.max(np.abs(vals - g["vals"])) <= 1e-8 * np.max(np.abs(g["vals"]))          # north_star: 1e-8 relative
    assert np.max(np.abs(old["vals"] - g["vals"])) <= 1e-8 * np.max(np.abs(g["vals"]))   # the round-1 oracle-made fixture
    xy = g["sorted_nodes"]
    exact = np.sin(np.pi * xy[:, 0]) * np.cosh(np.pi * xy[:, 1]) / np.cosh(np.pi)
    assert np.max(np.abs(g["vals"] - exact)) <= 2e-2                                         # demos/Laplace/00_...:109-110


@pytest.mark.parametrize("tag", ["vel", "phi"])
def test_gmsh_cloud_of_the_reference_equals_the_restatement(oracle, tag):
    """The reference's GmshCloud on its own fixture mesh.msh: nodes, renumbering, facet lists and the computed
    outward normals equal the oracle-made fixtures the GPU tests use (tests/golden/mesh_msh_cloud_*.npz), bit for bit;
    boundary rows (Neumann rows with those normals) to 1e-12."""
    g = rc.load("ref_mesh_msh_" + tag)
    old = np.load(os.path.join(rc.GOLDEN, "mesh_msh_cloud_%s.npz" % tag))
    for k in ("sorted_nodes", "sorted_outward_normals", "counts", "Np", "facet_names", "facet_sizes", "facet_nodes", "facet_types", "old_of_new"):
        assert np.array_equal(g[k], old[k]), k
    if tag == "phi":
        from helpers import cloud_from_golden
        cloud, _ = cloud_from_golden("mesh_msh_cloud_phi.npz")
        bdPhi, bdP = oracle.assemble_bd_Phi_P(cloud, "polyharmonic", 1, 3)
        r = g["bd_rows"]
        assert rel_err_rowscaled(bdPhi[r], g["bdPhi_sample"]) <= 1e-12 and rel_err_rowscaled(bdP[r], g["bdP_sample"]) <= 1e-12


# ---- the product's host layer against the same files -----------------------------------------------------------------
def test_product_clouds_lowering_and_bc_preparation_match_the_reference():
    for name, case in (("ref_laplace_12x9", rc.laplace(u, 12, 9)), ("ref_robin_11x8", rc.robin(u)),
                       ("ref_periodic_10x10", rc.periodic(u)), ("ref_config1_30x20", rc.laplace(u, 30, 20))):
        g = rc.load(name)
        cloud = u.SquareCloud(**case.cloud_args)
        rc.assert_cloud_equals_golden(cloud, g)
        coef_phi, coef_pol = u.lower_diff_operator(case.op, cloud, case.rbf, case.diff_args)
        assert np.allclose(coef_phi, case.coef(cloud), rtol=1e-15, atol=0) and np.array_equal(coef_phi, coef_pol), name
        # boundary-condition preparation (operators.py:512-555, :628-647): Robin betas per node, arrays per facet
        bc_arr = u.boundary_conditions_func_to_arr(case.bcs, cloud)
        robin, bc_arr = u.duplicate_robin_coeffs(dict(bc_arr), cloud)
        betas = np.array([robin[k] for k in sorted(robin)], dtype=np.float64) if robin else np.zeros(0)
        assert np.allclose(betas, g["betas"], rtol=1e-15, atol=0), name
        if name != "ref_periodic_10x10":               # (its right-hand side evaluates a field: GPU test)
            bc_arr = u.zerofy_periodic_cond(bc_arr, cloud)
            M = u.compute_nb_monomials(case.max_degree, 2)
            q = u.assemble_q(case.rhs, bc_arr, cloud, case.rbf, M, None)
            assert np.allclose(q, g["q"], rtol=1e-15, atol=1e-300), name
    # kernels case: field-dependent coefficients through the symbolic lowering
    g = rc.load("ref_kernels_7x6")
    cloud = u.SquareCloud(**rc.KERNELS_CLOUD)
    rc.assert_cloud_equals_golden(cloud, g)
    coef_phi, coef_pol = u.lower_diff_operator(rc.kernels_operator(u), cloud, rc.kernel_rbf(u, "gaussian", 4.0), [g["f0"], g["f1"]])
    assert np.allclose(coef_phi, rc.kernels_coef(g, cloud.Ni), rtol=1e-15, atol=0) and np.array_equal(coef_phi, coef_pol)


def test_product_gmsh_reader_matches_the_reference_cloud(tmp_path):
    """GmshCloud of the product on a copy of the fixture (the .msh file itself is reference data and is not committed:
    only run where /root/reference is mounted)."""
    mesh = os.path.join(REFERENCE, "updes/tests/data/mesh.msh")
    if not os.path.exists(mesh):
        pytest.skip("needs /root/reference (build container)")
    for tag, ft in (("vel", {"Wall": "d", "Inflow": "d", "Outflow": "n", "Blowing": "d", "Suction": "d"}),
                    ("phi", {"Wall": "n", "Inflow": "n", "Outflow": "d", "Blowing": "n", "Suction": "n"})):
        cloud = u.GmshCloud(filename=mesh, facet_types=ft)
        rc.assert_cloud_equals_golden(cloud, rc.load("ref_mesh_msh_" + tag))


# ---- the generator reproduces the committed files; the reference's own tests pass over the stand-in ------------------
def _run(cmd, cwd, timeout):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    return subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)


def test_generator_reproduces_committed_goldens(tmp_path):
    if not os.path.isdir(REFERENCE):
        pytest.skip("needs /root/reference (build container)")
    r = _run([sys.executable, os.path.join(rc.GOLDEN, "make_reference_golden.py"), "--out", str(tmp_path),
              "--only", "ref_robin_11x8,ref_periodic_10x10"], ROOT, 900)
    assert r.returncode == 0, r.stderr[-2000:]
    for name in ("ref_robin_11x8", "ref_periodic_10x10"):
        new, old = np.load(str(tmp_path / (name + ".npz"))), rc.load(name)
        assert sorted(new.files) == sorted(old.files)
        for k in old.files:
            if old[k].dtype.kind in "fc":
                assert np.allclose(new[k], old[k], rtol=1e-13, atol=1e-13 * max(1.0, float(np.max(np.abs(old[k]))) if old[k].size else 1.0)), (name, k)
            else:
                assert np.array_equal(new[k], old[k]), (name, k)


def test_reference_own_tests_pass_over_the_stand_in():
    """updes/tests/test_{interpolation,integrals,operators}.py of the reference, unmodified, with oracle/refshim on the
    path: the stand-in is faithful enough for the reference's own known-answer tests (constant-field gradient and
    divergence on mesh.msh with a gaussian kernel, the pi/12 integral, the cloud-to-cloud permutation)."""
    if not os.path.isdir(REFERENCE):
        pytest.skip("needs /root/reference (build container)")
    env_path = os.path.join(ROOT, "oracle", "refshim")
    r = subprocess.run([sys.executable, "-W", "ignore", "-m", "pytest", "updes/tests/test_interpolation.py", "updes/tests/test_integrals.py",
                        "updes/tests/test_operators.py", "-q", "-p", "no:cacheprovider"], cwd=REFERENCE,
                       env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=env_path), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "3 passed" in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]


def test_oracle_config3_projection_loop_matches_the_reference_demo(oracle):
    """Config 3: two iterations of simulate_forward_navier_stokes as the reference's demo script itself runs it (its source
    executed unchanged by the generator) against the oracle's restatement of the loop.  Every solve goes through
    inv(A) at cond ~ 1e9 in both pipelines and feeds the next one, hence the cond-scaled bound (measured: 5e-6)."""
    import configs_path  # noqa: F401
    import configs
    from helpers import cloud_from_golden
    g = rc.load("ref_config3_ns_2iter")
    cv, _ = cloud_from_golden("ref_mesh_msh_vel.npz")
    cp, _ = cloud_from_golden("ref_mesh_msh_phi.npz")
    hist = rc.oracle_config3_loop(oracle, u.interpolate_field, configs.config3_boundary_arrays(cv, cp), cv, cp, int(g["nb_iter"]))
    rel = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
    worst = 0.0
    for it, (uu, vv, pp) in enumerate(hist):
        for nm, got, want in (("u", uu, g["u"][it + 1]), ("v", vv, g["v"][it + 1]), ("p", pp, g["p"][it + 1])):
            worst = max(worst, rel(got, want))
            print("iteration %d %s oracle-vs-reference-demo %.2e" % (it, nm, rel(got, want)))
    assert worst <= 2e-5, worst
    assert np.abs(g["u"][-1]).max() > 0.5           # a developed channel flow


def test_oracle_picard_loop_matches_the_reference_multi_solver(oracle):
    """pde_multi_solver of the reference (operators.py:696-771) on two genuinely coupled equations: every sweep solves
    all equations with the PREVIOUS sweep's values.  Restated with explicit coefficient tables on the oracle."""
    g = rc.load("ref_multi_solver_9x8")
    cloud = oracle.RefSquareCloud(9, 8, rc.MULTI_CLOUD["facet_types"])
    rc.assert_cloud_equals_golden(cloud, g)
    Ni, xs = cloud.Ni, cloud.sorted_nodes[:cloud.Ni, 0]
    north, west = np.asarray(cloud.facet_nodes["North"]), np.asarray(cloud.facet_nodes["West"])
    bc = lambda ids: {f: (np.ones(len(v)) if v is ids else np.zeros(len(v))) for f, v in ((k, np.asarray(cloud.facet_nodes[k])) for k in cloud.facet_nodes)}
    bc0 = {f: (np.ones(len(v)) if f == "North" else np.zeros(len(v))) for f, v in cloud.facet_nodes.items()}
    bc1 = {f: (np.ones(len(v)) if f == "West" else np.zeros(len(v))) for f, v in cloud.facet_nodes.items()}
    q0, q1 = oracle.assemble_q(cloud, np.zeros(Ni), bc0), oracle.assemble_q(cloud, -np.ones(Ni), bc1)
    prev = [np.zeros(cloud.N), np.zeros(cloud.N)]
    for k in range(1, int(g["nb_iters"]) + 1):
        c0 = np.stack([-(1.0 + prev[1][:Ni] ** 2), np.zeros(Ni), np.zeros(Ni), np.ones(Ni), np.ones(Ni)], axis=1)
        c1 = np.stack([np.zeros(Ni), xs + prev[0][:Ni], np.zeros(Ni), np.ones(Ni), np.ones(Ni)], axis=1)
        v0, _, _ = oracle.reference_solve(cloud, "polyharmonic", 1, 1, c0, q0)
        v1, _, _ = oracle.reference_solve(cloud, "polyharmonic", 1, 1, c1, q1)
        prev = [v0, v1]
        for i, v in enumerate(prev):
            want = g["vals%d_after_%d" % (i, k)]
            assert np.max(np.abs(v - want)) <= 1e-8 * np.max(np.abs(want)), (k, i)
    assert np.max(np.abs(g["vals0_after_3"] - g["vals0_after_1"])) >= 1e-4 * np.max(np.abs(g["vals0_after_3"]))     # the coupling is real


def test_random_problems_oracle_and_product_host_match_the_reference(oracle):
    """16 random small problems run through the reference (tests/golden/ref_fuzz_16.npz): facet types d / n / r /
    periodic pairs in a random dict order (corner precedence, periodic-class suffixes and the renumbering depend on it),
    Neumann + Robin + periodic mixes, every kernel, degrees 0-4, a fully general operator with five nodal fields."""
    seen = set()
    for k, nx, ny, facets, kind, param, M, fields, betas, g in rc.fuzz_cases():
        ocloud = oracle.RefSquareCloud(nx, ny, dict(facets))
        pcloud = u.SquareCloud(Nx=nx, Ny=ny, facet_types=dict(facets))
        rc.assert_cloud_equals_golden(ocloud, g)
        rc.assert_cloud_equals_golden(pcloud, g)
        coef = fields[:, :ocloud.Ni].T.copy()
        D = oracle.assemble_diffMat(ocloud, kind, param, M, coef, betas)
        assert rel_err_rowscaled(D, g["diffMat"]) <= 1e-12 and true_rel_err(D, g["diffMat"]) <= 1e-10, k
        # the product's symbolic lowering of the same operator gives exactly these coefficients
        def op(x, center, rbf, monomial, f):
            return (f[0] * u.nodal_value(x, center, rbf, monomial) + u.dot([f[1], f[2]], u.nodal_gradient(x, center, rbf, monomial))
                    + u.nodal_div_grad(x, center, rbf, monomial, (f[3], f[4])))
        cphi, cpol = u.lower_diff_operator(op, pcloud, rc.kernel_rbf(u, kind, param), list(fields))
        assert np.array_equal(cphi, coef) and np.array_equal(cpol, coef), k
        seen.add((pcloud.Nn > 0, pcloud.Nr > 0, len(pcloud.Np)))
    assert {(True, True, 1), (False, False, 2), (False, True, 0)} <= seen          # Neumann + Robin + periodic; doubly periodic; Robin only


def test_gmsh_readers_match_the_reference_on_generated_meshes(tmp_path, oracle):
    """Gmsh-4.0 channel meshes written by tests/golden/make_msh.py (regenerated here): the product's reader and the
    oracle's restatement against what the reference's GmshCloud made of the same files -- corner nodes go to the facet
    that comes first in the facet_types dict, normals point away from the nearest internal node."""
    from golden.make_msh import write_channel_msh
    g = rc.load("ref_generated_msh")
    for tag in ("a", "b"):
        nx, ny = (int(v) for v in g[tag + "_grid"])
        facets = {str(f): str(t) for f, t in g[tag + "_facets_in"]}
        path = str(tmp_path / ("channel_%s.msh" % tag))
        write_channel_msh(path, nx=nx, ny=ny)
        sub = {k: g["%s_%s" % (tag, k)] for k in rc.CLOUD_KEYS}
        for cloud in (u.GmshCloud(path, facet_types=dict(facets)), oracle.RefGmshCloud(path, dict(facets))):
            rc.assert_cloud_equals_golden(cloud, sub)
            old_of_new = [o for o, _ in sorted(cloud.renumbering_map.items(), key=lambda kv: kv[1])]
            assert old_of_new == g[tag + "_old_of_new"].tolist()


def test_reference_style_operators_with_a_foreign_array_library_lower_through_numeric_probes():
    """Operators as the reference's demos write them hand the nodal terms to jnp (`jnp.dot(U_prev, nodal_gradient(...))`):
    symbolic lowering cannot follow that, the numeric-probe fallback can -- and still rejects non-linear / affine bodies.
    Runs in a subprocess because it puts the JAX stand-in on sys.path; with /root/reference mounted the operator source is
    the demo's own text (demos/NavierStokes/30_channel_flow_blowing_suction.py:61-65)."""
    r = _run([sys.executable, "-W", "ignore", os.path.join(ROOT, "tests", "lowering_with_foreign_arrays.py")], ROOT, 600)
    assert r.returncode == 0 and "OK bc + rhs" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_oracle_time_steps_match_the_reference_advection_demo(oracle):
    """Config 2 as the reference's demo defines it (demos/Advection/01_adv_diff_periodic.py:34-113: 35x35 doubly periodic
    cloud, K = 0.08, VEL = (100, 0), DT = 1e-4, gaussian bump; definitions executed from the demo's source by the
    generator): three implicit steps, each restated on the oracle from the reference's previous field."""
    g = rc.load("ref_config2_advdiff_3steps")
    cloud = oracle.RefSquareCloud(35, 35, {"South": "p1", "North": "p1", "West": "p2", "East": "p2"})
    rc.assert_cloud_equals_golden(cloud, g)
    assert list(cloud.Np) == [70, 66] and int(g["max_degree"]) == 0
    DT, K, VEL = float(g["DT"]), float(g["K"]), g["VEL"]
    coef = np.tile([1 / DT, VEL[0], VEL[1], -K, -K], (cloud.Ni, 1))
    A = oracle.assemble_A(cloud, "polyharmonic", 1, 1)
    zero_bc = {k: np.zeros(len(cloud.facet_nodes[k])) for k in cloud.facet_types}
    xy = cloud.sorted_nodes
    assert np.allclose(g["u"][0], np.exp(-((xy[:, 0] - 0.35) ** 2 + (xy[:, 1] - 0.5) ** 2) / (2 * 0.1 ** 2)), rtol=1e-14, atol=1e-16)
    for s in range(g["u"].shape[0] - 1):
        cprev = np.linalg.solve(A, np.concatenate([g["u"][s], np.zeros(1)]))
        q_int = oracle.eval_field(xy[:cloud.Ni], xy, cprev, "polyharmonic", 1, "value") / DT
        vals, _, _ = oracle.reference_solve(cloud, "polyharmonic", 1, 0, coef, oracle.assemble_q(cloud, q_int, zero_bc))
        assert np.max(np.abs(vals - g["u"][s + 1])) <= 1e-8 * np.max(np.abs(g["u"][s + 1])), s


def test_reference_quirks_are_facts_of_the_reference_output():
    """SURVEY.md Q1-Q4 read off the reference's own matrices (not off the restatement): Q1 the diagonal of Phi is 0 even
    where phi(0) = 1 (the node is dropped from its own support, cloud.py:110-112); Q2 periodic rows do write their own
    column; Q3 Robin rows of bd(Phi) use the Neumann nodes' normal slots when Nn > 0 (assembly.py:206); Q4 derivatives
    at r = 0 are 0 (nan_to_num) in the field evaluators."""
    g = rc.load("ref_kernels_7x6")
    N = g["sorted_nodes"].shape[0]
    for name in ("gaussian", "multiquadric", "inverse_multiquadric"):                    # phi(0) = 1 for these
        Phi = g[name + "_A"][:N, :N]
        assert np.all(np.diag(Phi) == 0.0) and np.all(Phi[~np.eye(N, dtype=bool)] != 0.0)
    g = rc.load("ref_periodic_10x10")
    Ni = int(g["counts"][1])
    assert np.allclose([g["bdPhi"][k, Ni + k] for k in range(10)], -1.0)       # phi(0) - phi(1) with phi = r^3: own column written
    # Q3: on the 11x8 Robin + Neumann cloud the first Robin row (West facet, true normal (-1, 0)) carries the South facet's
    # normal (0, -1): with beta known, (row - beta * phi) must equal grad(phi) . (0, -1), not grad(phi) . (-1, 0)
    g = rc.load("ref_robin_11x8")
    Ni, Nd, Nn = (int(v) for v in g["counts"][1:4])
    xy, eps = g["sorted_nodes"], 3.0
    i = Ni + Nd + Nn                                                           # first Robin node
    assert np.allclose(g["sorted_outward_normals"][0], [0.0, -1.0]) and np.allclose(g["sorted_outward_normals"][Nn], [-1.0, 0.0])
    d = xy[i] - xy                                                             # x_i - x_j
    phi = np.exp(-eps ** 2 * (d ** 2).sum(1))
    grad = -2 * eps ** 2 * d * phi[:, None]
    row = g["bdPhi"][i - Ni]
    cols = np.arange(xy.shape[0]) != i
    quirk = g["betas"][0] * phi + grad @ np.array([0.0, -1.0])
    true_normal = g["betas"][0] * phi + grad @ np.array([-1.0, 0.0])
    assert np.allclose(row[cols], quirk[cols], rtol=1e-12, atol=1e-15) and not np.allclose(row[cols], true_normal[cols], rtol=1e-3, atol=1e-6)
    # Q4: the Laplacian of a gaussian field evaluated AT a node omits that node's own term (true value -4 eps^2 c_i)
    g = rc.load("ref_kernels_7x6")
    xy, pts, c, eps = g["sorted_nodes"], g["eval_pts"], g["gaussian_coeffs"], 4.0
    N = xy.shape[0]
    k = 0                                                                       # eval_pts[0] is node 0
    assert np.array_equal(pts[k], xy[0])
    s = ((pts[k] - xy) ** 2).sum(1)
    lap_terms = (4 * eps ** 4 * s - 4 * eps ** 2) * np.exp(-eps ** 2 * s) * c[:N]
    poly = 2 * c[N + 3] + 2 * c[N + 5] + 6 * c[N + 6] * pts[k, 0] + 2 * c[N + 7] * pts[k, 1] + 2 * c[N + 8] * pts[k, 0] + 6 * c[N + 9] * pts[k, 1] \
        + 12 * c[N + 10] * pts[k, 0] ** 2 + 6 * c[N + 11] * pts[k, 0] * pts[k, 1] + 2 * c[N + 12] * (pts[k, 0] ** 2 + pts[k, 1] ** 2) \
        + 6 * c[N + 13] * pts[k, 0] * pts[k, 1] + 12 * c[N + 14] * pts[k, 1] ** 2
    with_self, without_self = lap_terms.sum() + poly, lap_terms.sum() - lap_terms[0] + poly
    assert np.isclose(g["gaussian_laplacian"][k], without_self, rtol=1e-10) and not np.isclose(g["gaussian_laplacian"][k], with_self, rtol=1e-6)


def test_stand_in_has_the_jax_semantics_the_hot_path_relies_on():
    """x64 dtypes, functional .at[].set, NaN gradients of the distance at coincident points + nan_to_num, jacfwd(grad),
    vmap with in_axes=None and Python-number outputs, fori_loop / Partial / tree_map, LAPACK inv, QR solve
    (tests/refshim_semantics.py, in a subprocess because it imports the stand-in)."""
    r = _run([sys.executable, "-W", "ignore", os.path.join(ROOT, "tests", "refshim_semantics.py")], ROOT, 300)
    assert r.returncode == 0 and "SEMANTICS OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_oracle_matches_the_unmodified_laplace_demo(oracle):
    """demos/Laplace/00_laplace_with_rbf.py, the whole script run unmodified by the generator (30x30): its solution, the
    Laplacian of its solution at the nodes (r = 0 terms dropped by nan_to_num), and the mean-square errors it prints."""
    g = rc.load("ref_laplace_demo_30x30")
    cloud = oracle.RefSquareCloud(30, 30, {"South": "n", "West": "d", "North": "d", "East": "d"})
    rc.assert_cloud_equals_golden(cloud, g)
    xy = cloud.sorted_nodes
    bc = {f: (np.sin(np.pi * xy[ids, 0]) if f == "North" else np.zeros(len(ids))) for f, ids in cloud.facet_nodes.items()}
    vals, _, _ = oracle.reference_solve(cloud, "polyharmonic", 1, 1, np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1)),
                                        oracle.assemble_q(cloud, np.zeros(cloud.Ni), bc))
    assert np.max(np.abs(vals - g["vals"])) <= 1e-8 * np.max(np.abs(g["vals"]))
    exact = np.sin(np.pi * xy[:, 0]) * np.cosh(np.pi * xy[:, 1]) / np.cosh(np.pi)
    assert np.allclose(exact, g["exact"], rtol=1e-14, atol=1e-16)
    assert np.isclose(np.mean((exact - vals) ** 2), float(g["mse_total"]), rtol=1e-5)                 # the figure the demo prints
    south = np.asarray(cloud.facet_nodes["South"])
    assert np.isclose(np.mean((exact[south] - vals[south]) ** 2), float(g["mse_neumann"]), rtol=1e-5)
    lap = oracle.eval_field(xy, xy, g["coeffs"], "polyharmonic", 1, "laplacian")
    assert np.max(np.abs(lap - g["laplacian_at_nodes"])) <= 1e-11 * np.max(np.abs(g["laplacian_at_nodes"]))


def test_oracle_matches_the_unmodified_darcy_demo(oracle):
    """demos/Darcy/00_darcy_flow.py run unmodified by the generator (20x20, all Dirichlet): an identity-operator solve with
    polyharmonic a = 2 and -div(k grad u) = 1 through nodal_div_grad with a nodal field, thin_plate a = 3, degree 2.
    cond(A) ~ 1e9 - 1e10 here, so two inverse-based pipelines agree to a few 1e-8: the bound is 4x the sum of their own
    distances from the exactly solved discrete system."""
    g = rc.load("ref_darcy_demo_20x20")
    cloud = oracle.RefSquareCloud(20, 20, {"South": "d", "North": "d", "West": "d", "East": "d"})
    rc.assert_cloud_equals_golden(cloud, g)
    M, Ni, N = 6, cloud.Ni, cloud.N
    zero_bc = {k: np.zeros(len(v)) for k, v in cloud.facet_nodes.items()}

    def check(kind, a, coef, q, want, what):
        vals, _, _ = oracle.reference_solve(cloud, kind, a, 2, coef, q)
        K, A = oracle.assemble_K(cloud, kind, a, M, coef), oracle.assemble_A(cloud, kind, a, M)
        exact, _ = exact_solution(K, np.concatenate([q, np.zeros(M)]), A[:N])
        s = np.max(np.abs(exact))
        e_gold, e_orc, d = np.max(np.abs(want - exact)) / s, np.max(np.abs(vals - exact)) / s, np.max(np.abs(vals - want)) / s
        print("%s: reference-vs-exact %.2e oracle-vs-exact %.2e oracle-vs-reference %.2e" % (what, e_gold, e_orc, d))
        assert d <= max(1e-8, 4.0 * (e_gold + e_orc)) and e_gold <= 1e-6, (what, d, e_gold, e_orc)

    A1 = oracle.assemble_A(cloud, "polyharmonic", 2, M)
    cperm = np.linalg.solve(A1, np.concatenate([g["permeability"], np.zeros(M)]))
    q1 = oracle.assemble_q(cloud, oracle.eval_field(cloud.sorted_nodes[:Ni], cloud.sorted_nodes, cperm, "polyharmonic", 2, "value"), zero_bc)
    check("polyharmonic", 2, np.tile([1.0, 0, 0, 0, 0], (Ni, 1)), q1, g["perm_vals"], "permeability (identity operator)")
    k = g["perm_vals"][:Ni]
    coef2 = np.stack([np.zeros(Ni), np.zeros(Ni), np.zeros(Ni), -k, -k], axis=1)
    check("thin_plate", 3, coef2, oracle.assemble_q(cloud, np.ones(Ni), zero_bc), g["u_vals"], "Darcy solution")


def test_integrate_field_host_logic_matches_the_reference(oracle, monkeypatch):
    """integrate_field (operators.py:381-451; the function behind the reference's test_integrals.py): the product's point
    sets and weights against the number the reference returned for the same coefficients.  The evaluator is replaced by
    the oracle's here (host logic only; the GPU test uses the real one)."""
    from updes_b200 import operators as ops
    g = rc.load("ref_integrals_12x12")
    cloud = u.SquareCloud(Nx=12, Ny=12, facet_types={"North": "d", "South": "d", "East": "d", "West": "d"})
    rc.assert_cloud_equals_golden(cloud, g)
    monkeypatch.setattr(ops, "value", lambda x, cf, centers, rbf=None, clip_val=None: oracle.eval_field(
        np.asarray(x, dtype=float).reshape(-1, 2), np.asarray(centers), np.ascontiguousarray(cf), "polyharmonic", 5, "value"))
    val = ops.integrate_field(g["coeffs"], cloud, rc.kernel_rbf(u, "polyharmonic", 5), 3)
    assert np.isclose(val, float(g["integral"]), rtol=1e-12)
    assert abs(float(g["integral"]) - np.pi / 12) < 1e-1                                   # test_integrals.py:83-84
    with pytest.raises(AssertionError):
        ops.integrate_field(g["coeffs"], object(), u.polyharmonic, 3)
    # the oracle agrees on the rebuilt field too (value_vec at the nodes from the reference's coefficients)
    rebuilt = oracle.eval_field(cloud.sorted_nodes, cloud.sorted_nodes, g["coeffs"], "polyharmonic", 5, "value")
    assert np.max(np.abs(rebuilt - g["rebuilt"])) <= 1e-9 * max(1.0, np.max(np.abs(g["rebuilt"])))


@pytest.mark.parametrize("name,args", [
    ("ref_laplace_12x9", dict(Nx=12, Ny=9, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})),
    ("ref_periodic_10x10", dict(Nx=10, Ny=10, facet_types={"South": "p1", "North": "p1", "West": "p2", "East": "p2"})),
    ("ref_robin_11x8", None)])
def test_lazy_local_supports_have_the_reference_neighbour_order_up_to_ties(name, args):
    """cloud.local_supports[i] (reference cloud.py:83-112; read by demos/Advection/00_...:68): computed per access instead of
    stored as an (N, N-1) table.  Against the first rows of the reference's own table: the same nodes, self excluded, at the
    same sequence of distances; equidistant nodes may come in another order (BallTree's there, by id here)."""
    g = rc.load(name)
    cloud = u.SquareCloud(**(args if args is not None else rc.robin(u).cloud_args))
    ref = g["supports_first_rows"]
    ls = cloud.local_supports
    assert len(ls) == cloud.N and list(ls.keys()) == list(range(cloud.N)) and (cloud.N - 1) in ls and cloud.N not in ls
    xy = cloud.sorted_nodes
    for i in range(ref.shape[0]):
        ours = np.array(ls[i])
        assert ours.shape == (cloud.N - 1,) and i not in ours and sorted(ours) == sorted(ref[i])
        dist = lambda idx: np.linalg.norm(xy[idx] - xy[i], axis=1)
        assert np.all(np.diff(dist(ours)) >= -1e-15)
        assert np.allclose(dist(ours), dist(ref[i]), rtol=0, atol=1e-14)
    table = cloud.sorted_local_supports
    assert table.shape == (cloud.N, cloud.N - 1) and np.array_equal(table[2], np.array(ls[2]))
    with pytest.raises(KeyError):
        ls[cloud.N]


ADV00_FACETS = {"South": "d", "West": "d", "North": "d", "East": "n"}
ADV02_FACETS = {"South": "p1", "North": "p1", "West": "p2", "East": "p2"}


def _oracle_advection_steps(oracle, g, cloud, M, sink=None):
    """One implicit step per golden step, restated on the oracle from the REFERENCE's previous field: coefficients of the
    previous field through A, value(x)/DT on internal nodes, reference formulation of the solve."""
    DT, K, VEL = float(g["DT"]), float(g["K"]), g["VEL"]
    Ni, xy = cloud.Ni, cloud.sorted_nodes
    c0 = np.full(Ni, 1 / DT) + (0.0 if sink is None else sink[:Ni])
    coef = np.stack([c0, np.full(Ni, VEL[0]), np.full(Ni, VEL[1]), np.full(Ni, -K), np.full(Ni, -K)], axis=1)
    A = oracle.assemble_A(cloud, "polyharmonic", 1, M)
    zero_bc = {k: np.zeros(len(cloud.facet_nodes[k])) for k in cloud.facet_types}
    worst = 0.0
    for s in range(g["u"].shape[0] - 1):
        cprev = np.linalg.solve(A, np.concatenate([g["u"][s], np.zeros(M)]))
        q_int = oracle.eval_field(xy[:Ni], xy, cprev, "polyharmonic", 1, "value") / DT
        vals, _, _ = oracle.reference_solve(cloud, "polyharmonic", 1, int(g["max_degree"]), coef, oracle.assemble_q(cloud, q_int, zero_bc))
        worst = max(worst, float(np.max(np.abs(vals - g["u"][s + 1])) / np.max(np.abs(g["u"][s + 1]))))
    return worst


def test_oracle_matches_the_advection_demo_with_outflow(oracle):
    """demos/Advection/00_advection_with_rbf.py (definitions executed from its source by the generator): 40x20, Dirichlet on
    three sides and a Neumann outflow, degree 1, u0 = 0.95 on the 20 nearest neighbours of node 8 read from
    cloud.local_supports; two implicit steps."""
    g = rc.load("ref_advection00_2steps")
    cloud = oracle.RefSquareCloud(40, 20, ADV00_FACETS)
    rc.assert_cloud_equals_golden(cloud, g)
    assert int(g["max_degree"]) == 1 and int(g["source_id"]) == 8 and len(g["source_neighbors"]) == 20
    u0 = np.zeros(cloud.N); u0[g["source_neighbors"]] = 0.95
    assert np.array_equal(g["u"][0], u0)
    assert _oracle_advection_steps(oracle, g, cloud, 3) <= 1e-8
    # the product's per-access supports give the demo the same neighbourhood: the same 20 nodes unless the 20th and 21st
    # neighbours are equidistant (ties are ordered by BallTree in the reference), and always the same distances
    pc = u.SquareCloud(Nx=40, Ny=20, facet_types=ADV00_FACETS)
    ours = np.array(pc.local_supports[8])
    xy = pc.sorted_nodes
    dist = lambda idx: np.linalg.norm(xy[idx] - xy[8], axis=1)
    assert np.allclose(dist(ours), dist(g["source_support"]), rtol=0, atol=1e-14)
    if dist(ours[19:20])[0] < dist(ours[20:21])[0] - 1e-12:
        assert set(ours[:20].tolist()) == set(g["source_neighbors"].tolist())


def test_oracle_matches_the_advection_demo_with_a_sink_field(oracle):
    """demos/Advection/02_adv_diff_periodic_with_sink.py: pure advection (K = 0, VEL = 500) on the doubly periodic 35x35
    cloud, degree 0, the operator's `fields[0] * val` term fed by a nodal sink field through diff_args; two steps."""
    g = rc.load("ref_advection02_sink_2steps")
    cloud = oracle.RefSquareCloud(35, 35, ADV02_FACETS)
    rc.assert_cloud_equals_golden(cloud, g)
    assert float(g["K"]) == 0.0 and int(g["max_degree"]) == 0 and g["u_sink"].shape == (cloud.N,)
    assert _oracle_advection_steps(oracle, g, cloud, 1, sink=g["u_sink"]) <= 1e-8


GS_FACETS = {"South": "p0", "North": "p0", "West": "p1", "East": "p1"}
ALL_NEUMANN = {"South": "n", "North": "n", "West": "n", "East": "n"}


def test_oracle_matches_the_gray_scott_demo_periodic_ids_with_degree_one(oracle):
    """demos/Gray-Scott/001_gray-scott.py: periodic ids "p0" / "p1" (suffixed "p00", "p01", "p12", "p13" by facet position),
    three monomial columns beside the periodic rows, u0 through cloud.local_supports; two implicit steps."""
    g = rc.load("ref_grayscott001_2steps")
    cloud = oracle.RefSquareCloud(40, 20, GS_FACETS)
    rc.assert_cloud_equals_golden(cloud, g)
    assert [str(t) for t in g["facet_types"]] == ["p00", "p01", "p12", "p13"] and int(g["max_degree"]) == 1
    assert _oracle_advection_steps(oracle, g, cloud, 3) <= 1e-8
    pc = u.SquareCloud(Nx=40, Ny=20, facet_types=GS_FACETS)
    rc.assert_cloud_equals_golden(pc, g)
    sid = int(g["source_id"])
    xy = pc.sorted_nodes
    dist = lambda idx: np.linalg.norm(xy[idx] - xy[sid], axis=1)
    assert np.allclose(dist(np.array(pc.local_supports[sid])), dist(g["source_support"]), rtol=0, atol=1e-14)


def wave_exact_steps(oracle, g, cloud):
    """For every golden step of the wave demo: (exact next field, oracle's next field).  'Exact' = the discrete step solved
    with extended-precision refinement throughout (coefficients of the two previous fields through A, then K), from the
    REFERENCE's previous fields.  cond(A) = 2e12 and cond(K) = 2e17 (rows of size 1/DT^2 = 4e6 beside rows of size 1) here,
    so the inverse-based pipelines -- the reference's and the oracle's -- each sit ~5e-6 from it."""
    DT, C, M = float(g["DT"]), float(g["C"]), 6
    Ni, N, xy = cloud.Ni, cloud.N, cloud.sorted_nodes
    coef = np.tile([1 / DT ** 2, 0.0, 0.0, C, C], (Ni, 1))
    A = oracle.assemble_A(cloud, "polyharmonic", 3, M)
    K = oracle.assemble_K(cloud, "polyharmonic", 3, M, coef)
    zero_bc = {k: np.zeros(len(cloud.facet_nodes[k])) for k in cloud.facet_types}
    val = lambda c: oracle.eval_field(xy[:Ni], xy, c, "polyharmonic", 3, "value")
    out = []
    for s in range(1, g["u"].shape[0] - 1):
        cx = [exact_solution(A, np.concatenate([g["u"][t], np.zeros(M)]), A[:N])[1] for t in (s, s - 1)]
        q = oracle.assemble_q(cloud, (2 * val(cx[0]) - val(cx[1])) / DT ** 2, zero_bc)
        exact, _ = exact_solution(K, np.concatenate([q, np.zeros(M)]), A[:N])
        cn = [np.linalg.solve(A, np.concatenate([g["u"][t], np.zeros(M)])) for t in (s, s - 1)]
        qn = oracle.assemble_q(cloud, (2 * val(cn[0]) - val(cn[1])) / DT ** 2, zero_bc)
        vals, _, _ = oracle.reference_solve(cloud, "polyharmonic", 3, 2, coef, qn)
        out.append((exact, vals))
    return out


def test_oracle_matches_the_wave_demo_all_neumann_two_fields(oracle):
    """demos/Wave/00_wave.py: no Dirichlet row at all (four Neumann facets), polyharmonic a = 3, degree 2, the right-hand side
    evaluates two nodal fields; two steps, each restated from the reference's two previous fields.  Agreement is asserted to
    the accuracy the inverse-based pipelines have on this system (see wave_exact_steps)."""
    g = rc.load("ref_wave00_2steps")
    cloud = oracle.RefSquareCloud(25, 25, ALL_NEUMANN)
    rc.assert_cloud_equals_golden(cloud, g)
    assert cloud.Nd == 0 and cloud.Nn == 96 and int(g["max_degree"]) == 2 and g["u"].shape[0] == 4
    for s, (exact, vals) in enumerate(wave_exact_steps(oracle, g, cloud), start=1):
        sc = np.max(np.abs(exact))
        e_gold, e_ours, d = (np.max(np.abs(g["u"][s + 1] - exact)) / sc, np.max(np.abs(vals - exact)) / sc,
                             np.max(np.abs(vals - g["u"][s + 1])) / sc)
        print("wave demo step %d: reference-vs-exact %.2e, oracle-vs-exact %.2e, oracle-vs-reference %.2e" % (s, e_gold, e_ours, d))
        assert d <= max(1e-8, 4.0 * e_gold) and e_ours <= max(1e-8, 4.0 * e_gold)


def test_cartesian_finite_difference_helpers_match_the_reference(tmp_path):
    """cartesian_gradient(_vec), enforce_cartesian_gradient_neumann and apply_neumann_conditions (operators.py:211-291,
    :483-509; demos/NavierStokes/11_...:187, 16_...:181) against what the reference returned on two square clouds and a
    generated channel mesh -- including its quirks: the divisor of cartesian_gradient is the distance to the FARTHEST node of
    the support (a variable left over from the loop), a node without a neighbour behind a direction returns early, and
    enforce_... writes one number into both components."""
    from golden.make_msh import write_channel_msh
    g = rc.load("ref_cartesian_helpers")
    path = str(tmp_path / "channel.msh")
    write_channel_msh(path, nx=9, ny=6)
    clouds = {"sq_a": u.SquareCloud(Nx=7, Ny=6, facet_types={"South": "n", "West": "d", "North": "d", "East": "n"}),
              "sq_b": u.SquareCloud(Nx=8, Ny=5, facet_types={"South": "d", "West": "n", "North": "r", "East": "d"}),
              "msh": u.GmshCloud(path, facet_types={"Wall": "n", "Inflow": "d", "Outflow": "n"})}
    for tag, c in clouds.items():
        assert np.array_equal(c.sorted_nodes, g[tag + "_nodes"])
        f, g0 = g[tag + "_f"], g[tag + "_g0"]
        grad = u.cartesian_gradient_vec(range(c.N), f, c)
        assert np.allclose(grad, g[tag + "_grad"], rtol=1e-13, atol=1e-15), tag
        assert np.any(np.all(grad == 0.0, axis=1)), "nodes with nothing behind +x return (0, 0)"
        assert np.allclose(u.cartesian_gradient(3, f, c, clip_val=0.05), g[tag + "_grad3_clipped"], rtol=1e-13, atol=1e-15)
        enforced = u.enforce_cartesian_gradient_neumann(f, g0, {}, c)
        assert np.allclose(enforced, g[tag + "_enforced"], rtol=1e-13, atol=1e-15), tag
        assert np.allclose(u.apply_neumann_conditions(f, {}, c), g[tag + "_applied"], rtol=0, atol=0), tag
        assert np.array_equal(g0, g[tag + "_g0"]) and not np.shares_memory(enforced, g0)      # inputs are not modified
