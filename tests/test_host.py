"""Host logic and ABI surface (CPU only): clouds vs the oracle's literal restatement, operator
lowering, BC preparation, row descriptors, and that the C-ABI library loads and exports every
symbol declared in include/updes_b200.h (no compute calls without a GPU)."""
import os
import re
import sys
from functools import partial

import numpy as np
import pytest

import updes_b200 as u
from updes_b200 import assembly as asm
from helpers import CONFIG1_FACETS, CONFIG2_FACETS, advdiff_op, cloud_from_golden, laplace_op

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_MESH = "/root/reference/updes/tests/data/mesh.msh"


def _same_cloud(a, b):
    assert (a.N, a.Ni, a.Nd, a.Nn, a.Nr, list(a.Np)) == (b.N, b.Ni, b.Nd, b.Nn, b.Nr, list(b.Np))
    assert np.array_equal(a.sorted_nodes, b.sorted_nodes)
    assert np.array_equal(a.sorted_outward_normals, b.sorted_outward_normals)
    assert a.facet_nodes == b.facet_nodes and a.facet_types == b.facet_types
    assert a.node_types == b.node_types and a.renumbering_map == b.renumbering_map


@pytest.mark.parametrize("facets", [CONFIG1_FACETS, CONFIG2_FACETS,
                                    {"South": "r", "West": "n", "North": "d", "East": "p1"},
                                    {"North": "p3", "East": "r", "West": "r", "South": "p3"},
                                    {k: "d" for k in ("South", "West", "North", "East")}])
@pytest.mark.parametrize("shape,seed", [((9, 7), None), ((6, 11), 3), ((30, 20), None)])
def test_square_cloud_matches_reference_restatement(oracle, facets, shape, seed):
    a = u.SquareCloud(Nx=shape[0], Ny=shape[1], facet_types=facets, noise_key=seed)
    b = oracle.RefSquareCloud(shape[0], shape[1], facets, noise_seed=seed)
    _same_cloud(a, b)


def test_square_cloud_counts_of_the_configs():
    c1 = u.SquareCloud(Nx=30, Ny=20, facet_types=CONFIG1_FACETS)
    assert (c1.N, c1.Ni, c1.Nd, c1.Nn) == (600, 504, 66, 30)
    c2 = u.SquareCloud(Nx=35, Ny=35, facet_types=CONFIG2_FACETS)
    assert (c2.N, c2.Ni, c2.Np) == (1225, 1089, [70, 66])
    # periodic pairs (i, i + Np/2) are geometric opposites
    s = c2.Ni
    for g, nb in enumerate(c2.Np):
        xy1, xy2 = c2.sorted_nodes[s:s + nb // 2], c2.sorted_nodes[s + nb // 2:s + nb]
        same = 0 if g == 0 else 1
        assert np.allclose(xy1[:, same], xy2[:, same])
        s += nb
    c4 = u.SquareCloud(Nx=300, Ny=300, facet_types=CONFIG1_FACETS)
    assert (c4.N, c4.Ni) == (90000, 88804)


@pytest.mark.skipif(not os.path.exists(REF_MESH), reason="reference fixture not mounted (GPU box)")
@pytest.mark.parametrize("tag,facets", [("vel", {"Wall": "d", "Inflow": "d", "Outflow": "n", "Blowing": "d", "Suction": "d"}),
                                        ("phi", {"Wall": "n", "Inflow": "n", "Outflow": "d", "Blowing": "n", "Suction": "n"})])
def test_gmsh_cloud_on_reference_fixture(oracle, tag, facets):
    """updes/tests/data/mesh.msh: product reader == literal restatement == committed golden arrays."""
    a = u.GmshCloud(REF_MESH, facet_types=facets)
    _same_cloud(a, oracle.RefGmshCloud(REF_MESH, facets))
    g = np.load(os.path.join(GOLDEN, "mesh_msh_cloud_%s.npz" % tag))
    assert np.array_equal(a.sorted_nodes, g["sorted_nodes"])
    assert np.array_equal(a.sorted_outward_normals, g["sorted_outward_normals"])
    assert list(g["counts"]) == [a.N, a.Ni, a.Nd, a.Nn, a.Nr]
    assert a.N == 1385 and a.Ni == 1227
    assert np.allclose(np.linalg.norm(a.sorted_outward_normals, axis=1), 1.0)


def test_gmsh_reader_on_generated_mesh(tmp_path, oracle):
    """A Gmsh-4.0 ASCII channel mesh written by tests/golden/make_msh.py (own data)."""
    from golden.make_msh import write_channel_msh
    path = str(tmp_path / "channel.msh")
    write_channel_msh(path, nx=13, ny=9)
    facets = {"Wall": "d", "Inflow": "n", "Outflow": "r"}
    a = u.GmshCloud(path, facet_types=facets)
    _same_cloud(a, oracle.RefGmshCloud(path, facets))
    assert a.N == 13 * 9 and a.Ni == 11 * 7
    # inflow normals point to -x, outflow to +x
    inflow = [i - a.Ni - a.Nd for i in a.facet_nodes["Inflow"]]
    assert np.allclose(a.sorted_outward_normals[inflow], [-1.0, 0.0])


def test_cloud_from_golden_arrays_round_trip(oracle):
    """Cloud.from_arrays on the committed fixtures reproduces the clouds they were made from."""
    c1, g = cloud_from_golden("config1_30x20_phs3.npz")
    ref1 = oracle.RefSquareCloud(30, 20, CONFIG1_FACETS)       # (the original-id renumbering map is not stored)
    assert np.array_equal(c1.sorted_nodes, ref1.sorted_nodes) and c1.facet_nodes == ref1.facet_nodes
    assert np.array_equal(c1.sorted_outward_normals, ref1.sorted_outward_normals) and c1.node_types == ref1.node_types
    assert (c1.N, c1.Ni, c1.Nd, c1.Nn, c1.Nr) == (ref1.N, ref1.Ni, ref1.Nd, ref1.Nn, ref1.Nr)
    c2, _ = cloud_from_golden("config2_35x35_periodic.npz")
    ref2 = oracle.RefSquareCloud(35, 35, CONFIG2_FACETS, noise_seed=7)
    assert np.array_equal(c2.sorted_nodes, ref2.sorted_nodes) and c2.Np == ref2.Np and c2.facet_nodes == ref2.facet_nodes
    assert np.array_equal(c2.sorted_outward_normals, ref2.sorted_outward_normals) and c2.node_types == ref2.node_types
    c3, _ = cloud_from_golden("mesh_msh_cloud_vel.npz")
    assert (c3.N, c3.Ni, c3.Nd, c3.Nn) == (1385, 1227, 109, 49)
    assert {k: len(v) for k, v in c3.facet_nodes.items()} == {"Wall": 44, "Inflow": 49, "Outflow": 49, "Blowing": 8, "Suction": 8}


def test_rbf_identification():
    assert u.identify_rbf(u.polyharmonic) == ("polyharmonic", 1.0)
    assert u.identify_rbf(partial(u.polyharmonic, a=2)) == ("polyharmonic", 2.0)
    assert u.identify_rbf(partial(u.gaussian, eps=10.0)) == ("gaussian", 10.0)
    assert u.identify_rbf(partial(partial(u.thin_plate, a=3))) == ("thin_plate", 3.0)
    with pytest.raises(TypeError):
        u.identify_rbf(lambda x, c: 0.0)
    with pytest.raises(TypeError):
        u.identify_rbf(partial(u.gaussian, a=1))
    assert u.compute_nb_monomials(1, 2) == 3 and u.compute_nb_monomials(4, 2) == 15
    with pytest.raises(NotImplementedError):
        u.compute_nb_monomials(5, 2)


def test_operator_lowering_of_the_shipped_operators():
    cloud = u.SquareCloud(Nx=9, Ny=7, facet_types=CONFIG1_FACETS)
    Ni = cloud.Ni
    lap, _ = u.lower_diff_operator(laplace_op(u), cloud, u.polyharmonic)
    assert np.array_equal(lap, np.tile([0, 0, 0, 1.0, 1.0], (Ni, 1)))                     # README.md:46-47
    adv, adv_p = u.lower_diff_operator(advdiff_op(u), cloud, u.polyharmonic)
    assert np.allclose(adv, np.tile([1e4, 100.0, 0.0, -0.08, -0.08], (Ni, 1))) and np.array_equal(adv, adv_p)
    # Navier-Stokes momentum: U . grad - lap / Re with fields = (u, v)   (demos/NavierStokes/30_...:61-65)
    uu, vv = np.linspace(0, 1, cloud.N), np.linspace(2, 3, cloud.N)

    def ns(x, center, rbf, monomial, fields):
        U = np.array([fields[0], fields[1]])
        return u.dot(U, u.nodal_gradient(x, center, rbf, monomial)) - u.nodal_laplacian(x, center, rbf, monomial) / 100.0
    c, _ = u.lower_diff_operator(ns, cloud, u.polyharmonic, [uu, vv])
    assert np.allclose(c[:, 1], uu[:Ni]) and np.allclose(c[:, 2], vv[:Ni]) and np.allclose(c[:, 3:], -0.01)
    # Darcy: -div(k grad) through nodal_div_grad with a per-row field   (demos/Darcy/00_darcy_flow.py:85-89)
    kk = np.linspace(1, 2, cloud.N)
    darcy = lambda x, center, rbf, monomial, fields: -u.nodal_div_grad(x, center, rbf, monomial, (fields[0], fields[0]))
    c, _ = u.lower_diff_operator(darcy, cloud, u.gaussian, [kk])
    assert np.allclose(c[:, 3], -kk[:Ni]) and np.allclose(c[:, 4], -kk[:Ni]) and np.all(c[:, :3] == 0)
    # coefficients may depend on x
    c, _ = u.lower_diff_operator(lambda x, ce, r, m, f: np.sin(x[0]) * u.nodal_value(x, ce, r, m), cloud, u.gaussian)
    assert np.allclose(c[:, 0], np.sin(cloud.sorted_nodes[:Ni, 0]))


@pytest.mark.parametrize("bad", [
    lambda x, c, r, m, f: u.nodal_value(x, c, r, m) * u.nodal_gradient(x, c, r, m)[0],       # demos/NavierStokes/10_...:89-94
    lambda x, c, r, m, f: u.nodal_value(x, c, r, m) ** 2,
    lambda x, c, r, m, f: 1.0 / u.nodal_laplacian(x, c, r, m),
    lambda x, c, r, m, f: u.nodal_value(x, c, r, m) + 1.0,
    lambda x, c, r, m, f: 3.0,
    lambda x, c, r, m, f: np.exp(u.nodal_value(x, c, r, m)),
    lambda x, c, r, m, f: r(x, c),
])
def test_operators_outside_the_term_set_raise(bad):
    cloud = u.SquareCloud(Nx=6, Ny=5, facet_types=CONFIG1_FACETS)
    with pytest.raises((u.OperatorLoweringError, TypeError)):
        u.lower_diff_operator(bad, cloud, u.polyharmonic)


def test_bc_preparation_matches_reference_rules():
    cloud = u.SquareCloud(Nx=8, Ny=6, facet_types={"South": "r", "West": "d", "North": "p1", "East": "n"})
    bcs = {"South": (lambda c: c[0], lambda c: 2.0 + c[0]), "West": lambda c: 1.0, "North": lambda c: 5.0, "East": 0.5}
    arr = u.boundary_conditions_func_to_arr(bcs, cloud)
    south = np.asarray(cloud.facet_nodes["South"])
    assert np.allclose(arr["South"][0], cloud.sorted_nodes[south, 0]) and np.allclose(arr["South"][1], 2 + cloud.sorted_nodes[south, 0])
    robin, new = u.duplicate_robin_coeffs(arr, cloud)
    assert sorted(robin) == sorted(south.tolist()) and np.allclose([robin[i] for i in south], 2 + cloud.sorted_nodes[south, 0])
    assert np.allclose(new["South"], cloud.sorted_nodes[south, 0])
    new = u.zerofy_periodic_cond(new, cloud)
    assert np.all(new["North"] == 0)
    with pytest.raises(ValueError):                                               # reference: AttributeError (Q6)
        u.duplicate_robin_coeffs({**arr, "South": np.zeros(len(south))}, cloud)


def test_row_descriptors_periodic_layout():
    """Boundary row order d, n, r, periodic-value rows of all groups, then periodic-flux rows (assembly.py:157-267)."""
    cloud = u.SquareCloud(Nx=7, Ny=7, facet_types=CONFIG2_FACETS)
    t = asm.build_operator_rows(cloud, np.tile([0, 0, 0, 1.0, 1.0], (cloud.Ni, 1)))
    Ni, (n0, n1) = cloud.Ni, cloud.Np
    half = (n0 + n1) // 2
    v0 = np.arange(Ni, Ni + n0 // 2)
    assert np.array_equal(t.p1[v0], v0) and np.array_equal(t.p2[v0], v0 + n0 // 2) and np.all(t.skip[v0] == -1)
    v1 = np.arange(Ni + n0 // 2, Ni + half)
    assert np.array_equal(t.p1[v1], np.arange(Ni + n0, Ni + n0 + n1 // 2))
    f0 = v0 + half
    assert np.array_equal(t.p1[f0], v0) and np.all(t.cphi1[f0, 0] == 0) and np.all(t.cphi1[v0, 0] == 1) and np.all(t.cphi2[v0, 0] == -1)
    assert np.array_equal(t.skip[:Ni], np.arange(Ni))
    assert t.masks() == (asm.JET_ISO, asm.JET_VAL | asm.JET_GRAD)


def test_c_abi_exports_every_declared_symbol():
    from updes_b200 import _lib
    header = open(os.path.join(ROOT, "include", "updes_b200.h")).read()
    declared = set(re.findall(r"\b(updes_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), "libupdes_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree: %s" % (declared ^ set(_lib.SIGNATURES))
    assert b"sm_100a" in lib.updes_b200_version()


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cloud = u.SquareCloud(Nx=6, Ny=5, facet_types=CONFIG1_FACETS)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        u.pde_solver(laplace_op(u), lambda x, c, r, f: 0.0, cloud, {k: (lambda c: 0.0) for k in cloud.facet_types},
                     u.polyharmonic, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        u.value(np.zeros(2), np.zeros(33), cloud.sorted_nodes, u.polyharmonic)


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "updes_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs the CPU oracle port and prints one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-nx", "16", "--cpu-sizes", "10x8,14x14,20x20"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
    sweep = line["cpu_baseline"]["size_sweep"]               # measured sizes + cubic extrapolation, labelled
    assert [r["n"] for r in sweep["measured"]] == [83, 199, 403] and "EXTRAPOLATED" in sweep["note"]
    assert sweep["extrapolated_seconds_at_n_90003"] > sweep["measured"][-1]["seconds"]


def test_bench_reference_arm_under_torchrun_env_restores_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must lift it (rank 0 re-executes itself) and the other
    ranks must exit 0 without output (round 1: the N >= 2 reference arms ran on one thread)."""
    import json
    import subprocess
    import sys
    base = dict(os.environ, OMP_NUM_THREADS="1", WORLD_SIZE="2", LOCAL_RANK="0")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
           "--cpu-nx", "12", "--no-cpu-sweep"]
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(base, RANK="1"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    r0 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(base, RANK="0"))
    assert r0.returncode == 0, r0.stderr[-2000:]
    line = json.loads(r0.stdout.strip().splitlines()[-1])
    assert line["config"]["host_threads_env"] == str(os.cpu_count()) and line["n_gpus"] == 2
    assert line["cpu_baseline"]["cores"] == os.cpu_count() or os.cpu_count() == 1


def test_c_abi_rejects_bad_arguments_without_a_gpu():
    """Argument validation comes before any CUDA call: negative status = index of the offending argument."""
    import ctypes
    from updes_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.updes_lu_create(ctypes.byref(h), 100, 7) == -3            # ld not a multiple of 16
    assert lib.updes_lu_create(ctypes.byref(h), 0, 16) == -2             # n <= 0
    assert lib.updes_lu_create(None, 10, 16) == -1
    assert lib.updes_lu_factor(None, None, None, None, None) == -1
    assert lib.updes_lu_solve(None, None, None, None, 0, 1, 0, None) == -1
    rows = _lib.UpdesRows()
    assert lib.updes_assemble_rows(0, 1.0, 0, 3, None, ctypes.byref(rows), 0, 1, 7, None, 16, None) == -3   # N <= 0
    assert lib.updes_assemble_rows(0, 1.0, 10, 99, None, ctypes.byref(rows), 0, 1, 7, None, 16, None) == -4  # M > 15
    assert lib.updes_assemble_rows(0, 1.0, 10, 3, None, ctypes.byref(rows), 0, 1, 7, None, 16, None) == -5   # no centres
    assert lib.updes_eval_jets(0, 1.0, 0, 3, None, None, 0, 1, None, 1, None, None, None, None, None) == -3
    assert lib.updes_lu_set_gemm_variant(None, 0) == -1


# ---- round 2: boundary semantics ------------------------------------------------------------------
def test_steadysol_behaves_like_the_reference_namedtuple():
    """utils.py:148: SteadySol = namedtuple('PDESolution', ['vals', 'coeffs', 'mat']); mat is lazy here."""
    calls = []
    sol = u.SteadySol(np.arange(3.0), np.arange(4.0), None, lambda: calls.append(1) or np.eye(2))
    assert sol._fields == ("vals", "coeffs", "mat") and len(sol) == 3
    assert np.array_equal(sol[0], np.arange(3.0)) and np.array_equal(sol[1], np.arange(4.0)) and np.array_equal(sol[-3], sol.vals)
    assert calls == []                                   # nothing computed so far
    v, c, m = sol                                         # unpacking reads mat
    assert np.array_equal(m, np.eye(2)) and calls == [1]
    assert np.array_equal(sol[2], np.eye(2)) and np.array_equal(sol.mat, np.eye(2)) and calls == [1]    # cached
    sol2 = sol._replace(vals=np.zeros(3))
    assert np.array_equal(sol2.vals, np.zeros(3)) and sol2.coeffs is sol.coeffs and np.array_equal(sol2.mat, np.eye(2))
    assert set(sol._asdict()) == {"vals", "coeffs", "mat"}
    with pytest.raises(IndexError):
        sol[3]
    with pytest.raises(ValueError):
        sol._replace(values=1)
    assert len(sol[0:2]) == 2
    # a lazy mat stays lazy through _replace, and eager construction works as on the namedtuple
    lazy = u.SteadySol(1, 2, None, lambda: calls.append(2) or 7)._replace(coeffs=3)
    assert calls == [1] and lazy.mat == 7 and calls == [1, 2]
    assert u.SteadySol(1, 2, 3).mat == 3


def test_operator_with_per_node_python_control_flow_falls_back_to_row_evaluation():
    """The reference vmaps operators over nodes (assembly.py:126-130): x is a (2,) point there.  A Python
    `if` on a coordinate cannot run on the batched (2, Ni) coordinates; it is evaluated node by node."""
    cloud = u.SquareCloud(Nx=9, Ny=7, facet_types=CONFIG1_FACETS)
    seen_shapes = set()

    def op(x, center, rbf, monomial, fields):
        seen_shapes.add(np.shape(x))
        k = 2.0 if x[0] > 0.5 else 1.0                      # ambiguous for an array -> ValueError in batch mode
        return k * u.nodal_laplacian(x, center, rbf, monomial) + float(fields[0]) * u.nodal_value(x, center, rbf, monomial)

    f = np.linspace(1, 2, cloud.N)
    cphi, cpol = u.lower_diff_operator(op, cloud, u.polyharmonic, [f])
    xs = cloud.sorted_nodes[:cloud.Ni, 0]
    want = np.where(xs > 0.5, 2.0, 1.0)
    assert np.array_equal(cphi[:, 3], want) and np.array_equal(cphi[:, 4], want) and np.allclose(cphi[:, 0], f[:cloud.Ni])
    assert np.array_equal(cphi, cpol)
    assert (2,) in seen_shapes and (2, cloud.Ni) in seen_shapes
    # an operator that is wrong in BOTH modes still raises the explicit lowering error
    with pytest.raises(u.OperatorLoweringError):
        u.lower_diff_operator(lambda x, c, r, m, fl: (1.0 if float(x[0]) > 0 else 2.0) * u.nodal_value(x, c, r, m) ** 2,
                              cloud, u.polyharmonic)


def test_diff_args_of_coefficient_length_and_bad_lengths():
    cloud = u.SquareCloud(Nx=8, Ny=6, facet_types=CONFIG1_FACETS)
    op = lambda x, c, r, m, f: f[0] * u.nodal_value(x, c, r, m)
    long_field = np.arange(cloud.N + 3, dtype=float)         # a coefficient vector (N+M): the reference indexes fields[i]
    c, _ = u.lower_diff_operator(op, cloud, u.polyharmonic, [long_field])
    assert np.array_equal(c[:, 0], long_field[:cloud.Ni])
    with pytest.raises(ValueError):
        u.lower_diff_operator(op, cloud, u.polyharmonic, [np.zeros(cloud.Ni - 1)])


def test_points_layouts_inside_and_outside_operators():
    from updes_b200 import operators as ops
    pts, lay = ops._points(np.array([0.1, 0.2]))
    assert lay == "single" and pts.shape == (1, 2)
    pts, lay = ops._points(np.zeros((7, 2)))
    assert lay == "rows" and pts.shape == (7, 2)
    x = ops.BatchPoints(np.arange(10.0).reshape(2, 5))
    pts, lay = ops._points(x)
    assert lay == "batch" and pts.shape == (5, 2) and np.array_equal(pts[:, 0], np.arange(5.0))
    rebuilt = np.stack([x[0], x[1]])                      # the usual port of jnp.array([x[0], x[1]]): loses the type
    with pytest.raises(ValueError):
        ops._points(rebuilt)                              # outside an operator call a (2, 5) array is ambiguous -> refuse
    with ops._batch_rows(5):
        pts, lay = ops._points(rebuilt)
        assert lay == "batch" and np.array_equal(pts[:, 1], np.arange(5.0, 10.0))
    with pytest.raises(ValueError):
        ops._points(np.zeros(3))


def test_interpolate_field_is_the_reference_permutation():
    """updes/tests/test_interpolation.py:60-63 restated: two clouds on the same grid with swapped boundary types."""
    f1 = {"South": "d", "West": "d", "North": "n", "East": "d"}
    f2 = {"South": "n", "West": "n", "North": "d", "East": "n"}
    c1 = u.SquareCloud(Nx=8, Ny=8, facet_types=f1)
    c2 = u.SquareCloud(Nx=8, Ny=8, facet_types=f2)
    field1 = np.sin(3 * c1.sorted_nodes[:, 0]) + c1.sorted_nodes[:, 1] ** 2
    field2 = u.interpolate_field(field1, c1, c2)
    assert np.allclose(field2, np.sin(3 * c2.sorted_nodes[:, 0]) + c2.sorted_nodes[:, 1] ** 2, atol=1e-12)
    assert np.array_equal(u.interpolate_field(field2, c2, c1), field1)
    vec = np.stack([field1, -field1], axis=-1)             # (N, 2) fields are carried row-wise (gradphi in the NS demo)
    assert np.array_equal(u.interpolate_field(vec, c1, c2)[:, 1], -field2)
    # the committed parse of the reference's mesh fixture keeps its original numbering too
    from helpers import cloud_from_golden
    cv, _ = cloud_from_golden("mesh_msh_cloud_vel.npz")
    cp, _ = cloud_from_golden("mesh_msh_cloud_phi.npz")
    xv = u.interpolate_field(cp.sorted_nodes, cp, cv)
    assert np.array_equal(xv, cv.sorted_nodes)


def test_config3_operators_lower_to_the_reference_rows():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import configs
    from helpers import cloud_from_golden
    cv, _ = cloud_from_golden("mesh_msh_cloud_vel.npz")
    du, ru, dv, rv, dphi, rphi = configs.config3_operators(u, Re=100.0)
    uu, vv = np.cos(cv.sorted_nodes[:, 0]), np.sin(cv.sorted_nodes[:, 1])
    c, cp = u.lower_diff_operator(du, cv, u.polyharmonic, [uu, vv])
    assert np.array_equal(c[:, 1], uu[:cv.Ni]) and np.array_equal(c[:, 2], vv[:cv.Ni]) and np.all(c[:, 3:] == -0.01) and np.all(c[:, 0] == 0)
    assert np.array_equal(c, cp)
    c, _ = u.lower_diff_operator(dphi, cv, u.polyharmonic, None)
    assert np.array_equal(c, np.tile([0, 0, 0, 1.0, 1.0], (cv.Ni, 1)))


def test_distributed_path_is_opt_in():
    """An initialised process group alone must not switch pde_solver to the sharded path."""
    from updes_b200 import operators as ops
    assert ops._dist_world() == (1, 0, None)
    with pytest.raises(RuntimeError):
        u.enable_distributed()                           # no process group
    with pytest.raises(RuntimeError):
        ops._dist_world(distributed=True)


def test_distributed_layout_selection(monkeypatch):
    """enable_distributed(grid=(P, Q)): P > 1 builds the 2-D system, everything else the measured 1 x Q system;
    the builder passes the row table through and the predicted footprints are sane (no GPU: stub systems)."""
    import torch.distributed as dist
    from updes_b200 import operators as ops
    made = []

    class Stub1:
        def __init__(self, cloud, kind, param, M, table, world, rank, group=None):
            made.append(("1xQ", table, world, rank, group))
        predict_nbytes = staticmethod(ops._DistSystem.predict_nbytes)

    class Stub2:
        def __init__(self, cloud, kind, param, M, table, grid, rank, group=None):
            made.append(("PxQ", table, grid, rank, group))
        predict_nbytes = staticmethod(ops._DistSystem2D.predict_nbytes)

    monkeypatch.setattr(ops, "_DistSystem", Stub1)
    monkeypatch.setattr(ops, "_DistSystem2D", Stub2)
    monkeypatch.setattr(dist, "is_initialized", lambda: True)
    monkeypatch.setattr(dist, "get_world_size", lambda group=None: 4)
    monkeypatch.setattr(dist, "get_rank", lambda group=None: 3)
    cloud = u.SquareCloud(Nx=8, Ny=8, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})
    try:
        ops.enable_distributed()
        assert ops._dist_world() == (4, 3, None) and ops._DIST["grid"] is None
        build, need = ops._make_dist_system(cloud, "polyharmonic", 1.0, 3, lambda: "TABLE", 4, 3, None)
        build()
        assert made[-1] == ("1xQ", "TABLE", 4, 3, None) and need == ops._DistSystem.predict_nbytes(cloud.N + 3, 4)
        ops.enable_distributed(grid=(1, 4))                     # 1 x Q spelled as a grid is still the default path
        assert ops._DIST["grid"] is None
        ops.enable_distributed(grid=(2, 2))
        build, need = ops._make_dist_system(cloud, "polyharmonic", 1.0, 3, lambda: "TABLE", 4, 3, None)
        build()
        assert made[-1] == ("PxQ", "TABLE", (2, 2), 3, None) and need == ops._DistSystem2D.predict_nbytes(cloud.N + 3, (2, 2))
        with pytest.raises(ValueError):
            ops.enable_distributed(grid=(3, 2))                 # 6 != 4 ranks
        # footprints at the sizes the layouts are for: 250k nodes on 8 GPUs is ~62.5 GB of matrix per rank either way
        n = 250003
        assert 62e9 < ops._DistSystem.predict_nbytes(n, 8) < 75e9
        assert 62e9 < ops._DistSystem2D.predict_nbytes(n, (2, 4)) < 85e9   # + gathered panel, panel rows, U12 and row-exchange staging
    finally:
        ops.disable_distributed()
    assert ops._DIST == {"enabled": False, "group": None, "grid": None}


def test_numeric_probe_lowering_of_operators_that_leave_the_symbolic_world():
    """An operator whose body needs numbers (float(), np.asarray, a foreign array library) cannot be traced symbolically;
    lower_diff_operator then probes it numerically row by row: unit jets in, coefficients out, linearity checked."""
    cloud = u.SquareCloud(Nx=8, Ny=6, facet_types=CONFIG1_FACETS)
    Ni = cloud.Ni
    f0, f1 = np.linspace(1, 2, cloud.N), np.linspace(-1, 1, cloud.N)

    def op(x, center, rbf, monomial, fields):
        g = np.asarray(u.nodal_gradient(x, center, rbf, monomial), dtype=float)        # forces numbers
        lap = float(u.nodal_laplacian(x, center, rbf, monomial))
        dg = u.nodal_div_grad(x, center, rbf, monomial, (fields[0], 2.0))
        return float(fields[1]) * float(u.nodal_value(x, center, rbf, monomial)) + float(np.dot([x[0], -0.5], g)) - 0.25 * lap + dg

    cphi, cpol = u.lower_diff_operator(op, cloud, u.polyharmonic, [f0, f1])
    xs = cloud.sorted_nodes[:Ni, 0]
    want = np.stack([f1[:Ni], xs, np.full(Ni, -0.5), f0[:Ni] - 0.25, np.full(Ni, 2.0 - 0.25)], axis=1)
    assert np.allclose(cphi, want, rtol=1e-15, atol=0) and np.array_equal(cphi, cpol)
    # the same machinery rejects what the symbolic path rejects
    for bad in (lambda x, c, r, m, f: float(u.nodal_value(x, c, r, m)) ** 2,
                lambda x, c, r, m, f: float(u.nodal_value(x, c, r, m)) + 1.0,
                lambda x, c, r, m, f: float(u.nodal_value(x, c, r, m)) * float(u.nodal_laplacian(x, c, r, m))):
        with pytest.raises(u.OperatorLoweringError):
            u.lower_diff_operator(bad, cloud, u.polyharmonic)
    # and the term set is still unavailable outside an operator
    with pytest.raises(u.OperatorLoweringError):
        u.nodal_value(np.zeros(2), np.zeros(2), u.polyharmonic, None)


def test_numeric_probe_batch_result_is_validated_against_per_node_evaluation():
    """A body that forces numbers and reduces over 'all axes' means one node's two gradient components in the
    reference's vmap semantics, but would sum over every row on batched arrays: the batched table is discarded."""
    cloud = u.SquareCloud(Nx=7, Ny=6, facet_types=CONFIG1_FACETS)
    op = lambda x, c, r, m, f: np.sum(np.asarray(u.nodal_gradient(x, c, r, m), dtype=float)) + 0.0 * np.sum(x)
    cphi, _ = u.lower_diff_operator(op, cloud, u.polyharmonic)
    assert np.array_equal(cphi, np.tile([0.0, 1.0, 1.0, 0.0, 0.0], (cloud.Ni, 1)))


def test_batched_symbolic_lowering_is_validated_against_per_node_evaluation():
    """np.sum(x) is x + y of ONE node in the reference's semantics; on the batched (2, Ni) coordinates it would silently
    become the sum over all rows.  Sample rows are re-evaluated per node and the batched table is discarded."""
    cloud = u.SquareCloud(Nx=7, Ny=6, facet_types=CONFIG1_FACETS)
    op = lambda x, c, r, m, f: np.sum(x) * u.nodal_value(x, c, r, m) + u.nodal_laplacian(x, c, r, m)
    cphi, cpol = u.lower_diff_operator(op, cloud, u.polyharmonic)
    xy = cloud.sorted_nodes[:cloud.Ni]
    assert np.allclose(cphi[:, 0], xy[:, 0] + xy[:, 1], rtol=1e-15, atol=0) and np.all(cphi[:, 3:] == 1.0) and np.array_equal(cphi, cpol)


def test_explicit_assembly_names_exist_and_refuse_large_clouds():
    """The reference's assemble_* names are exported (updes/assembly.py:10-401); explicit host matrices are refused
    beyond N = 20 000 before anything touches the GPU."""
    for name in ("assemble_Phi", "assemble_P", "assemble_A", "assemble_invert_A", "assemble_op_Phi_P", "assemble_bd_Phi_P",
                 "assemble_B", "assemble_q", "core_compute_coefficients", "compute_coefficients", "get_field_coefficients"):
        assert callable(getattr(u, name)), name

    class Big:
        N = 20001
    for fn, args in ((u.assemble_A, (Big, u.polyharmonic, 3)), (u.assemble_invert_A, (Big, u.polyharmonic, 3)),
                     (u.assemble_bd_Phi_P, (Big, u.polyharmonic, 3)), (u.assemble_B, (None, Big, u.polyharmonic, 3, None, {}))):
        with pytest.raises(MemoryError):
            fn(*args)
    from updes_b200.explicit import _betas
    cloud = u.SquareCloud(Nx=6, Ny=5, facet_types={"South": "r", "West": "d", "North": "d", "East": "n"})
    ids = cloud.facet_nodes["South"]
    assert np.array_equal(_betas(cloud, {i: 1.0 + k for k, i in enumerate(reversed(ids))}), np.arange(len(ids), 0, -1.0))
    assert np.array_equal(_betas(cloud, None), np.zeros(cloud.Nr))
    with pytest.raises(ValueError):
        _betas(cloud, {ids[0]: 1.0})


def test_cloud_plotting_helpers_of_the_reference_surface():
    """visualize_cloud / visualize_normals / visualize_field / animate_fields / average_spacing exist with the reference's signatures
    (cloud.py:62-70, :175-285; the README example ends with cloud.visualize_field(...)).  Host-only: without matplotlib
    they raise ImportError; with a pyplot module they draw (here: the no-op stand-in of oracle/refshim, in a subprocess)."""
    import importlib.util
    import subprocess
    cloud = u.SquareCloud(Nx=5, Ny=4, facet_types=CONFIG1_FACETS)
    xy = cloud.sorted_nodes
    d = np.sqrt(((xy[:, None] - xy[None]) ** 2).sum(-1))
    assert np.isclose(cloud.average_spacing(), d[np.triu_indices(cloud.N)].mean(), rtol=1e-15)
    if importlib.util.find_spec("matplotlib") is None:
        with pytest.raises(ImportError):
            cloud.visualize_field(np.zeros(cloud.N))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, updes_b200 as u\n"
            "c = u.SquareCloud(Nx=6, Ny=5, facet_types={'South': 'n', 'West': 'd', 'North': 'r', 'East': 'd'})\n"
            "c.visualize_cloud(s=6); c.visualize_normals()\n"
            "ax, img = c.visualize_field(np.arange(c.N, dtype=float), cmap='jet', projection='3d', title='RBF solution')\n"
            "ax, img = c.visualize_field(np.arange(c.N, dtype=float)[:, None], levels=20)\n"
            "hist = [np.full(c.N, float(k)) for k in range(4)]\n"
            "axes = c.animate_fields([hist, np.stack(hist)], cmaps='jet', filename='x.gif', titles=['a', 'b'], vmin=0, vmax=1)\n"
            "assert len(axes) == 2 and len(c.animate_fields([hist], titles=['only'])) == 1\n"
            "print('drawn')\n" % (root, os.path.join(root, "oracle", "refshim")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "drawn" in r.stdout, r.stderr[-1500:]


def test_small_host_utilities_of_the_reference_surface(tmp_path):
    """make_dir / random_name / dot_vec / dot_mat / RK4 (updes/utils.py:152-246, operators.py:777-779): every demo script of
    the reference calls some of them next to the solver."""
    d = str(tmp_path / "run")
    u.make_dir(d); u.make_dir(d)
    assert os.path.isdir(d) and len(u.random_name(7)) == 7 and u.random_name().isdigit()
    a, b = np.arange(6.0).reshape(3, 2), np.ones((3, 2))
    assert np.array_equal(u.dot_vec(a, b), a.sum(1)) and np.array_equal(u.dot_mat(np.tile(2 * np.eye(2), (3, 1, 1)), a), 2 * a)
    te = np.linspace(0, 1, 11)
    ys = u.RK4(lambda t, y: -y, (0.0, 1.0), np.array([1.0, 3.0]), t_eval=te, subdivisions=4)
    assert ys.shape == (11, 2) and np.allclose(ys[:, 0], np.exp(-te), rtol=1e-7) and np.array_equal(ys[0], [1.0, 3.0])
    assert u.RK4(lambda t, y: -y, (0.0, 1.0), np.array([1.0])).shape == (2, 1)
    with pytest.raises(ValueError):
        u.RK4(lambda t, y: -y, (0.0, None), np.array([1.0]))


def test_support_size_selects_only_the_global_path():
    """The reference turns 'max' into N (cloud.py:97-98) and then drops the node itself; N - 1 already drops the farthest
    node of every support (RBF-FD), which is out of scope.  Such a cloud can be BUILT (geometry and numbering do not depend
    on the supports; the reference's own test_interpolation.py builds one and only permutes fields) but assembling on it
    is refused, not silently treated as global."""
    ft = {"South": "n", "West": "d", "North": "n", "East": "d"}
    for ss in ("max", None, 64):
        c = u.SquareCloud(Nx=8, Ny=8, facet_types=ft, support_size=ss)
        assert c.N == 64 and c.support_size == 64 and len(c.local_supports[5]) == 63
    full = u.SquareCloud(Nx=8, Ny=8, facet_types=ft)
    for ss in (63, 10):
        c = u.SquareCloud(Nx=8, Ny=8, facet_types=ft, support_size=ss)
        assert c.support_size == ss and np.array_equal(c.sorted_nodes, full.sorted_nodes) and c.facet_nodes == full.facet_nodes
        assert c.local_supports[5] == full.local_supports[5][:ss - 1] and c.sorted_local_supports.shape == (64, ss - 1)
        with pytest.raises(NotImplementedError):
            asm.DeviceRows(c, asm.build_interpolation_rows(c))              # refused before any device work
        with pytest.raises(NotImplementedError):
            u.pde_solver_jit(laplace_op(u), lambda x, centers, rbf, fields: 0.0, c, {k: (lambda p: 0.0) for k in ft}, u.polyharmonic, 1)
    with pytest.raises(ValueError):
        u.SquareCloud(Nx=8, Ny=8, facet_types=ft, support_size="all")
    with pytest.raises(AssertionError):
        u.SquareCloud(Nx=8, Ny=8, facet_types=ft, support_size=65)          # cloud.py:99-100


def test_symbolic_and_numeric_lowering_agree_on_random_linear_operators():
    """Random operators a0 phi + a1 phi_x + a2 phi_y + a3 phi_xx + a4 phi_yy with coefficients mixing constants, x and
    fields: the symbolic path and the numeric-probe path (forced by converting the terms to numbers) give the same table."""
    rng = np.random.default_rng(5)
    cloud = u.SquareCloud(Nx=7, Ny=6, facet_types=CONFIG1_FACETS)
    Ni = cloud.Ni
    F = rng.normal(size=(3, cloud.N))
    for trial in range(8):
        w = rng.normal(size=(5, 3))                       # weights of (1, x[0], fields[trial % 3]) per coefficient

        def coeffs(x, f, k):
            return w[k, 0] + w[k, 1] * x[0] + w[k, 2] * f[trial % 3]

        def symbolic(x, c, r, m, f):
            g = u.nodal_gradient(x, c, r, m)
            return (coeffs(x, f, 0) * u.nodal_value(x, c, r, m) + coeffs(x, f, 1) * g[0] + coeffs(x, f, 2) * g[1]
                    + u.nodal_div_grad(x, c, r, m, (coeffs(x, f, 3), coeffs(x, f, 4))))

        def numeric(x, c, r, m, f):
            g = np.asarray(u.nodal_gradient(x, c, r, m), dtype=float)             # leaves the symbolic world
            return (coeffs(x, f, 0) * float(u.nodal_value(x, c, r, m)) + coeffs(x, f, 1) * g[0] + coeffs(x, f, 2) * g[1]
                    + u.nodal_div_grad(x, c, r, m, (coeffs(x, f, 3), coeffs(x, f, 4))))

        a, ap = u.lower_diff_operator(symbolic, cloud, u.gaussian, list(F))
        b, bp = u.lower_diff_operator(numeric, cloud, u.gaussian, list(F))
        xs = cloud.sorted_nodes[:Ni, 0]
        want = np.stack([w[k, 0] + w[k, 1] * xs + w[k, 2] * F[trial % 3, :Ni] for k in range(5)], axis=1)
        assert np.allclose(a, want, rtol=1e-14, atol=1e-15) and np.allclose(b, want, rtol=1e-14, atol=1e-15)
        assert np.array_equal(a, ap) and np.array_equal(b, bp)


def test_robin_betas_follow_the_reference_offset_rule_when_facets_interleave():
    """Q8 (operators.py:535-536): betas[i - node_ids[0]] with the jax out-of-bounds clamp.  North owns the corners, so
    with East also Robin its last node is numbered after East's nodes: offset > len - 1 -> the last beta."""
    import warnings
    cloud = u.SquareCloud(Nx=5, Ny=4, facet_types={"South": "d", "West": "d", "North": "r", "East": "r"})
    north, east = cloud.facet_nodes["North"], cloud.facet_nodes["East"]
    assert north[-1] - north[0] > len(north) - 1 and east[-1] - east[0] == len(east) - 1      # North interleaved, East contiguous
    bn, be = np.arange(10.0, 10.0 + len(north)), np.arange(20.0, 20.0 + len(east))
    bcs = {"South": 0.0, "West": 0.0, "North": (np.zeros(len(north)), bn), "East": (np.zeros(len(east)), be)}
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        robin, new_bc = u.duplicate_robin_coeffs(bcs, cloud)
    assert any("not numbered contiguously" in str(m.message) for m in w)
    for k, i in enumerate(east):
        assert robin[i] == be[k]
    for i in north:
        assert robin[i] == bn[min(i - north[0], len(north) - 1)]
    assert robin[north[-1]] == bn[-1] and set(robin) == set(north) | set(east)
    assert new_bc["North"] is bcs["North"][0] and new_bc["South"] == 0.0


def test_star_import_gives_a_reference_script_the_names_it_takes_from_updes():
    """The reference's demos start with `from updes import *` and use, beside the solver surface, names that updes/utils.py
    merely imports (partial, Partial, os) and its small helpers (plot, dataloader, make_dir, RK4 ...)."""
    ns = {}
    exec("from updes_b200 import *", ns)
    for name in ("SquareCloud", "GmshCloud", "pde_solver", "pde_solver_jit", "pde_multi_solver", "polyharmonic", "gaussian",
                 "nodal_value", "nodal_gradient", "nodal_laplacian", "nodal_div_grad", "value", "gradient", "laplacian",
                 "divergence", "gradient_vec", "interpolate_field", "integrate_field", "get_field_coefficients",
                 "partial", "Partial", "os", "make_dir", "random_name", "RK4", "plot", "dataloader"):
        assert name in ns, name
    assert ns["partial"](ns["polyharmonic"], a=2).keywords == {"a": 2} and ns["Partial"] is ns["partial"]
    data = np.arange(20).reshape(10, 2)
    batches = list(u.dataloader(data, 3, 5))
    assert [b.shape for b in batches] == [(3, 2)] * 3                         # `while end < dataset_size`: 3, 6, 9 -- never the rest
    rows = np.concatenate(batches)
    assert len({tuple(r) for r in rows}) == 9 and all(tuple(r) in {tuple(d) for d in data} for r in rows)
    assert [b.tolist() for b in u.dataloader(data, 3, 5)] == [b.tolist() for b in batches]   # same key, same batches
    assert list(u.dataloader(data, 10, 0)) == []


def test_global_indices_hold_the_renumbered_ids_and_radial_profiles_match_the_kernels(capsys):
    """cloud.py:153-157: after renumbering, global_indices[k, l] is the NEW id of grid node (k, l) (and global_indices_rev
    its inverse); print_global_indices lays them out as drawn.  utils.py:30-89: the *_func radial profiles are what the
    two-point kernels apply to the distance."""
    c = u.SquareCloud(Nx=6, Ny=4, facet_types={"South": "p1", "West": "r", "North": "p1", "East": "n"})
    xs, ys = np.linspace(0, 1, 6), np.linspace(0, 1, 4)
    for k in range(6):
        for l in range(4):
            assert np.allclose(c.sorted_nodes[c.global_indices[k, l]], (xs[k], ys[l]), rtol=0, atol=1e-15)
            assert c.global_indices_rev[int(c.global_indices[k, l])] == (k, l)
    assert sorted(c.global_indices.ravel().tolist()) == list(range(c.N))
    c.print_global_indices()
    assert str(int(c.global_indices[0, 3])) in capsys.readouterr().out.splitlines()[0]          # top-left of the drawing
    x, ctr = np.array([0.3, 0.4]), np.zeros(2)
    assert u.polyharmonic(x, ctr, a=2) == u.polyharmonic_func(0.5, 2) and u.thin_plate(x, ctr, a=2) == u.thin_plate_func(0.5, 2)
    assert u.gaussian(x, ctr, eps=3.0) == u.gaussian_func(0.5, 3.0) and u.multiquadric(x, ctr, eps=3.0) == u.multiquadric_func(0.5, 3.0)
    assert u.inverse_multiquadric(x, ctr, eps=3.0) == u.inv_multiquadric_func(0.5, 3.0) and u.thin_plate_func(0.0, 1) == 0.0
    assert u.make_nodal_rbf(x, ctr, lambda r: 2 * r) == 1.0 and u.value_vec_ is u.value and u.gradient_vec_ is u.gradient
