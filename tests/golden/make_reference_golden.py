"""Golden vectors produced by the REFERENCE'S OWN CODE (run from the repo root in the build container, where
/root/reference is mounted):

    python tests/golden/make_reference_golden.py [--only NAME] [--out DIR]

`import updes` below imports the unmodified reference package from /root/reference.  Its third-party dependencies
(jax, lineax, matplotlib, seaborn) are absent from the image; oracle/refshim/ provides stand-ins for exactly the
API surface the hot path touches, on torch float64 (see oracle/refshim/README.md for what that does and does not
pin).  Every array written here is the return value of a reference function -- SquareCloud / GmshCloud,
assemble_A, assemble_op_Phi_P, assemble_bd_Phi_P, assemble_q, pde_solver_jit -- called the way the reference's
README, demos and tests call them.

Files (tests/golden/ref_*.npz), each with the cloud arrays and:
  ref_laplace_12x9        README problem (config 1 shape): A, opPhi, opP, bdPhi, bdP, q, B (= sol.mat), vals, coeffs
  ref_robin_11x8          Neumann + Robin facets together (quirk Q3), gaussian eps=3, degree 2, operator with value,
                          gradient and Laplacian terms, Robin (value, beta) tuples with callables
  ref_periodic_10x10      config 2 shape: doubly periodic advection-diffusion, one implicit step with rhs = value(u)/DT
  ref_kernels_7x6         all five kernels with max_degree 4 (15 monomials) and a field-dependent operator that
                          uses every term of the set incl. nodal_div_grad; field evaluators value / gradient / laplacian
  ref_config1_30x20       config 1 at full size: q, vals, coeffs, a row sample of B
  ref_fuzz_16             16 random small problems: facet types incl. periodic pairs in random dict order, Robin + Neumann mixes,
                          all kernels and degrees 0-4, fully general operator with five nodal fields: diffMat of each
  ref_generated_msh       GmshCloud on two channel meshes written by tests/golden/make_msh.py, two facet-type orders
  ref_multi_solver_9x8    pde_multi_solver on two genuinely coupled equations, the state after each of three sweeps
  ref_integrals_12x12     the setting of updes/tests/test_integrals.py: coefficients of s = x^2/(1+y^2), the rebuilt field, integrate_field
  ref_laplace_demo_30x30  demos/Laplace/00_laplace_with_rbf.py run unmodified as a whole script: solution, Laplacian at the nodes, its printed errors
  ref_darcy_demo_20x20    demos/Darcy/00_darcy_flow.py run unmodified: identity-operator solve (polyharmonic a=2) and -div(k grad u) = 1 (thin_plate a=3)
  ref_config2_advdiff_3steps  config 2: the Advection demo's own definitions (35x35 periodic cloud, operators, u0), three time steps
  ref_advection00_2steps  demos/Advection/00_advection_with_rbf.py: its definitions (40x20, d/d/d/n, u0 from cloud.local_supports), two steps
  ref_advection02_sink_2steps  demos/Advection/02_adv_diff_periodic_with_sink.py: its definitions (sink field through diff_args), two steps
  ref_grayscott001_2steps demos/Gray-Scott/001_gray-scott.py: periodic ids 'p0' / 'p1' with degree 1, two steps
  ref_wave00_2steps       demos/Wave/00_wave.py: all-Neumann cloud, polyharmonic a=3, degree 2, two nodal fields in the rhs, two steps
  ref_cartesian_helpers   cartesian_gradient(_vec), enforce_cartesian_gradient_neumann, apply_neumann_conditions on three small clouds
  ref_config3_ns_2iter    config 3: two iterations of the demo's own projection loop (u, v, phi solves on the two mesh.msh clouds)
  ref_mesh_msh_{vel,phi}  the reference's fixture updes/tests/data/mesh.msh through GmshCloud for the two facet-type
                          sets of demos/NavierStokes/30_...:40-41; for phi also a row sample of bdPhi / bdP (Neumann
                          rows with the reference's computed normals)
"""
import argparse
import os
import sys
import time
import warnings
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
sys.path.insert(0, REFERENCE)
warnings.filterwarnings("ignore", message="The use of `x.T` on tensors")

import jax.numpy as jnp  # noqa: E402  (the stand-in)
import updes  # noqa: E402  (the reference itself)

assert os.path.realpath(updes.__file__).startswith(REFERENCE), "must import the reference package, found %s" % updes.__file__


def npa(t):
    return np.asarray(t.detach().numpy() if hasattr(t, "detach") else t)


def cloud_arrays(c):
    """Same keys as tests/golden/make_golden.py so tests/helpers.py:cloud_from_golden can rebuild the cloud."""
    names = list(c.facet_nodes.keys())
    normals = npa(c.sorted_outward_normals) if hasattr(c, "sorted_outward_normals") else np.zeros((0, 2))
    return dict(sorted_nodes=npa(c.sorted_nodes), sorted_outward_normals=normals,
                counts=np.array([c.N, c.Ni, c.Nd, c.Nn, c.Nr]), Np=np.array(list(c.Np), dtype=np.int64),
                facet_names=np.array(names), facet_sizes=np.array([len(c.facet_nodes[k]) for k in names]),
                facet_nodes=np.concatenate([np.asarray(c.facet_nodes[k], dtype=np.int64) for k in names]),
                facet_types=np.array([c.facet_types[k] for k in names]),
                old_of_new=np.array([o for o, _ in sorted(c.renumbering_map.items(), key=lambda kv: kv[1])], dtype=np.int64),
                supports_first_rows=npa(c.sorted_local_supports[:4]))


def blocks(diff_operator, cloud, rbf, max_degree, diff_args, robin_coeffs):
    """The four blocks of diffMat exactly as assemble_B builds them (assembly.py:384-385), plus A (:62-85)."""
    M = updes.compute_nb_monomials(max_degree, cloud.dim)
    opPhi, opP = updes.assemble_op_Phi_P(diff_operator, cloud, rbf, M, diff_args)
    bdPhi, bdP = updes.assemble_bd_Phi_P(cloud, rbf, M, robin_coeffs)
    A = updes.assemble_A(cloud, rbf, M)
    return dict(opPhi=npa(opPhi), opP=npa(opP), bdPhi=npa(bdPhi), bdP=npa(bdP), A=npa(A), M=np.array(M))


def solve(diff_operator, rhs_operator, cloud, bcs, rbf, max_degree, diff_args=None, rhs_args=None):
    """pde_solver_jit as a user calls it (operators.py:650-683), plus q and the Robin coefficients it derives."""
    sol = updes.pde_solver_jit(diff_operator=diff_operator, rhs_operator=rhs_operator, cloud=cloud, boundary_conditions=bcs,
                               rbf=rbf, max_degree=max_degree, diff_args=diff_args, rhs_args=rhs_args)
    bc_arr = updes.boundary_conditions_func_to_arr(bcs, cloud)
    robin, bc_arr = updes.duplicate_robin_coeffs(bc_arr, cloud)
    bc_arr = updes.zerofy_periodic_cond(bc_arr, cloud)
    M = updes.compute_nb_monomials(max_degree, cloud.dim)
    q = updes.assemble_q(rhs_operator, bc_arr, cloud, rbf, M, rhs_args)
    betas = np.array([float(robin[k]) for k in sorted(robin)]) if robin else np.zeros(0)
    return dict(vals=npa(sol.vals), coeffs=npa(sol.coeffs), B=npa(sol.mat), q=npa(q), betas=betas), robin


# ---- cases -----------------------------------------------------------------------------------------------
def laplace_op(x, center, rbf, monomial, fields):
    return updes.nodal_laplacian(x, center, rbf, monomial)


def zero_rhs(x, centers, rbf, fields):
    return 0.0


def case_laplace(nx, ny, keep_blocks=True):
    cloud = updes.SquareCloud(Nx=nx, Ny=ny, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})
    bcs = {"South": lambda c: 0.0, "West": lambda c: 0.0, "North": lambda c: jnp.sin(jnp.pi * c[0]), "East": lambda c: 0.0}
    out, robin = solve(laplace_op, zero_rhs, cloud, bcs, updes.polyharmonic, 1)
    if keep_blocks:
        out.update(blocks(laplace_op, cloud, updes.polyharmonic, 1, None, robin))
    else:
        rows = np.arange(0, cloud.N, 37)
        out["B_rows"], out["B_sample"] = rows, out.pop("B")[rows]
    out.update(cloud_arrays(cloud))
    return out


def case_robin():
    cloud = updes.SquareCloud(Nx=11, Ny=8, facet_types={"South": "n", "West": "r", "North": "d", "East": "r"})
    rbf = partial(updes.gaussian, eps=3.0)

    def op(x, center, rbf, monomial, fields):
        val = updes.nodal_value(x, center, rbf, monomial)
        grad = updes.nodal_gradient(x, center, rbf, monomial)
        lap = updes.nodal_laplacian(x, center, rbf, monomial)
        return 2.5 * val + jnp.dot(jnp.array([1.5, -0.5]), grad) - 0.3 * lap

    def rhs(x, centers, rbf, fields):
        return jnp.cos(3.0 * x[0]) * x[1]

    bcs = {"South": lambda c: 0.25 * c[0], "West": (lambda c: 1.0 + c[1], lambda c: 2.0 + c[1]),
           "North": lambda c: jnp.sin(jnp.pi * c[0]), "East": (lambda c: -0.5, 0.75 * jnp.ones(6))}
    out, robin = solve(op, rhs, cloud, bcs, rbf, 2)
    out.update(blocks(op, cloud, rbf, 2, None, robin))
    out.update(cloud_arrays(cloud))
    return out


def case_periodic():
    DT, VEL, K = 1e-4, (100.0, 0.0), 0.08
    cloud = updes.SquareCloud(Nx=10, Ny=10, facet_types={"South": "p1", "North": "p1", "West": "p2", "East": "p2"})
    rbf = partial(updes.polyharmonic, a=1)

    def op(x, center, rbf, monomial, fields):
        val = updes.nodal_value(x, center, rbf, monomial)
        grad = updes.nodal_gradient(x, center, rbf, monomial)
        lap = updes.nodal_laplacian(x, center, rbf, monomial)
        return (val / DT) + jnp.dot(jnp.array(VEL), grad) - K * lap

    def rhs(x, centers, rbf, fields):
        return updes.value(x, fields[:, 0], centers, rbf) / DT

    xy = npa(cloud.sorted_nodes)
    u0 = np.exp(-((xy[:, 0] - 0.35) ** 2 + (xy[:, 1] - 0.5) ** 2) / (2 * 0.1 ** 2))
    bcs = {k: (lambda c: 0.0) for k in cloud.facet_types}
    out, robin = solve(op, rhs, cloud, bcs, rbf, 0, rhs_args=[jnp.array(u0)])
    out.update(blocks(op, cloud, rbf, 0, None, robin))
    out.update(u0=u0, **cloud_arrays(cloud))
    return out


KERNELS = [("polyharmonic", "a", 2), ("thin_plate", "a", 1), ("gaussian", "eps", 4.0), ("multiquadric", "eps", 2.0),
           ("inverse_multiquadric", "eps", 1.5)]


def case_kernels():
    cloud = updes.SquareCloud(Nx=7, Ny=6, facet_types={"South": "n", "West": "d", "North": "d", "East": "n"})
    xy = npa(cloud.sorted_nodes)
    f0, f1 = 1.0 + xy[:, 0] * xy[:, 1], np.cos(2.0 * xy[:, 0]) - xy[:, 1]

    def op(x, center, rbf, monomial, fields):
        val = updes.nodal_value(x, center, rbf, monomial)
        grad = updes.nodal_gradient(x, center, rbf, monomial)
        lap = updes.nodal_laplacian(x, center, rbf, monomial)
        dg = updes.nodal_div_grad(x, center, rbf, monomial, (fields[0], fields[1]))
        return fields[0] * val + jnp.dot(jnp.array([fields[1], 0.3]), grad) - 0.7 * lap + dg

    out = dict(f0=f0, f1=f1, **cloud_arrays(cloud))
    rng = np.random.default_rng(4)
    pts = np.concatenate([xy[[0, 5, 17, 30, 41]], rng.uniform(0.05, 0.95, size=(6, 2))])     # nodes (r = 0 terms) and free points
    out["eval_pts"] = pts
    for name, pname, pval in KERNELS:
        rbf = partial(getattr(updes, name), **{pname: pval})
        b = blocks(op, cloud, rbf, 4, [jnp.array(f0), jnp.array(f1)], {})
        for k in ("opPhi", "opP", "bdPhi", "bdP", "A"):
            out["%s_%s" % (name, k)] = b[k]
        coeffs = jnp.array(rng.normal(size=cloud.N + 15))
        out["%s_coeffs" % name] = npa(coeffs)
        out["%s_value" % name] = npa(updes.value_vec(jnp.array(pts), coeffs, cloud.sorted_nodes, rbf))
        out["%s_gradient" % name] = npa(updes.gradient_vec(jnp.array(pts), coeffs, cloud.sorted_nodes, rbf))
        out["%s_laplacian" % name] = npa(updes.laplacian_vec(jnp.array(pts), coeffs, cloud.sorted_nodes, rbf))
        coeffs2 = jnp.array(rng.normal(size=cloud.N + 15))
        out["%s_coeffs2" % name] = npa(coeffs2)
        out["%s_divergence" % name] = npa(updes.divergence_vec(jnp.array(pts), jnp.stack([coeffs, coeffs2], axis=-1), cloud.sorted_nodes, rbf))
        # coefficients of a nodal field: inv(A) [field; 0] (assembly.py:404-430), degree 2 keeps cond(A) moderate
        out["%s_field_coeffs" % name] = npa(updes.get_field_coefficients(jnp.array(f1), cloud, rbf, 2))
    out["kernel_names"] = np.array([k[0] for k in KERNELS])
    out["kernel_params"] = np.array([float(k[2]) for k in KERNELS])
    return out


MESH_FACETS = {"vel": {"Wall": "d", "Inflow": "d", "Outflow": "n", "Blowing": "d", "Suction": "d"},
               "phi": {"Wall": "n", "Inflow": "n", "Outflow": "d", "Blowing": "n", "Suction": "n"}}


def case_mesh(tag):
    cloud = updes.GmshCloud(filename=os.path.join(REFERENCE, "updes/tests/data/mesh.msh"), facet_types=MESH_FACETS[tag])
    out = cloud_arrays(cloud)
    if tag == "phi":
        # boundary rows only (assembly.py:141-362): Dirichlet + Neumann rows with the normals GmshCloud computed
        bdPhi, bdP = updes.assemble_bd_Phi_P(cloud, updes.polyharmonic, 3, {})
        rows = np.arange(0, bdPhi.shape[0], 9)
        out.update(bd_rows=rows, bdPhi_sample=npa(bdPhi)[rows], bdP_sample=npa(bdP)[rows])
    return out


GENERATED_MSH = [("a", {"Wall": "d", "Inflow": "n", "Outflow": "r"}, 13, 9), ("b", {"Outflow": "n", "Wall": "r", "Inflow": "d"}, 10, 12)]


def case_generated_msh():
    """The reference's GmshCloud on Gmsh-4.0 meshes written by tests/golden/make_msh.py (own data, regenerated by the
    tests): corner-to-facet assignment by facet_types ORDER (cloud.py:39-40, :680-688) and the computed normals."""
    import tempfile
    sys.path.insert(0, HERE)
    from make_msh import write_channel_msh
    out = {}
    for tag, ft, nx, ny in GENERATED_MSH:
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "channel.msh")
            write_channel_msh(path, nx=nx, ny=ny)
            cloud = updes.GmshCloud(filename=path, facet_types=dict(ft))
        for k, v in cloud_arrays(cloud).items():
            out["%s_%s" % (tag, k)] = v
        out["%s_facets_in" % tag] = np.array([[f, t] for f, t in ft.items()])
        out["%s_grid" % tag] = np.array([nx, ny])
    return out


FUZZ_KERNELS = [("polyharmonic", "a", [1, 2, 3]), ("thin_plate", "a", [1, 2]), ("gaussian", "eps", [0.7, 2.0, 5.0]),
                ("multiquadric", "eps", [0.5, 1.0, 3.0]), ("inverse_multiquadric", "eps", [0.8, 2.5])]


def fuzz_configs(seed=2024, ncases=16):
    """Random small problems: grid size, facet types (d / n / r / periodic pairs) in a random DICT ORDER (the reference's
    corner precedence, periodic-class suffixes and renumbering depend on it), kernel + parameter, polynomial degree."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(ncases):
        nx, ny = int(rng.integers(5, 9)), int(rng.integers(5, 8))
        pattern = k % 4
        t = {f: str(rng.choice(["d", "n", "r"])) for f in ("South", "North", "West", "East")}
        if pattern == 1:
            t["South"] = t["North"] = "p1"
        elif pattern == 2:
            t["West"] = t["East"] = "p1"
        elif pattern == 3:
            a, b = ("p1", "p2") if rng.random() < 0.5 else ("p2", "p1")
            t["South"] = t["North"] = a
            t["West"] = t["East"] = b
        order = [str(v) for v in rng.permutation(["South", "North", "West", "East"])]
        name, pname, vals = FUZZ_KERNELS[int(rng.integers(len(FUZZ_KERNELS)))]
        out.append(dict(Nx=nx, Ny=ny, facets={f: t[f] for f in order}, kernel=name, pname=pname,
                        param=vals[int(rng.integers(len(vals)))], degree=int(rng.integers(0, 5)), seed=int(rng.integers(1 << 30))))
    return out


def case_fuzz():
    """diffMat = [[opPhi opP], [bdPhi bdP]] of every random problem, with a fully general operator
    f0 phi + f1 phi_x + f2 phi_y + f3 phi_xx + f4 phi_yy (five random nodal fields) and random Robin coefficients
    given the way a user gives them: (value, beta array) tuples per Robin facet."""
    def op(x, center, rbf, monomial, fields):
        val = updes.nodal_value(x, center, rbf, monomial)
        grad = updes.nodal_gradient(x, center, rbf, monomial)
        dg = updes.nodal_div_grad(x, center, rbf, monomial, (fields[3], fields[4]))
        return fields[0] * val + jnp.dot(jnp.array([fields[1], fields[2]]), grad) + dg

    out = {}
    cfgs = fuzz_configs()
    for k, cfg in enumerate(cfgs):
        rng = np.random.default_rng(cfg["seed"])
        cloud = updes.SquareCloud(Nx=cfg["Nx"], Ny=cfg["Ny"], facet_types=dict(cfg["facets"]))
        rbf = partial(getattr(updes, cfg["kernel"]), **{cfg["pname"]: cfg["param"]})
        fields = rng.normal(size=(5, cloud.N))
        bcs = {}
        for f, ft in cloud.facet_types.items():
            nf = len(cloud.facet_nodes[f])
            bcs[f] = (jnp.zeros((nf,)), jnp.array(rng.uniform(0.2, 3.0, size=nf))) if ft == "r" else jnp.zeros((nf,))
        robin, _ = updes.duplicate_robin_coeffs(bcs, cloud)
        b = blocks(op, cloud, rbf, cfg["degree"], [jnp.array(f) for f in fields], robin)
        pre = "c%02d_" % k
        out[pre + "diffMat"] = np.concatenate([np.concatenate([b["opPhi"], b["opP"]], axis=1), np.concatenate([b["bdPhi"], b["bdP"]], axis=1)], axis=0)
        out[pre + "fields"] = fields
        out[pre + "betas"] = np.array([float(robin[i]) for i in sorted(robin)]) if robin else np.zeros(0)
        ca = cloud_arrays(cloud)
        for key in ("sorted_nodes", "sorted_outward_normals", "counts", "Np", "facet_names", "facet_sizes", "facet_nodes", "facet_types"):
            out[pre + key] = ca[key]
        out[pre + "config"] = np.array([cfg["Nx"], cfg["Ny"], cfg["degree"]])
        out[pre + "kernel"] = np.array([cfg["kernel"], str(cfg["param"])])
        out[pre + "facets_in"] = np.array([[f, t] for f, t in cfg["facets"].items()])
    out["ncases"] = np.array(len(cfgs))
    return out


MULTI_CLOUD = dict(Nx=9, Ny=8, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})


def case_multi(nb_iters=3):
    """pde_multi_solver (operators.py:696-771) on two genuinely coupled equations,
        lap(u0) - (1 + u1^2) u0 = 0,     lap(u1) + (x + u0) d(u1)/dx = -1."""
    cloud = updes.SquareCloud(**MULTI_CLOUD)
    rbf = partial(updes.polyharmonic, a=1)
    zero, one = (lambda c: 0.0), (lambda c: 1.0)
    bcs = [{"South": zero, "West": zero, "North": one, "East": zero}, {"South": zero, "West": one, "North": zero, "East": zero}]
    op0 = lambda x, c, r, m, f: updes.nodal_laplacian(x, c, r, m) - (1.0 + f[1] ** 2) * updes.nodal_value(x, c, r, m)
    op1 = lambda x, c, r, m, f: updes.nodal_laplacian(x, c, r, m) + (x[0] + f[0]) * updes.nodal_gradient(x, c, r, m)[0]
    rhs0 = lambda x, centers, rbf, fields: 0.0
    rhs1 = lambda x, centers, rbf, fields: -1.0
    z = jnp.zeros((cloud.N,))
    out = dict(cloud_arrays(cloud), nb_iters=np.array(nb_iters))
    for k in range(1, nb_iters + 1):        # the state after every sweep (the reference returns only the last one)
        sols = updes.pde_multi_solver([op0, op1], [rhs0, rhs1], cloud, bcs, rbf, 1, nb_iters=k, diff_args=[[z, z], [z, z]],
                                      rhs_args=[None, None])
        out["vals0_after_%d" % k], out["vals1_after_%d" % k] = npa(sols[0].vals), npa(sols[1].vals)
    out["coeffs0"], out["coeffs1"] = npa(sols[0].coeffs), npa(sols[1].coeffs)
    return out


def case_laplace_demo():
    """demos/Laplace/00_laplace_with_rbf.py run UNMODIFIED, whole script (runpy, in a scratch directory because it creates
    ./data/TempFolder; its plotting calls go to the no-op pyplot stand-in): 30x30 Laplace problem, the solution, the
    Laplacian of the solution at the nodes, and the two error figures the script prints."""
    import runpy
    import tempfile
    demo = os.path.join(REFERENCE, "demos/Laplace/00_laplace_with_rbf.py")
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.mkdir(os.path.join(d, "data"))
        os.chdir(d)
        try:
            ns = runpy.run_path(demo, run_name="__main__")
        finally:
            os.chdir(cwd)
    cloud, sol = ns["cloud"], ns["sol"]
    return dict(cloud_arrays(cloud), vals=npa(sol.vals), coeffs=npa(sol.coeffs), laplacian_at_nodes=npa(ns["lap"]),
                exact=npa(ns["exact_sol"]), mse_total=np.array(float(jnp.mean(ns["error"] ** 2))),
                mse_neumann=np.array(float(jnp.mean(ns["error_neumann"] ** 2))))


def _run_demo_unmodified(relpath):
    """runpy of a reference demo script in a scratch directory (the demos create ./data/<run name>/ and save figures
    through the no-op pyplot stand-in); returns the script's global namespace."""
    import runpy
    import tempfile
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.mkdir(os.path.join(d, "data"))
        os.chdir(d)
        try:
            return runpy.run_path(os.path.join(REFERENCE, relpath), run_name="__main__")
        finally:
            os.chdir(cwd)


def case_darcy_demo():
    """demos/Darcy/00_darcy_flow.py run UNMODIFIED (20x20, all Dirichlet): (1) the permeability field passed through an
    identity-operator solve with polyharmonic a = 2, degree 2; (2) -div(k grad u) = 1 through nodal_div_grad with the
    nodal field k as diff_args, thin_plate a = 3, degree 2."""
    ns = _run_demo_unmodified("demos/Darcy/00_darcy_flow.py")
    cloud, perm, uf = ns["cloud"], ns["perm_field"], ns["ufield"]
    return dict(cloud_arrays(cloud), permeability=npa(ns["permeability"]), perm_vals=npa(perm.vals), perm_coeffs=npa(perm.coeffs),
                u_vals=npa(uf.vals), u_coeffs=npa(uf.coeffs), max_degree=np.array(ns["MAX_DEGREE"]))


def case_integrals():
    """The setting of the reference's own test updes/tests/test_integrals.py (12x12 all-Dirichlet cloud, polyharmonic a = 5,
    degree 3, s = x^2 / (1 + y^2)): get_field_coefficients, the field rebuilt with value_vec, integrate_field (pi/12 to 1e-1)."""
    cloud = updes.SquareCloud(Nx=12, Ny=12, facet_types={"North": "d", "South": "d", "East": "d", "West": "d"}, support_size="max", noise_key=None)
    rbf = partial(updes.polyharmonic, a=5)
    xy = cloud.sorted_nodes
    s = xy[:, 0] ** 2 / (1 + xy[:, 1] ** 2)
    coeffs = updes.get_field_coefficients(s, cloud, rbf, 3)
    rebuilt = updes.value_vec(cloud.sorted_nodes, coeffs, cloud.sorted_nodes, rbf)
    return dict(cloud_arrays(cloud), s=npa(s), coeffs=npa(coeffs), rebuilt=npa(rebuilt),
                integral=np.array(float(updes.integrate_field(coeffs, cloud, rbf, 3))))


def case_config2(nb_steps=3):
    """Config 2 as the reference's demo defines it: constants, cloud (35x35, doubly periodic, key = None), operators and
    initial field are the source text of demos/Advection/01_adv_diff_periodic.py:34-93 executed unchanged; the time loop
    below is the demo's own pde_solver_jit call (:106-113), run for three of its hundred steps."""
    import jax
    src = open(os.path.join(REFERENCE, "demos/Advection/01_adv_diff_periodic.py")).read()
    body = src[src.index("RBF = partial(polyharmonic, a=1)"):src.index("## Begin timestepping for 100 steps")]
    ns = {k: getattr(updes, k) for k in dir(updes) if not k.startswith("_")}
    ns.update(jax=jax, jnp=jnp, partial=partial, key=None)
    exec(compile(body, "01_adv_diff_periodic.py", "exec"), ns)
    cloud, u = ns["cloud"], ns["u0"]
    ulist = [npa(u)]
    for _ in range(nb_steps):
        ufield = updes.pde_solver_jit(diff_operator=ns["my_diff_operator"], rhs_operator=ns["my_rhs_operator"], rhs_args=[u],
                                      cloud=cloud, boundary_conditions=ns["boundary_conditions"], rbf=ns["RBF"],
                                      max_degree=ns["MAX_DEGREE"])
        u = ufield.vals
        ulist.append(npa(u))
    return dict(cloud_arrays(cloud), u=np.stack(ulist), DT=np.array(ns["DT"]), K=np.array(ns["K"]), VEL=npa(ns["VEL"]),
                max_degree=np.array(ns["MAX_DEGREE"]), coeffs_last=npa(ufield.coeffs))


def _advection_demo(relpath, nb_steps, with_sink):
    """Constants, cloud, operators and initial field = the demo's source text executed unchanged (up to its time loop);
    the loop below is the demo's own pde_solver_jit call, run for `nb_steps` of its steps."""
    import jax
    src = open(os.path.join(REFERENCE, relpath)).read()
    body = src[src.index("RBF = partial(polyharmonic, a=1)"):src.index("## Begin timestepping for 100 steps")]
    ns = {k: getattr(updes, k) for k in dir(updes) if not k.startswith("_")}
    ns.update(jax=jax, jnp=jnp, partial=partial, key=None)
    exec(compile(body, os.path.basename(relpath), "exec"), ns)
    cloud, u = ns["cloud"], ns["u0"]
    ulist = [npa(u)]
    for _ in range(nb_steps):
        kw = dict(diff_args=[ns["u_sink"]]) if with_sink else {}
        ufield = updes.pde_solver_jit(diff_operator=ns["my_diff_operator"], rhs_operator=ns["my_rhs_operator"], rhs_args=[u],
                                      cloud=cloud, boundary_conditions=ns["boundary_conditions"], rbf=ns["RBF"],
                                      max_degree=ns["MAX_DEGREE"], **kw)
        u = ufield.vals
        ulist.append(npa(u))
    out = dict(cloud_arrays(cloud), u=np.stack(ulist), DT=np.array(ns["DT"]), K=np.array(ns["K"]), VEL=npa(ns["VEL"]),
               max_degree=np.array(ns["MAX_DEGREE"]), coeffs_last=npa(ufield.coeffs))
    return out, ns, cloud


def case_advection00(nb_steps=2):
    """demos/Advection/00_advection_with_rbf.py (40x20, Dirichlet on three sides, Neumann outflow, degree 1): its u0 is
    0.95 on the N // 40 nearest neighbours of one node, read from cloud.local_supports (:66-69)."""
    out, ns, cloud = _advection_demo("demos/Advection/00_advection_with_rbf.py", nb_steps, False)
    sid = int(ns["source_id"])
    out.update(source_id=np.array(sid), source_neighbors=npa(ns["source_neighbors"]).astype(np.int64),
               source_support=np.array(cloud.local_supports[sid], dtype=np.int64))
    return out


def case_advection02(nb_steps=2):
    """demos/Advection/02_adv_diff_periodic_with_sink.py (35x35 doubly periodic, pure advection K = 0, VEL = 500, degree 0):
    the operator carries a nodal sink field through diff_args (`fields[0] * val`, :58-62)."""
    out, ns, _ = _advection_demo("demos/Advection/02_adv_diff_periodic_with_sink.py", nb_steps, True)
    out.update(u_sink=npa(ns["u_sink"]))
    return out


def case_grayscott001(nb_steps=2):
    """demos/Gray-Scott/001_gray-scott.py (despite its name an advection-diffusion loop): 40x20 cloud with periodic ids
    "p0" / "p1", degree 1 (three monomials beside periodic rows), u0 through cloud.local_supports of the middle node."""
    out, ns, cloud = _advection_demo("demos/Gray-Scott/001_gray-scott.py", nb_steps, False)
    sid = int(ns["source_id"])
    out.update(source_id=np.array(sid), source_neighbors=npa(ns["source_neighbors"]).astype(np.int64),
               source_support=np.array(cloud.local_supports[sid], dtype=np.int64))
    return out


def case_wave00(nb_steps=2):
    """demos/Wave/00_wave.py: all four facets Neumann (no Dirichlet row at all), polyharmonic a = 3, degree 2, operator
    val / DT^2 + C lap, right-hand side from TWO nodal fields (2 u_prev - u_prev_prev) / DT^2; definitions executed from the
    demo's source, the loop below is its own (:88-100), two of its 500 steps."""
    import jax
    src = open(os.path.join(REFERENCE, "demos/Wave/00_wave.py")).read()
    body = src[src.index("RBF = partial(polyharmonic, a=3)"):src.index("## Begin timestepping for 100 steps")]
    ns = {k: getattr(updes, k) for k in dir(updes) if not k.startswith("_")}
    ns.update(jax=jax, jnp=jnp, partial=partial, key=None)
    exec(compile(body, "00_wave.py", "exec"), ns)
    cloud, u0, DT = ns["cloud"], ns["u0"], ns["DT"]
    ulist = [u0, 1 * DT + u0]
    for _ in range(nb_steps):
        ufield = updes.pde_solver_jit(diff_operator=ns["my_diff_operator"], rhs_operator=ns["my_rhs_operator"],
                                      rhs_args=[ulist[-1], ulist[-2]], cloud=cloud, boundary_conditions=ns["boundary_conditions"],
                                      rbf=ns["RBF"], max_degree=ns["MAX_DEGREE"])
        ulist.append(ufield.vals)
    return dict(cloud_arrays(cloud), u=np.stack([npa(x) for x in ulist]), DT=np.array(DT), C=np.array(ns["C"]),
                max_degree=np.array(ns["MAX_DEGREE"]), coeffs_last=npa(ufield.coeffs))


def case_cartesian_helpers():
    """cartesian_gradient / cartesian_gradient_vec / enforce_cartesian_gradient_neumann / apply_neumann_conditions
    (operators.py:211-291, :483-509; called by demos/NavierStokes/11_...:187 and 16_...:181) on two square clouds and on a
    generated channel mesh, for f = sin(3x) cos(2y) + x^2 and a fixed random gradient table."""
    import tempfile
    sys.path.insert(0, HERE)
    from make_msh import write_channel_msh
    out = {}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "channel.msh")
        write_channel_msh(path, nx=9, ny=6)
        clouds = [("sq_a", updes.SquareCloud(Nx=7, Ny=6, facet_types={"South": "n", "West": "d", "North": "d", "East": "n"}, support_size="max", noise_key=None)),
                  ("sq_b", updes.SquareCloud(Nx=8, Ny=5, facet_types={"South": "d", "West": "n", "North": "r", "East": "d"}, support_size="max", noise_key=None)),
                  ("msh", updes.GmshCloud(filename=path, facet_types={"Wall": "n", "Inflow": "d", "Outflow": "n"}))]
        for tag, c in clouds:
            xy = npa(c.sorted_nodes)
            f = np.sin(3 * xy[:, 0]) * np.cos(2 * xy[:, 1]) + xy[:, 0] ** 2
            g0 = np.random.default_rng(0).normal(size=(c.N, 2))
            fj = jnp.array(f)
            out[tag + "_nodes"] = xy
            out[tag + "_f"], out[tag + "_g0"] = f, g0
            out[tag + "_grad"] = npa(updes.cartesian_gradient_vec(range(c.N), fj, c))
            out[tag + "_grad3_clipped"] = npa(updes.cartesian_gradient(3, fj, c, clip_val=0.05))
            out[tag + "_enforced"] = npa(updes.enforce_cartesian_gradient_neumann(fj, jnp.array(g0), {}, c))
            out[tag + "_applied"] = npa(updes.apply_neumann_conditions(fj, {}, c))
    return out


def case_config3(nb_iter=2):
    """Config 3 as the reference's demo runs it: the source text of simulate_forward_navier_stokes and its six operators
    is read from demos/NavierStokes/30_channel_flow_blowing_suction.py:61-250 and executed unchanged (the rest of that
    script builds its clouds with the gmsh package and plots); both clouds come from the reference's mesh.msh."""
    import jax
    from jax.tree_util import Partial
    src = open(os.path.join(REFERENCE, "demos/NavierStokes/30_channel_flow_blowing_suction.py")).read()
    body = src[src.index("def diff_operator_u("):src.index("def diff_operator_id(")]
    ns = {k: getattr(updes, k) for k in dir(updes) if not k.startswith("_")}
    ns.update(jax=jax, jnp=jnp, Partial=Partial, partial=partial, RBF=updes.polyharmonic, MAX_DEGREE=1, Re=100, Pa=0., NB_ITER=5)
    exec(compile(body, "30_channel_flow_blowing_suction.py", "exec"), ns)
    mesh = os.path.join(REFERENCE, "updes/tests/data/mesh.msh")
    cloud_vel = updes.GmshCloud(filename=mesh, facet_types=MESH_FACETS["vel"])
    cloud_phi = updes.GmshCloud(filename=mesh, facet_types=MESH_FACETS["phi"])
    u_list, v_list, vel_list, p_list = ns["simulate_forward_navier_stokes"](cloud_vel, cloud_phi, NB_ITER=nb_iter)
    return dict(u=np.stack([npa(x) for x in u_list]), v=np.stack([npa(x) for x in v_list]), p=np.stack([npa(x) for x in p_list]),
                vel=np.stack([npa(x) for x in vel_list]), nb_iter=np.array(nb_iter), Re=np.array(100.0))


CASES = {"ref_laplace_12x9": lambda: case_laplace(12, 9), "ref_robin_11x8": case_robin, "ref_periodic_10x10": case_periodic,
         "ref_kernels_7x6": case_kernels, "ref_config1_30x20": lambda: case_laplace(30, 20, keep_blocks=False),
         "ref_mesh_msh_vel": lambda: case_mesh("vel"), "ref_mesh_msh_phi": lambda: case_mesh("phi"),
         "ref_integrals_12x12": case_integrals, "ref_laplace_demo_30x30": case_laplace_demo, "ref_darcy_demo_20x20": case_darcy_demo, "ref_config2_advdiff_3steps": case_config2, "ref_config3_ns_2iter": case_config3, "ref_multi_solver_9x8": case_multi, "ref_fuzz_16": case_fuzz, "ref_generated_msh": case_generated_msh,
         "ref_advection00_2steps": case_advection00, "ref_advection02_sink_2steps": case_advection02,
         "ref_grayscott001_2steps": case_grayscott001, "ref_wave00_2steps": case_wave00,
         "ref_cartesian_helpers": case_cartesian_helpers}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default=HERE)
    args = ap.parse_args()
    for name, fn in CASES.items():
        if args.only and name not in args.only.split(","):
            continue
        t0 = time.time()
        out = fn()
        np.savez_compressed(os.path.join(args.out, name + ".npz"), **out)
        print("%-22s %6.1f s  %d arrays  %.0f KB" % (name, time.time() - t0, len(out),
                                                     os.path.getsize(os.path.join(args.out, name + ".npz")) / 1e3), flush=True)


if __name__ == "__main__":
    main()
