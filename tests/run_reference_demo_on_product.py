#!/usr/bin/env python
"""Run the reference's demo SCRIPTS, unmodified, against the PRODUCT: `updes` is aliased to `updes_b200` (so the demo's
`from updes import *` takes the product's surface), jax / jax.numpy / matplotlib / seaborn -- absent from the image -- come
from the stand-ins of oracle/refshim (torch-backed arrays go in, the product converts them with numpy.asarray), and the
C-ABI is the CPU emulation of tests/cpu_abi_emulation.py (the build container has no GPU).  What the script computed is
compared with what the REFERENCE computed for the same script (tests/golden/ref_*.npz).  The reference's OWN three tests
(updes/tests/test_*.py) are run the same way first.  Needs /root/reference.

    python tests/run_reference_demo_on_product.py                 # Laplace/00, Darcy/00, the NS/30 projection loop (seconds)
    python tests/run_reference_demo_on_product.py --all           # + the 100-step Advection / Gray-Scott loops (minutes on the emulation)
"""
import os
import runpy
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))       # before updes_b200 is imported: plt / sns are optional exports

import cpu_abi_emulation as emu  # noqa: E402

emu.install()
import updes_b200  # noqa: E402
import reference_cases as rc  # noqa: E402

sys.modules["updes"] = updes_b200


def run(relpath):
    cwd = os.getcwd()
    updes_b200.clear_cache()
    with tempfile.TemporaryDirectory() as d:
        os.mkdir(os.path.join(d, "data"))
        os.chdir(d)
        try:
            return runpy.run_path(os.path.join(REFERENCE, "demos", relpath), run_name="__main__")
        finally:
            os.chdir(cwd)


def npa(t):
    return t.detach().numpy() if hasattr(t, "detach") else np.asarray(t)


def rel(a, b):
    return float(np.max(np.abs(npa(a) - b)) / np.max(np.abs(b)))


def reference_own_tests():
    """updes/tests/test_{interpolation,integrals,operators}.py of the reference, unmodified, on the product: the files are
    loaded under another module name (inside the reference tree pytest would import the `updes` package they sit in) and
    every test_* function is called; they read "updes/tests/data/mesh.msh" relative to the reference's root."""
    import importlib.util
    cwd = os.getcwd()
    os.chdir(REFERENCE)
    ran = 0
    try:
        for name in ("test_interpolation", "test_integrals", "test_operators"):
            updes_b200.clear_cache()
            spec = importlib.util.spec_from_file_location("reference_" + name, os.path.join(REFERENCE, "updes", "tests", name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            for fn in sorted(n for n in dir(mod) if n.startswith("test_") and callable(getattr(mod, n))):
                getattr(mod, fn)()
                ran += 1
                print("passed on the product: updes/tests/%s.py::%s" % (name, fn), flush=True)
    finally:
        os.chdir(cwd)
    assert ran == 3
    print("the reference's own %d tests pass on the product" % ran)


def main():
    assert os.path.isdir(REFERENCE), "needs /root/reference (build container)"
    reference_own_tests()
    # the README example (README.md:37-68, BASELINE.json configs[0]): the python block as printed there, executed literally
    readme = open(os.path.join(REFERENCE, "README.md")).read()
    block = readme[readme.index("```python") + len("```python"):]
    block = block[:block.index("```")]
    assert "import updes" in block and "pde_solver_jit" in block
    ns = {}
    updes_b200.clear_cache()
    exec(compile(block, "README.md", "exec"), ns)
    g = rc.load("ref_config1_30x20")
    d = rel(ns["sol"].vals, g["vals"])
    print("README.md example (config 1, 30x20), executed literally: solution vs the reference's pde_solver_jit %.2e" % d)
    assert d <= 1e-8

    ns = run("Laplace/00_laplace_with_rbf.py")
    g = rc.load("ref_laplace_demo_30x30")
    d, mse = rel(ns["sol"].vals, g["vals"]), float(np.mean(npa(ns["error"]) ** 2))
    print("Laplace/00_laplace_with_rbf.py: solution vs the reference's %.2e; MSE it prints %.6e (reference: %.6e)" % (d, mse, float(g["mse_total"])))
    assert d <= 1e-8 and abs(mse - float(g["mse_total"])) <= 1e-5 * float(g["mse_total"])
    assert rel(ns["lap"], g["laplacian_at_nodes"]) <= 1e-6

    ns = run("Darcy/00_darcy_flow.py")
    g = rc.load("ref_darcy_demo_20x20")
    d1, d2 = rel(ns["perm_field"].vals, g["perm_vals"]), rel(ns["ufield"].vals, g["u_vals"])
    print("Darcy/00_darcy_flow.py: permeability solve vs the reference's %.2e, Darcy solution %.2e" % (d1, d2))
    assert d1 <= 2e-7 and d2 <= 2e-7            # both pipelines carry ~1e-8 at cond(K) = 3e11

    # config 3: the projection loop of demos/NavierStokes/30_channel_flow_blowing_suction.py -- the script itself builds its
    # clouds with the gmsh package, so (as the golden generator does with the reference) the source text of its operators
    # and of simulate_forward_navier_stokes is executed unchanged, here against the product, on the reference's mesh.msh
    import jax
    import jax.numpy as jnp
    from functools import partial
    src = open(os.path.join(REFERENCE, "demos/NavierStokes/30_channel_flow_blowing_suction.py")).read()
    body = src[src.index("def diff_operator_u("):src.index("def diff_operator_id(")]
    ns = {k: getattr(updes_b200, k) for k in dir(updes_b200) if not k.startswith("_")}
    ns.update(jax=jax, jnp=jnp, Partial=partial, partial=partial, RBF=updes_b200.polyharmonic, MAX_DEGREE=1, Re=100, Pa=0., NB_ITER=5)
    exec(compile(body, "30_channel_flow_blowing_suction.py", "exec"), ns)
    mesh = os.path.join(REFERENCE, "updes/tests/data/mesh.msh")
    vel = {"Wall": "d", "Inflow": "d", "Outflow": "n", "Blowing": "d", "Suction": "d"}
    phi = {"Wall": "n", "Inflow": "n", "Outflow": "d", "Blowing": "n", "Suction": "n"}
    updes_b200.clear_cache()
    u_list, v_list, _, p_list = ns["simulate_forward_navier_stokes"](updes_b200.GmshCloud(filename=mesh, facet_types=vel),
                                                                     updes_b200.GmshCloud(filename=mesh, facet_types=phi), NB_ITER=2)
    g = rc.load("ref_config3_ns_2iter")
    d = max(rel(lst[k], g[name][k]) for name, lst in (("u", u_list), ("v", v_list), ("p", p_list)) for k in (1, 2))
    print("NavierStokes/30 simulate_forward_navier_stokes (source unchanged, two iterations): u, v, p vs the reference's %.2e" % d)
    assert d <= 2e-5                            # both pipelines feed inv(A)-level errors (cond 1e9) into the next solve

    if "--all" in sys.argv:
        for relpath, golden in (("Advection/00_advection_with_rbf.py", "ref_advection00_2steps"),
                                ("Advection/01_adv_diff_periodic.py", "ref_config2_advdiff_3steps"),
                                ("Advection/02_adv_diff_periodic_with_sink.py", "ref_advection02_sink_2steps"),
                                ("Gray-Scott/001_gray-scott.py", "ref_grayscott001_2steps")):
            ns = run(relpath)
            g = rc.load(golden)
            steps = g["u"].shape[0]
            assert len(ns["ulist"]) == ns["NB_TIMESTEPS"] + 1
            d = max(rel(ns["ulist"][s], g["u"][s]) for s in range(1, steps))
            print("%s: all %d steps run; first %d fields vs the reference's trajectory %.2e" % (relpath, ns["NB_TIMESTEPS"], steps - 1, d))
            assert d <= 1e-7
    print("reference demos ran unmodified on the product")


if __name__ == "__main__":
    main()
