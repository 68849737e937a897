#!/usr/bin/env python
"""Run the reference's demo SCRIPTS, unmodified, against the PRODUCT: `updes` is aliased to `updes_b200` (so the demo's
`from updes import *` takes the product's surface), jax / jax.numpy / matplotlib / seaborn -- absent from the image -- come
from the stand-ins of oracle/refshim (torch-backed arrays go in, the product converts them with numpy.asarray), and the
C-ABI is the CPU emulation of tests/cpu_abi_emulation.py (the build container has no GPU).  What the script computed is
compared with what the REFERENCE computed for the same script (tests/golden/ref_*.npz).  Needs /root/reference.

    python tests/run_reference_demo_on_product.py                 # Laplace/00 and Darcy/00 (seconds)
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


def main():
    assert os.path.isdir(REFERENCE), "needs /root/reference (build container)"
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
