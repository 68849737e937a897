"""Run in a SUBPROCESS by tests/test_reference_golden.py: operators written the way the reference's demos write them --
`jnp.array([...])`, `jnp.dot(...)` on the nodal terms -- lowered by the product.  `jnp` here is the torch-backed stand-in
of oracle/refshim (real JAX is not installed); what matters is that the operator body hands the nodal terms to a foreign
array library, which symbolic lowering cannot follow and the numeric-probe fallback can."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))

import jax.numpy as jnp  # noqa: E402  (stand-in)
import updes_b200 as u  # noqa: E402
from updes_b200 import nodal_gradient, nodal_laplacian, nodal_value, nodal_div_grad  # noqa: E402,F401

cloud = u.SquareCloud(Nx=9, Ny=7, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})
Ni = cloud.Ni
uu, vv = np.linspace(0, 1, cloud.N), np.linspace(2, 3, cloud.N)
Re = 100

ns = {"jnp": jnp, "nodal_gradient": nodal_gradient, "nodal_laplacian": nodal_laplacian, "nodal_value": nodal_value, "Re": Re}
demo = "/root/reference/demos/NavierStokes/30_channel_flow_blowing_suction.py"
if os.path.exists(demo):
    # the demo's own operator definitions, source text executed unchanged against the product's term set
    src = open(demo).read()
    exec(compile(src[src.index("def diff_operator_u("):src.index("# @Partial(jax.jit, static_argnums=[2])\ndef rhs_operator_u")], demo, "exec"), ns)
    diff_operator_u = ns["diff_operator_u"]
    print("operator source: reference demo")
else:
    def diff_operator_u(x, center=None, rbf=None, monomial=None, fields=None):
        U_prev = jnp.array([fields[0], fields[1]])
        u_grad = nodal_gradient(x, center, rbf, monomial)
        u_lap = nodal_laplacian(x, center, rbf, monomial)
        return jnp.dot(U_prev, u_grad) - u_lap / Re
    print("operator source: inline copy of the demo's form")

cphi, cpol = u.lower_diff_operator(diff_operator_u, cloud, u.polyharmonic, [uu, vv])
want = np.stack([np.zeros(Ni), uu[:Ni], vv[:Ni], np.full(Ni, -1 / Re), np.full(Ni, -1 / Re)], axis=1)
assert np.allclose(cphi, want, rtol=1e-15, atol=0) and np.array_equal(cphi, cpol), np.abs(cphi - want).max()

# adv-diff of demos/Advection with jnp
DT, VEL, K = 1e-4, jnp.array([100.0, 0.0]), 0.08


def advdiff(x, center, rbf, monomial, fields):
    val = nodal_value(x, center, rbf, monomial)
    grad = nodal_gradient(x, center, rbf, monomial)
    lap = nodal_laplacian(x, center, rbf, monomial)
    return (val / DT) + jnp.dot(VEL, grad) - K * lap


c, _ = u.lower_diff_operator(advdiff, cloud, u.polyharmonic)
assert np.allclose(c, np.tile([1 / DT, 100.0, 0.0, -K, -K], (Ni, 1)), rtol=1e-15, atol=0)

# non-linear / affine operators written with the foreign library still raise the explicit error
for bad in (lambda x, c, r, m, f: jnp.sin(jnp.array(nodal_value(x, c, r, m))),
            lambda x, c, r, m, f: jnp.dot(jnp.array([1.0, 1.0]), nodal_gradient(x, c, r, m)) + 1.0,
            lambda x, c, r, m, f: jnp.dot(nodal_gradient(x, c, r, m), nodal_gradient(x, c, r, m)),
            lambda x, c, r, m, f: jnp.exp(jnp.array(nodal_laplacian(x, c, r, m))) - 1.0):
    try:
        u.lower_diff_operator(bad, cloud, u.polyharmonic)
    except u.OperatorLoweringError as e:
        print("rejected:", str(e)[:90])
    else:
        raise AssertionError("a non-linear operator was accepted")
print("OK")

# boundary functions and rhs operators written with the foreign library (README.md:52-58 uses jnp.sin(jnp.pi * coord[0]))
bcs = {"South": lambda c: 0.0, "West": lambda c: 0.0, "North": lambda c: jnp.sin(jnp.pi * c[0]), "East": lambda c: 0.0}
arr = u.boundary_conditions_func_to_arr(bcs, cloud)
north = np.asarray(cloud.facet_nodes["North"])
assert np.allclose(np.asarray(arr["North"], dtype=np.float64), np.sin(np.pi * cloud.sorted_nodes[north, 0]), rtol=1e-15, atol=1e-16)
q = u.assemble_q(lambda x, centers, rbf, fields: jnp.cos(3.0 * x[0]) * x[1], arr, cloud, u.polyharmonic, 3, None)
xy = cloud.sorted_nodes
assert np.allclose(q[:Ni], np.cos(3.0 * xy[:Ni, 0]) * xy[:Ni, 1], rtol=1e-15, atol=1e-16) and np.allclose(q[north], arr["North"])
print("OK bc + rhs")
