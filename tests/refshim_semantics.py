"""Run in a SUBPROCESS by tests/test_reference_golden.py: the JAX semantics the reference's hot path relies on, asserted on
the stand-in of oracle/refshim (documented JAX behaviour in the comments; none of it is specific to XLA arithmetic)."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))

import jax  # noqa: E402
import jax.numpy as jnp  # noqa: E402
from jax.tree_util import Partial  # noqa: E402
import lineax as lx  # noqa: E402
import torch  # noqa: E402

x64 = jnp.zeros((3,))
assert x64.dtype == torch.float64 and jnp.arange(4).dtype == torch.int64 and jnp.zeros((2, 2), dtype=int).dtype == torch.int64   # jax_enable_x64

# functional updates: x.at[i].set(v) returns a new array and leaves x alone
a = jnp.zeros((3, 3))
b = a.at[1, jnp.array([0, 2])].set(jnp.array([5.0, 7.0]))
assert float(a.sum()) == 0.0 and b[1].tolist() == [5.0, 0.0, 7.0] and a.at[:].add(1.0).sum() == 9.0

# grad / jacfwd differentiate w.r.t. the first argument; the distance has a NaN gradient at coincident points
# (d/dx sqrt(x.x) = x / sqrt(x.x) = 0/0), which the reference then removes with nan_to_num (operators.py:58)
dist = lambda p, c: jnp.sqrt((p - c).T @ (p - c))
p = jnp.array([0.3, 0.7])
g = jax.grad(dist)(p, jnp.array([0.0, 0.3]))
assert np.allclose(g.numpy(), [0.6, 0.8])
assert bool(torch.isnan(jax.grad(dist)(p, p)).all())
cube = lambda p, c: dist(p, c) ** 3
assert bool(torch.isnan(jax.grad(cube)(p, p)).all())                 # 3 r^2 * (0/0): NaN, not the true limit 0
H = jax.jacfwd(jax.grad(cube))(p, jnp.array([0.0, 0.3]))
r, d = 0.5, np.array([0.3, 0.4])
assert np.allclose(H.numpy(), 3 * r * np.eye(2) + 3 * np.outer(d, d) / r, rtol=1e-13)
assert bool(torch.isnan(jax.jacfwd(jax.grad(cube))(p, p)).any())
assert jnp.nan_to_num(jnp.array([math.nan, math.inf, -math.inf, 2.0]), posinf=0.0, neginf=0.0).tolist() == [0.0, 0.0, 0.0, 2.0]
assert float(jnp.nan_to_num(jnp.log(jnp.array(0.0)) * jnp.array(0.0) ** 2, neginf=0.0, posinf=0.0)) == 0.0     # thin plate at r = 0

# vmap: in_axes None broadcasts an argument (also non-arrays), Python-number outputs become arrays
f = lambda p, c, k, fn: fn(p, c) * k if fn is not None else 1.0
out = jax.vmap(f, in_axes=(None, 0, None, None), out_axes=0)(p, jnp.array([[0.0, 0.3], [0.3, 0.7]]), 2.0, dist)
assert np.allclose(out.numpy(), [1.0, 0.0])
assert jax.vmap(lambda q: 1.0)(jnp.zeros((4, 2))).tolist() == [1.0] * 4
assert jax.vmap(jax.grad(lambda q: 1.0))(jnp.zeros((4, 2))).tolist() == [[0.0, 0.0]] * 4          # gradient of a constant monomial
assert (p != None) is True                                                                           # noqa: E711  (operators.py:28)

# control flow, partial application, tree_map, jit as identity
assert jax.lax.fori_loop(2, 5, lambda i, v: v + i, 0) == 9
assert Partial(lambda a_, b_: a_ - b_, b_=2)(5) == 3 and Partial(jax.jit, static_argnums=1)(abs)(-3) == 3
assert jax.tree_util.tree_map(lambda i: i + 1, {"a": [1, 2], "b": (3,)}) == {"a": [2, 3], "b": (4,)}

# linear algebra: inv through LAPACK, lineax QR solve == least squares of a square full-rank system
rng = np.random.default_rng(0)
M = jnp.array(rng.normal(size=(6, 6)))
v = jnp.array(rng.normal(size=6))
assert np.allclose((jnp.linalg.inv(M) @ M).numpy(), np.eye(6), atol=1e-12)
sol = lx.linear_solve(lx.MatrixLinearOperator(M), v, solver=lx.QR()).value
assert np.allclose(sol.numpy(), np.linalg.solve(M.numpy(), v.numpy()), rtol=1e-10)
assert np.allclose(jnp.dot(jnp.ones((3, 2)), jnp.array([1.0, 2.0])).numpy(), [3.0] * 3) and jnp.linspace(0, 1.0, 5).tolist() == np.linspace(0, 1, 5).tolist()
print("SEMANTICS OK")
