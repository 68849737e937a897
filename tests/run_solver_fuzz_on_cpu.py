#!/usr/bin/env python
"""Randomised sweep of the host layer on the CPU emulation of the C-ABI (tests/cpu_abi_emulation.py): random facet-type
mixes (d / n / r / one or two periodic pairs) in random dict order, every kernel, degrees 0-3, a linear operator with
constant, coordinate-dependent and field-dependent coefficients (diff_args), a right-hand side that evaluates a nodal
field (rhs_args), boundary data as callables / arrays / (value, beta) tuples.  For each problem updes_b200.pde_solver's
coefficients must solve the system the ORACLE assembles from independently written inputs (hand-lowered coefficients,
per-facet arrays, Robin betas by the reference's offset rule): backward error <= 1e-12 and vals == [Phi P] c.
Checks lowering, BC preparation, row descriptors and rhs assembly -- not the kernels.

    python tests/run_solver_fuzz_on_cpu.py [seed] [cases]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import warnings

import numpy as np
from functools import partial
import cpu_abi_emulation as emu
emu.install()
import updes_b200 as u
from oracle import oracle as O
O.build()

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 30
worst = 0
bad = 0
for case in range(ncase):
    nx, ny = int(rng.integers(6, 12)), int(rng.integers(6, 12))
    names = ["South", "North", "West", "East"]
    mode = rng.integers(0, 4)
    types = {}
    if mode == 0:   # no periodic
        for n in names: types[n] = str(rng.choice(["d", "n", "r"]))
        if all(t != "d" for t in types.values()): types["North"] = "d"
    elif mode == 1:
        types = {"South": "p1", "North": "p1", "West": str(rng.choice(["d", "n", "r"])), "East": "d"}
    elif mode == 2:
        types = {"West": "p1", "East": "p1", "South": "d", "North": str(rng.choice(["d", "n", "r"]))}
    else:
        types = {"South": "p1", "North": "p1", "West": "p2", "East": "p2"}
    order = list(types); rng.shuffle(order)
    facets = {k: types[k] for k in order}
    kname = str(rng.choice(["polyharmonic", "thin_plate", "gaussian", "multiquadric", "inverse_multiquadric"]))
    if kname in ("polyharmonic", "thin_plate"):
        param = int(rng.integers(1, 3)); rbf = partial(getattr(u, kname), a=param)
    else:
        param = float(rng.uniform(1.0, 4.0)); rbf = partial(getattr(u, kname), eps=param)
    deg = int(rng.integers(0 if mode == 3 else 1, 4))
    M = u.compute_nb_monomials(deg, 2)
    cloud = u.SquareCloud(Nx=nx, Ny=ny, facet_types=facets)
    N, Ni = cloud.N, cloud.Ni
    f0, f1 = rng.normal(size=N), rng.normal(size=N)
    a = rng.normal(size=5)
    def op(x, center, rbf, monomial, fields):
        val = u.nodal_value(x, center, rbf, monomial)
        g = u.nodal_gradient(x, center, rbf, monomial)
        lap = u.nodal_laplacian(x, center, rbf, monomial)
        return (a[0] + fields[0]) * val + a[1] * x[0] * g[0] + (a[2] + fields[1]) * g[1] + a[3] * lap
    coef = np.stack([a[0] + f0[:Ni], a[1] * cloud.sorted_nodes[:Ni, 0], a[2] + f1[:Ni], np.full(Ni, a[3]), np.full(Ni, a[3])], axis=1)
    g0 = rng.normal(size=N)
    def rhs(x, centers, rbf, fields):
        return np.sin(x[0]) * x[1] + 0.5 * u.value(x, fields[:, 0], centers, rbf)
    A0 = O.assemble_A(cloud, kname, float(param), M)
    cg = np.linalg.solve(A0, np.concatenate([g0, np.zeros(M)]))
    q_int = np.sin(cloud.sorted_nodes[:Ni, 0]) * cloud.sorted_nodes[:Ni, 1] + 0.5 * O.eval_field(cloud.sorted_nodes[:Ni], cloud.sorted_nodes, cg, kname, float(param), "value")
    bcs, arrs = {}, {}
    betas_by_node = {}
    for f, t in cloud.facet_types.items():
        ids = np.asarray(cloud.facet_nodes[f]); pts = cloud.sorted_nodes[ids]
        form = rng.integers(0, 2)
        vals = np.cos(pts[:, 0] + 2 * pts[:, 1])
        if t == "r":
            beta = 1.0 + pts[:, 0]
            bcs[f] = ((lambda p: np.cos(p[0] + 2 * p[1])) if form else vals, (lambda p: 1.0 + p[0]) if rng.integers(0, 2) else beta)
            for i in ids: betas_by_node[int(i)] = beta[min(int(i - ids[0]), len(ids) - 1)]
            arrs[f] = vals
        elif t[0] == "p":
            bcs[f] = (lambda p: 3.0)            # ignored: periodic rhs := 0
            arrs[f] = np.zeros(len(ids))
        else:
            bcs[f] = (lambda p: np.cos(p[0] + 2 * p[1])) if form else vals
            arrs[f] = vals
    betas = np.array([betas_by_node[k] for k in sorted(betas_by_node)]) if betas_by_node else None
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sol = u.pde_solver(op, rhs, cloud, bcs, rbf, deg, diff_args=[f0, f1], rhs_args=[g0])
    except Exception as e:
        print("case", case, facets, kname, param, deg, "RAISED", repr(e)); raise
    K = O.assemble_K(cloud, kname, float(param), M, coef, betas)
    q = O.assemble_q(cloud, q_int, arrs)
    b = np.concatenate([q, np.zeros(M)])
    r = K @ sol.coeffs - b
    berr = np.max(np.abs(r)) / (np.abs(K).sum(1).max() * np.max(np.abs(sol.coeffs)) + np.max(np.abs(b)))
    A = O.assemble_A(cloud, kname, float(param), M)
    dv = np.max(np.abs(A[:N] @ sol.coeffs - sol.vals)) / max(np.max(np.abs(sol.vals)), 1e-300)
    worst = max(worst, berr)
    flag = "" if berr < 1e-12 and dv < 1e-8 else "   <<<<<<<< CHECK"
    bad = bad + (1 if flag else 0)
    print("case %2d %dx%d %s %s(%s) deg %d: backward error %.1e, vals-vs-[Phi P]c %.1e%s" % (case, nx, ny, dict(facets), kname, param, deg, berr, dv, flag))
print("worst backward error %.2e over %d problems, %d outside the bounds" % (worst, ncase, bad))
sys.exit(1 if bad else 0)
