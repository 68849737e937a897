"""Shared test helpers: reference-style problem definitions used by CPU and GPU tests."""
import numpy as np

import updes_b200 as u

CONFIG1_FACETS = {"South": "n", "West": "d", "North": "d", "East": "d"}          # README example
CONFIG2_FACETS = {"South": "p1", "North": "p1", "West": "p2", "East": "p2"}      # demos/Advection/01


def rel_err_rowscaled(got, want):
    """|got - want| / max(|want|, rowscale): the 1e-12 'relative per entry' test with the absolute
    floor SURVEY.md section 7 (hard part 4) asks for -- entries that pass through zero have no
    meaningful relative error, so the floor is the largest magnitude in the row."""
    got = np.asarray(got); want = np.asarray(want)
    scale = np.maximum(np.abs(want), np.max(np.abs(want), axis=-1, keepdims=True))
    scale = np.where(scale == 0, 1.0, scale)
    return np.max(np.abs(got - want) / scale)


def laplace_op(lib):
    def op(x, center, rbf, monomial, fields):
        return lib.nodal_laplacian(x, center, rbf, monomial)
    return op


def advdiff_op(lib, dt=1e-4, vel=(100.0, 0.0), k=0.08, dot=np.dot):
    def op(x, center, rbf, monomial, fields):
        val = lib.nodal_value(x, center, rbf, monomial)
        grad = lib.nodal_gradient(x, center, rbf, monomial)
        lap = lib.nodal_laplacian(x, center, rbf, monomial)
        return (val / dt) + dot(np.asarray(vel), grad) - k * lap
    return op


def cloud_from_golden(name):
    """Cloud rebuilt from a tests/golden/*.npz written by tests/golden/make_golden.py."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
    names = [str(v) for v in g["facet_names"]]
    sizes = [int(v) for v in g["facet_sizes"]]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    facet_nodes = {nm: g["facet_nodes"][offs[k]:offs[k + 1]].tolist() for k, nm in enumerate(names)}
    facet_types = {nm: str(t) for nm, t in zip(names, g["facet_types"])}
    return u.Cloud.from_arrays(g["sorted_nodes"], g["counts"], g["Np"], facet_types, facet_nodes, g["sorted_outward_normals"],
                               old_of_new=g["old_of_new"] if "old_of_new" in g else None), g


def true_rel_err(got, want, floor=1e-6):
    """TRUE per-entry relative error |got - want| / |want| over the entries that are not near-cancellations
    (|want| > floor * largest magnitude of the row) -- reported next to the row-scaled figure."""
    got = np.asarray(got); want = np.asarray(want)
    big = np.abs(want) > floor * np.max(np.abs(want), axis=-1, keepdims=True)
    return float(np.max(np.abs(got - want)[big] / np.abs(want)[big])) if np.any(big) else 0.0


def exact_solution(K, rhs, A_rows):
    """The discrete collocation solution to (almost) working-precision-independent accuracy: K c = rhs solved by
    LAPACK LU + iterative refinement with 80-bit (numpy longdouble) residuals, then vals = [Phi P] c in longdouble.
    Converges as long as cond(K) * 2^-53 < 1.  Lets a test say how far EACH formulation (the reference's
    inv + GEMM + QR, and the product's LU) is from the solution both approximate."""
    import scipy.linalg as sla
    lu = sla.lu_factor(K)
    Kl, rl = K.astype(np.longdouble), rhs.astype(np.longdouble)
    c = sla.lu_solve(lu, rhs).astype(np.longdouble)
    for _ in range(8):
        r = rl - Kl @ c
        d = sla.lu_solve(lu, np.asarray(r, dtype=np.float64))
        c = c + d.astype(np.longdouble)
        if np.max(np.abs(d)) <= 1e-17 * np.max(np.abs(c)):
            break
    vals = A_rows.astype(np.longdouble) @ c
    return np.asarray(vals, dtype=np.float64), np.asarray(c, dtype=np.float64)


def backward_error(K, c, rhs):
    """normwise backward error ||K c - rhs||_inf / (||K||_inf ||c||_inf + ||rhs||_inf)"""
    Kl = K.astype(np.longdouble)
    r = np.asarray(Kl @ c.astype(np.longdouble) - rhs.astype(np.longdouble), dtype=np.float64)
    return float(np.max(np.abs(r)) / (np.abs(K).sum(1).max() * np.max(np.abs(c)) + np.max(np.abs(rhs))))
