"""Full-size checks (BASELINE.json configs[3]: 300x300 = 90 000 nodes, n = 90 003, 64.8 GB matrix) through
size-independent properties, because no CPU reference can hold this system:
  * entry parity of the assembled matrix on a random sample of entries and on boundary / P^T rows,
    against the oracle's closed forms evaluated entry by entry;
  * normwise backward error of the factor + solve (matrix-free residual) <= 1e-13;
  * max-norm error against the analytic Laplace solution (demos/Laplace/00_laplace_with_rbf.py:109-110)."""
import ctypes

import numpy as np
import pytest

import updes_b200 as u
from updes_b200 import assembly as asm
from helpers import CONFIG1_FACETS

pytestmark = pytest.mark.gpu


def _oracle_entry(O, cloud, table, M, r, c):
    """K[r, c] from the oracle's jets (same composition rules as oracle.assemble_K, one entry at a time)."""
    N = cloud.N
    xy = cloud.sorted_nodes
    if r >= N:                                   # P^T rows
        return O.monomial_jet(r - N, xy[c])[0] if c < N else 0.0
    if c >= N:                                   # monomial columns
        v = float(np.dot(table.cpol1[r], O.monomial_jet(c - N, xy[table.p1[r]])))
        if table.p2[r] >= 0:
            v += float(np.dot(table.cpol2[r], O.monomial_jet(c - N, xy[table.p2[r]])))
        return v
    if table.skip[r] == c:
        return 0.0
    v = float(np.dot(table.cphi1[r], O.rbf_jet("polyharmonic", 1.0, xy[table.p1[r]], xy[c])))
    if table.p2[r] >= 0:
        v += float(np.dot(table.cphi2[r], O.rbf_jet("polyharmonic", 1.0, xy[table.p2[r]], xy[c])))
    return v


def test_headline_size_entry_parity_backward_error_and_analytic_solution(oracle):
    import torch
    from updes_b200.linalg import LUFactorization
    free, _ = torch.cuda.mem_get_info()
    if free < 70e9:
        pytest.skip("needs ~66 GB of free HBM")
    cloud = u.SquareCloud(Nx=300, Ny=300, facet_types=CONFIG1_FACETS)
    M, n = 3, cloud.N + 3
    coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1))
    table = asm.build_operator_rows(cloud, coef)
    rows = asm.DeviceRows(cloud, table)
    K = asm.assemble_system(rows, "polyharmonic", 1.0, M)
    # ---- sampled entry parity --------------------------------------------------------------------------
    rng = np.random.default_rng(0)
    rr = np.concatenate([rng.integers(0, n, 4000), rng.integers(cloud.Ni, n, 2000), np.arange(n - 3, n).repeat(50)])
    cc = np.concatenate([rng.integers(0, n, 4000), rng.integers(0, n, 2000), rng.integers(0, n, 150)])
    rr = np.concatenate([rr, np.arange(0, 2000, 7)]); cc = np.concatenate([cc, np.arange(0, 2000, 7)])   # diagonal (Q1)
    got = K[torch.as_tensor(rr).cuda(), torch.as_tensor(cc).cuda()].cpu().numpy()
    want = np.array([_oracle_entry(oracle, cloud, table, M, int(r), int(c)) for r, c in zip(rr, cc)])
    urows = np.unique(rr)
    rowmax = K[torch.as_tensor(urows).cuda(), :n].abs().max(dim=1).values.cpu().numpy()      # row-scale floor
    scale = np.maximum(np.abs(want), rowmax[np.searchsorted(urows, rr)])
    assert np.max(np.abs(got - want) / np.where(scale == 0, 1.0, scale)) <= 1e-12
    assert torch.all(K[:, n:] == 0)
    # ---- factor + solve ----------------------------------------------------------------------------------
    xy = cloud.sorted_nodes
    q = np.zeros(n)
    north = np.asarray(cloud.facet_nodes["North"])
    q[north] = np.sin(np.pi * xy[north, 0])
    b = torch.as_tensor(q).cuda()
    knorm = float(K[:, :n].abs().sum(dim=1).max().item())
    lu = LUFactorization(K, n).factor()
    x = lu.solve(b.clone())
    assert lu.zero_pivot() == 0
    piv = lu.ipiv.cpu().numpy()
    assert np.all(piv >= np.arange(n)) and np.all(piv < n)            # partial pivoting picks rows at or below the diagonal
    r = b - asm.apply_rows(rows, "polyharmonic", 1.0, M, x.view(1, -1))[0]
    berr = float(r.abs().max().item() / (knorm * x.abs().max().item() + b.abs().max().item()))
    assert berr <= 1e-13, berr                                         # north_star: cond-scaled backward error 1e-13
    own = torch.arange(cloud.N, dtype=torch.int32, device="cuda")
    jphi, jpol = asm.eval_jets("polyharmonic", 1.0, rows.centres, x.view(1, -1), rows.centres, own)
    vals = (jphi[0, :, 0] + jpol[0, :, 0]).cpu().numpy()
    exact = np.sin(np.pi * xy[:, 0]) * np.cosh(np.pi * xy[:, 1]) / np.cosh(np.pi)
    assert np.max(np.abs(vals - exact)) <= 1e-4
    # factor-once / solve-many: a second right-hand side reuses the factors; linearity of the solve
    x2 = lu.solve((2.0 * b).clone())
    assert float((x2 - 2.0 * x).abs().max() / x.abs().max()) <= 1e-9
    # ---- row equilibration at the headline size: same system, rows scaled by exact powers
    # of two before pivoting.  Reported: how many pivots change; asserted: the backward error stays <= 1e-13.
    x_plain = x.clone()
    del x2
    asm.assemble_system(rows, "polyharmonic", 1.0, M, out=K)
    lu2 = LUFactorization(K, n).factor(equilibrate=True)
    x = lu2.solve(b.clone())
    assert lu2.check() == 0
    same = float(np.mean(lu2.ipiv.cpu().numpy() == piv))
    r = b - asm.apply_rows(rows, "polyharmonic", 1.0, M, x.view(1, -1))[0]
    berr2 = float(r.abs().max().item() / (knorm * x.abs().max().item() + b.abs().max().item()))
    print("headline size: backward error plain %.2e, equilibrated %.2e; %.1f %% of the pivots unchanged; "
          "solutions differ by %.2e" % (berr, berr2, 100 * same, float((x - x_plain).abs().max() / x_plain.abs().max())))
    assert berr2 <= 1e-13, berr2
