"""Factor cache of the public API: least-recently-used eviction BEFORE a new system is built.  Written after the
round's GPU minutes were spent (dry-run on the emulated C-ABI), hence its own late-ordered file."""
import numpy as np
import pytest

import updes_b200 as u
from helpers import CONFIG1_FACETS

pytestmark = pytest.mark.gpu


def test_cache_evicts_least_recently_used_before_building(monkeypatch):
    """room for a new system is made BEFORE it is built, by evicting least-recently-used factorisations until
    the prediction fits the budget; a cached left-hand side costs two sweeps, an evicted one a new factorisation."""
    from updes_b200 import _lib, operators as ops
    cloud = u.SquareCloud(Nx=32, Ny=26, facet_types=CONFIG1_FACETS)
    one = ops._System.predict_nbytes(cloud.N + 3)
    monkeypatch.setattr(ops, "cache_budget_bytes", lambda: int(2.5 * one))
    zero = lambda c: 0.0
    bcs = {"South": zero, "West": zero, "North": lambda c: np.sin(np.pi * c[0]), "East": zero}
    rhs = lambda x, centers, rbf, fields: 1.0

    def solve(k):          # a different left-hand side per k: lap(u) - k u
        op = lambda x, center, rbf, monomial, fields: u.nodal_laplacian(x, center, rbf, monomial) - float(k) * u.nodal_value(x, center, rbf, monomial)
        l0 = _lib.launch_count()
        sol = u.pde_solver_jit(op, rhs, cloud, bcs, u.polyharmonic, 1)
        return sol, _lib.launch_count() - l0

    u.clear_cache()
    s1, first = solve(1)
    _, again = solve(1)
    assert again < 0.5 * first and len(ops._CACHE) == 1                       # cache hit: no assembly, no factorisation
    solve(2)
    assert len(ops._CACHE) == 2
    solve(1)                                                                  # touch 1: now 2 is the least recently used
    solve(3)
    assert len(ops._CACHE) == 2, "the budget holds two systems"
    assert sum(s.nbytes() for s in ops._CACHE.values()) <= 2.5 * one
    _, hit = solve(1)
    _, miss = solve(2)
    assert hit < 0.5 * first and miss >= 0.8 * first, (first, hit, miss)      # 1 survived, 2 was evicted and is rebuilt
    s1b, _ = solve(1)                                                         # (1 was evicted by rebuilding 2 after 3: rebuilt)
    assert np.max(np.abs(s1b.vals - s1.vals)) <= 1e-10 * np.max(np.abs(s1.vals))
    u.clear_cache()
    assert len(ops._CACHE) == 0
