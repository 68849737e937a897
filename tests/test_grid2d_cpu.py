"""gloo tests (CPU) of the P x Q block-cyclic driver updes_b200/grid2d.py: with the numpy kernels of
tests/numpy_kernels2d.py the factors, the pivots and the solutions must equal LAPACK's partial-pivoting LU of the
global matrix, on square and non-square grids, ragged last blocks, with and without row equilibration."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from updes_b200.grid2d import BlockCyclic2D, DistributedLU2D, compose_interchanges, permutation_from_pivots


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, P, Q, port, n, nb, seed, equil, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from numpy_kernels2d import NumpyKernels2D
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=P * Q)
    try:
        rng = np.random.default_rng(seed)
        K = rng.normal(size=(n, n))
        if equil:
            K *= (10.0 ** rng.integers(-6, 7, size=n))[:, None]          # rows of wildly different scale
        b = rng.normal(size=n)
        layout = BlockCyclic2D(n, nb, P, Q)
        kern = NumpyKernels2D()
        lu = DistributedLU2D(layout, rank, kern)
        lu.fill_from_global(K)
        if equil:
            lu.equilibrate()
        lu.factor()
        x = lu.solve(b).numpy().copy()
        x2 = lu.solve(2.0 * b).numpy().copy()                             # factor once, solve again
        pieces = [None] * (P * Q)
        dist.all_gather_object(pieces, (lu.local.numpy().copy(), lu.ipiv.copy(), kern.calls, lu.zero_pivot()))
        if rank == 0:
            npad = layout.nblocks * nb
            F = np.zeros((npad, npad))
            Fb = F.reshape(layout.nblocks, nb, layout.nblocks, nb)
            for r, (loc, _, _, _) in enumerate(pieces):
                p, q = layout.coords(r)
                Fb[p::P, :, q::Q, :] = loc.reshape(len(layout.row_blocks(p)), nb, len(layout.col_blocks(q)), nb)
            panels = [[c[1] for c in pc[2] if c[0] == "panel"] for pc in pieces]
            np.savez(out, F=F[:n, :n], pad_rows=F[n:, :], pad_cols=F[:, n:], ipiv=pieces[0][1], x=x, x2=x2, K=K, b=b,
                     same_piv=all(np.array_equal(pc[1], pieces[0][1]) for pc in pieces),
                     panels=np.array([len(v) for v in panels]), info=pieces[0][3],
                     scale=(lu.scale_global.numpy() if equil else np.ones(n)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P,Q,n,nb,equil", [
    (2, 2, 200, 32, False),     # square grid, ragged last block (200 = 6 * 32 + 8)
    (2, 3, 330, 32, False),     # P != Q, 11 blocks
    (3, 2, 257, 32, True),      # P > Q, one-row last block, row equilibration
    (2, 2, 256, 64, True),      # exact multiple of the block size
    (1, 3, 170, 32, False),     # degenerates to the 1 x Q layout
    (3, 1, 170, 32, False),     # P x 1: every panel gathered from all ranks
    (2, 2, 96, 32, False),      # fewer blocks than 2 per process
    (2, 4, 300, 32, True),      # the 8-GPU grid of SURVEY.md 8(e)
    (4, 2, 270, 32, False),
])
def test_grid2d_lu_matches_lapack(tmp_path, P, Q, n, nb, equil):
    import scipy.linalg as sla
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(P, Q, _free_port(), n, nb, 11, equil, out), nprocs=P * Q, join=True)
    r = np.load(out)
    K, b = r["K"], r["b"]
    Ks = K * r["scale"][:, None]
    lu, piv = sla.lu_factor(Ks)
    assert bool(r["same_piv"]), "every rank must end with the same pivot list"
    assert int(r["info"]) == 0
    assert np.array_equal(r["ipiv"], piv)
    assert np.allclose(r["F"], lu, rtol=1e-10, atol=1e-10 * np.abs(lu).max())
    assert not r["pad_rows"].any() and not r["pad_cols"].any(), "padding of the ragged last block must stay zero"
    xref = np.linalg.solve(K, b)
    assert np.allclose(r["x"], xref, rtol=1e-8, atol=1e-8 * np.abs(xref).max())
    assert np.allclose(r["x2"], 2.0 * r["x"], rtol=1e-12, atol=0)
    # panel k is factored on the diagonal owner (k % P, k % Q) and nowhere else
    nblocks = (n + nb - 1) // nb
    expect = np.zeros(P * Q, dtype=int)
    for k in range(nblocks):
        expect[(k % P) * Q + (k % Q)] += 1
    assert np.array_equal(r["panels"], expect)
    if equil:
        m = np.abs(K).max(axis=1) * r["scale"]
        assert np.all((m >= 1.0) & (m < 2.0)) and np.all(np.frexp(r["scale"])[0] == 0.5)


def test_layout_maps_2d():
    L = BlockCyclic2D(1000, 64, 2, 3)                # 16 blocks, last one 40 wide
    assert L.nblocks == 16 and L.width(15) == 40
    assert L.coords(4) == (1, 1) and L.rank_of(1, 1) == 4
    assert L.row_blocks(1) == [1, 3, 5, 7, 9, 11, 13, 15] and L.col_blocks(2) == [2, 5, 8, 11, 14]
    assert L.local_rows(1) == 8 * 64 and L.valid_rows(1) == 8 * 64 - 24 and L.valid_rows(0) == 8 * 64
    assert L.local_cols(0) == 6 * 64 and L.valid_cols(0) == 6 * 64 - 24 and L.valid_cols(2) == 5 * 64
    assert L.lrow(5) == 2 * 64 and L.lcol(5) == 1 * 64
    assert L.first_row_block_from(1, 4) == 5 and L.first_row_block_from(0, 4) == 4
    assert L.lrow_from(1, 4) == 2 * 64 and L.lrow_from(0, 15) == 8 * 64 and L.lrow_from(1, 16) == 8 * 64
    assert L.lcol_from(2, 12) == 4 * 64 and L.lcol_from(2, 15) == 5 * 64
    assert L.row_owner(64 * 5 + 3) == 1 and L.local_row_of(64 * 5 + 3) == 2 * 64 + 3
    assert sum(L.valid_rows(p) for p in range(2)) == 1000 and sum(L.valid_cols(q) for q in range(3)) == 1000
    with pytest.raises(ValueError):
        BlockCyclic2D(100, 64, 4, 1)                 # 2 blocks cannot feed 4 process rows
    with pytest.raises(ValueError):
        BlockCyclic2D(100, 48, 1, 1)


def test_compose_interchanges_equals_sequential_swaps():
    rng = np.random.default_rng(0)
    for trial in range(50):
        n, r0, w = 60, int(rng.integers(0, 20)), int(rng.integers(1, 17))
        piv = [int(rng.integers(r0 + t, n)) if rng.random() < 0.8 else r0 + t for t in range(w)]
        if trial % 5 == 0:
            piv = [min(n - 1, r0 + t + int(rng.integers(0, 3))) for t in range(w)]      # chains inside the diagonal block
        v = np.arange(n)
        for t, p in enumerate(piv):
            v[[r0 + t, p]] = v[[p, r0 + t]]
        moves = compose_interchanges(r0, piv)
        got = np.arange(n)
        for d, s in moves.items():
            got[d] = s
        assert np.array_equal(got, v)
        assert all(d != s for d, s in moves.items())
    ip = np.array([2, 1, 3, 3])
    v = np.arange(4)
    for k, p in enumerate(ip):
        v[[k, p]] = v[[p, k]]
    assert np.array_equal(permutation_from_pivots(ip), v)
