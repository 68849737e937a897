"""World-size-2 (and 3) gloo tests of the multi-GPU driver's host logic on CPUs: the column-block-cyclic
LU with look-ahead must reproduce LAPACK's partial-pivoting LU of the global matrix, and the
distributed solve must match.  Numerics come from tests/numpy_backend.py; the product path uses the
CUDA backend with the same driver (tests/test_gpu_distributed.py runs that on real GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from updes_b200.distributed import ColumnBlockCyclic, DistributedLU


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, nb, seed, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from numpy_backend import NumpyBackend
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        K = rng.normal(size=(n, n))
        b = rng.normal(size=n)
        layout = ColumnBlockCyclic(n, nb, world)
        be = NumpyBackend(layout, rank)
        be.fill_from_global(K)
        lu = DistributedLU(layout, rank, be).factor()
        x = lu.solve(b).numpy().copy()
        lu.solve_variant = "right"                       # first-generation column sweep: a second opinion
        x_right = lu.solve(b).numpy().copy()
        # gather the factored column blocks on rank 0
        pieces = [None] * world
        dist.all_gather_object(pieces, (be.local[:, :be.cols].copy(), be.ipiv.copy(), be.calls))
        if rank == 0:
            F = np.zeros((n, n))
            for r, (loc, _, _) in enumerate(pieces):
                for j in layout.local_blocks(r):
                    w, lc = layout.width(j), layout.local_offset(j)
                    F[:, j * nb:j * nb + w] = loc[:, lc:lc + w]
            np.savez(out, F=F, ipiv=pieces[0][1], x=x, x_right=x_right, K=K, b=b, same_piv=all(np.array_equal(p[1], pieces[0][1]) for p in pieces),
                     panels_r1=np.array([c[1] for c in pieces[1][2] if c[0] == "panel"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,nb", [(2, 200, 32), (2, 257, 64), (3, 330, 32), (2, 96, 32), (4, 301, 32), (2, 67, 64)])
def test_block_cyclic_lu_matches_lapack(tmp_path, world, n, nb):
    import scipy.linalg as sla
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, _free_port(), n, nb, 7, out), nprocs=world, join=True)
    r = np.load(out)
    lu, piv = sla.lu_factor(r["K"])
    assert bool(r["same_piv"]), "every rank must end with the same pivot list"
    assert np.array_equal(r["ipiv"], piv)
    assert np.allclose(r["F"], lu, rtol=1e-10, atol=1e-10)
    assert np.allclose(r["x"], np.linalg.solve(r["K"], r["b"]), rtol=1e-8, atol=1e-8)
    assert np.allclose(r["x_right"], r["x"], rtol=1e-9, atol=1e-10)      # left-looking and column-sweep solves agree
    # rank 1 factored exactly its own panels (global blocks 1, 1+world, ...)
    assert [int(v) // nb for v in r["panels_r1"]] == list(range(1, (n + nb - 1) // nb, world))


def _worker_equil(rank, world, port, n, nb, out):
    """Row equilibration + timeline + a sub-group whose ranks differ from the global ranks."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from numpy_backend import NumpyBackend
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        group = dist.new_group([1, 2])           # global ranks 1, 2 -> group ranks 0, 1; rank 0 sits out
        if rank == 0:
            return
        grank = dist.get_rank(group)
        rng = np.random.default_rng(3)
        K = rng.normal(size=(n, n)) * (10.0 ** rng.integers(-6, 7, size=n))[:, None]     # rows of wildly different scale
        b = rng.normal(size=n)
        layout = ColumnBlockCyclic(n, nb, 2)
        be = NumpyBackend(layout, grank)
        be.fill_from_global(K)
        lu = DistributedLU(layout, grank, be, group=group)
        lu.timeline = []
        lu.equilibrate().factor()
        x = lu.solve(b).numpy().copy()
        tl = lu.timeline_ms()
        if grank == 0:
            np.savez(out, x=x, K=K, b=b, scale=be.scale, nrec=len(tl), keys=sorted(set().union(*[set(r) for r in tl])))
    finally:
        dist.destroy_process_group()


def test_equilibrated_lu_on_a_subgroup_with_timeline(tmp_path):
    out = str(tmp_path / "res.npz")
    n, nb = 150, 32
    mp.spawn(_worker_equil, args=(3, _free_port(), n, nb, out), nprocs=3, join=True)
    r = np.load(out)
    K, b = r["K"], r["b"]
    assert np.allclose(r["x"], np.linalg.solve(K, b), rtol=1e-9, atol=1e-12)
    # exact powers of two that bring every row's largest magnitude into [1, 2)
    m = np.abs(K).max(axis=1) * r["scale"]
    assert np.all((m >= 1.0) & (m < 2.0)) and np.all(np.frexp(r["scale"])[0] == 0.5)
    assert int(r["nrec"]) == (n + nb - 1) // nb
    assert {"k", "wait", "update"} <= set(r["keys"].tolist()) and "panel" in r["keys"].tolist()


def test_layout_maps():
    L = ColumnBlockCyclic(1000, 64, 4)
    assert L.nblocks == 16 and L.width(15) == 40
    assert [L.owner(j) for j in range(6)] == [0, 1, 2, 3, 0, 1]
    assert L.local_blocks(3) == [3, 7, 11, 15] and L.local_cols(3) == 3 * 64 + 40
    assert L.local_offset(11) == 128
    assert L.first_local_block_after(2, 2) == 6 and L.first_local_block_after(2, 1) == 2
    assert L.first_local_block_after(3, 15) is None and L.local_offset_after(3, 15) == L.local_cols(3)
    assert L.local_offset_after(0, 0) == 64 and L.local_offset_after(1, 0) == 0
    assert sum(L.local_cols(r) for r in range(4)) == 1000
