"""GPU tests of the P x Q block-cyclic path (updes_b200/grid2d.py) through the C-ABI (CudaKernels2D).

The 1 x 1 grid runs on any GPU box: it goes through every CUDA wrapper (tile assembly, panel factorisation in the
gathered-panel buffer, TRSM / GEMM against the panel-row and U12 staging buffers, block GEMV, diagonal-block solves,
row equilibration).  2 x 1 / 1 x 2 need two GPUs, 2 x 2 four (gpurun --gpus N); the communication logic itself is
covered on CPUs by tests/test_grid2d_cpu.py."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FACETS = {"South": "n", "West": "d", "North": "d", "East": "d"}


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, P, Q, port, nx, ny, nb, out):
    import torch
    import torch.distributed as dist
    import updes_b200 as u
    from updes_b200 import assembly as asm
    from updes_b200.grid2d import BlockCyclic2D, CudaKernels2D, DistributedLU2D
    from updes_b200.linalg import LUFactorization
    world = P * Q
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        # ---- (1) random matrix, no ties, no equilibration: pivots must equal LAPACK's exactly ----------------
        n1 = nx * ny + 3
        g = torch.Generator().manual_seed(5)
        A = torch.randn((n1, n1), generator=g, dtype=torch.float64)
        bvec = torch.randn(n1, generator=g, dtype=torch.float64)
        layout = BlockCyclic2D(n1, nb, P, Q)
        d = DistributedLU2D(layout, rank, CudaKernels2D())
        d.fill_from_global(A.numpy())
        d.factor()
        x = d.solve(bvec.numpy())
        lu_ref, piv_ref = torch.linalg.lu_factor(A)
        piv_same = bool(np.array_equal(d.ipiv, piv_ref.numpy().astype(np.int64) - 1))
        npad = layout.nblocks * nb
        Fp = torch.zeros((npad, npad), dtype=torch.float64, device="cuda")
        Fp.view(layout.nblocks, nb, layout.nblocks, nb)[d.p::P, :, d.q::Q, :] = \
            d.local.view(d.mloc // nb, nb, d.ld // nb, nb)
        dist.all_reduce(Fp)
        fac_err = float((Fp[:n1, :n1].cpu() - lu_ref).abs().max() / lu_ref.abs().max())
        pad_clean = bool((Fp[n1:, :] == 0).all() and (Fp[:, n1:] == 0).all())
        sol_err = float((x.cpu() - torch.linalg.solve(A, bvec)).abs().max() / x.abs().max())
        zp = d.zero_pivot()
        d.close()
        del d

        # ---- (2) the collocation problem, row-equilibrated, against the single-GPU product path -----------------
        cloud = u.SquareCloud(Nx=nx, Ny=ny, facet_types=FACETS)
        M, n = 3, cloud.N + 3
        coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1))
        rows = asm.DeviceRows(cloud, asm.build_operator_rows(cloud, coef))
        xy = cloud.sorted_nodes
        q = np.zeros(n)
        north = np.asarray(cloud.facet_nodes["North"])
        q[north] = np.sin(np.pi * xy[north, 0])
        b = torch.as_tensor(q).cuda()
        K = asm.assemble_system(rows, "polyharmonic", 1.0, M)
        K0 = K.clone()
        ref = LUFactorization(K, n).factor(equilibrate=True)
        xref = ref.solve(b.clone())
        res = []
        for stage_u in (False, True):
            layout = BlockCyclic2D(n, nb, P, Q)
            d = DistributedLU2D(layout, rank, CudaKernels2D())
            d.always_stage_u = stage_u
            d.assemble(rows, "polyharmonic", 1.0, M)
            Kp = torch.zeros((layout.nblocks * nb,) * 2, dtype=torch.float64, device="cuda")
            Kp[:n, :n] = K0[:, :n]
            mine = Kp.view(layout.nblocks, nb, layout.nblocks, nb)[d.p::P, :, d.q::Q, :].reshape(d.mloc, d.ld)
            tiles_equal = bool(torch.equal(mine, d.local))
            d.equilibrate().factor()
            xs = d.solve(b)
            r = b - asm.apply_rows(rows, "polyharmonic", 1.0, M, xs.view(1, -1))[0]
            knorm = float(K0[:, :n].abs().sum(dim=1).max().item())
            berr = float(r.abs().max().item() / (knorm * xs.abs().max().item() + b.abs().max().item()))
            diff = float((xs - xref).abs().max().item() / xref.abs().max().item())
            status = d.zero_pivot()
            d.K.check_sweeps()
            res += [float(tiles_equal), berr, diff, float(status)]
            d.close()
            del d
        allres = [None] * world
        dist.all_gather_object(allres, [float(piv_same), fac_err, float(pad_clean), sol_err, float(zp)] + res)
        if rank == 0:
            np.save(out, np.array(allres))
    finally:
        dist.destroy_process_group()


def _run(P, Q, nx, ny, nb, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < P * Q:
        pytest.skip("needs %d GPUs" % (P * Q))
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(P, Q, _free_port(), nx, ny, nb, out), nprocs=P * Q, join=True)
    r = np.load(out)
    assert np.all(r[:, 0] == 1.0), "pivots differ from LAPACK partial pivoting on a tie-free matrix"
    assert np.all(r[:, 1] <= 1e-10), r          # factors == LAPACK's
    assert np.all(r[:, 2] == 1.0), "padding of the ragged last block must stay zero"
    assert np.all(r[:, 3] <= 1e-8), r           # solution of the random system
    assert np.all(r[:, 4] == 0)
    for o in (5, 9):                            # U12 from the local matrix / through the staging buffer
        assert np.all(r[:, o] == 1.0), "2-D tile assembly must be bit-identical to the single-GPU assembly"
        assert np.all(r[:, o + 1] <= 1e-13), r  # backward error (north_star: cond-scaled 1e-13)
        assert np.all(r[:, o + 2] <= 1e-6), r   # same discrete solution as the single-GPU path (cond ~ 1e9)
        assert np.all(r[:, o + 3] == 0)


@pytest.mark.parametrize("nx,ny,nb", [(30, 20, 32), (50, 50, 128)])
def test_grid2d_1x1(tmp_path, nx, ny, nb):
    _run(1, 1, nx, ny, nb, tmp_path)


@pytest.mark.parametrize("P,Q", [(2, 1), (1, 2)])
def test_grid2d_two_gpus(tmp_path, P, Q):
    _run(P, Q, 40, 40, 64, tmp_path)


@pytest.mark.parametrize("nx,nb", [(40, 64), (64, 256)])
def test_grid2d_2x2(tmp_path, nx, nb):
    _run(2, 2, nx, nx, nb, tmp_path)
