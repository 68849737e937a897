"""GPU tests of the multi-GPU path through the C-ABI building blocks (CudaBackend).

world = 1 runs on any GPU box (it still goes through panel buffers, TRSM/GEMM against a separate
buffer, block sweeps and NCCL self-broadcasts); world = 2 needs two GPUs (gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gather_factors(layout, be, world, n, nb):
    """Full factored matrix (on every rank) from the owned column blocks."""
    import torch
    import torch.distributed as dist
    F = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    for j in layout.local_blocks(be.rank):
        w, lc = layout.width(j), layout.local_offset(j)
        F[:, j * nb:j * nb + w] = be.local[:, lc:lc + w]
    dist.all_reduce(F)
    return F


def _reconstruction_error(F, piv, K):
    """max |L U - P K| / max |K| for factors F and 0-based interchange list piv."""
    import torch
    n = K.shape[0]
    L = torch.tril(F, -1) + torch.eye(n, dtype=torch.float64, device=F.device)
    U = torch.triu(F)
    PK = K.clone()
    for k, p in enumerate(piv.tolist()):
        if p != k:
            PK[[k, p]] = PK[[p, k]]
    return float((L @ U - PK).abs().max() / K.abs().max()), float(L.abs().max())


def _worker(rank, world, port, nx, nb, out):
    import torch
    import torch.distributed as dist
    import updes_b200 as u
    from updes_b200 import assembly as asm
    from updes_b200.distributed import ColumnBlockCyclic, CudaBackend, DistributedLU
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        # ---- (1) random matrix, no ties: pivots must equal LAPACK's exactly ---------------------------
        n1 = nx * nx + 3
        g = torch.Generator().manual_seed(5)
        A = torch.randn((n1, n1), generator=g, dtype=torch.float64)
        layout = ColumnBlockCyclic(n1, nb, world)
        be = CudaBackend(layout, rank, gemm_sms_reserved=8)
        for j in layout.local_blocks(rank):
            w, lc = layout.width(j), layout.local_offset(j)
            be.local[:, lc:lc + w] = A[:, j * nb:j * nb + w].cuda()
        lu = DistributedLU(layout, rank, be).factor()
        bvec = torch.randn(n1, generator=g, dtype=torch.float64)
        x = lu.solve(bvec.numpy())
        lu.solve_variant = "right"                       # first-generation column sweep as a second opinion
        x_right = lu.solve(bvec.numpy())
        lu.solve_variant = "left"
        solves_agree = float((x - x_right).abs().max() / x.abs().max())
        lu_ref, piv_ref = torch.linalg.lu_factor(A)
        F = _gather_factors(layout, be, world, n1, nb)
        piv_same = bool(np.array_equal(be.ipiv.cpu().numpy(), piv_ref.numpy() - 1))
        fac_err = float((F.cpu() - lu_ref).abs().max() / lu_ref.abs().max())
        sol_err = float((x.cpu() - torch.linalg.solve(A, bvec)).abs().max() / x.abs().max())
        zp = be.zero_pivot()
        del be, lu

        # ---- (2) the collocation problem on a jittered cloud ----------------------------------------------
        cloud = u.SquareCloud(Nx=nx, Ny=nx, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"}, noise_key=3)
        M, n = 3, cloud.N + 3
        coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1))
        rows = asm.DeviceRows(cloud, asm.build_operator_rows(cloud, coef))
        layout = ColumnBlockCyclic(n, nb, world)
        be = CudaBackend(layout, rank, gemm_sms_reserved=8)
        be.assemble(rows, "polyharmonic", 1.0, M)
        Kfull = asm.assemble_system(rows, "polyharmonic", 1.0, M)[:, :n].contiguous()
        asm_same = True
        for j in layout.local_blocks(rank):
            w, lc = layout.width(j), layout.local_offset(j)
            asm_same = asm_same and bool(torch.equal(be.local[:, lc:lc + w], Kfull[:, j * nb:j * nb + w]))
        lu = DistributedLU(layout, rank, be).factor()
        xy = cloud.sorted_nodes
        q = np.zeros(n)
        north = np.asarray(cloud.facet_nodes["North"])
        q[north] = np.sin(np.pi * xy[north, 0])
        x = lu.solve(q)
        F = _gather_factors(layout, be, world, n, nb)
        rec_err, lmax = _reconstruction_error(F, be.ipiv.cpu().numpy(), Kfull)
        bq = torch.as_tensor(q).cuda()
        berr = float((Kfull @ x - bq).abs().max() / (Kfull.abs().sum(dim=1).max() * x.abs().max() + bq.abs().max()))
        res = [None] * world
        dist.all_gather_object(res, (float(piv_same), fac_err, sol_err, float(zp), float(asm_same), rec_err, lmax, berr,
                                     float(be.zero_pivot()), solves_agree))
        if rank == 0:
            np.save(out, np.array(res))
    finally:
        dist.destroy_process_group()


def _run(world, nx, nb, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(world, _free_port(), nx, nb, out), nprocs=world, join=True)
    r = np.load(out)
    assert np.all(r[:, 0] == 1.0), "pivots differ from LAPACK partial pivoting on a tie-free matrix"
    assert np.all(r[:, 1] <= 1e-10), r          # factors == LAPACK's
    assert np.all(r[:, 2] <= 1e-8), r           # solution
    assert np.all(r[:, 3] == 0)
    assert np.all(r[:, 4] == 1.0), "distributed assembly must be bit-identical to the single-GPU assembly"
    assert np.all(r[:, 5] <= 1e-12), r          # P K = L U
    assert np.all(r[:, 6] <= 1.0 + 1e-12), r    # partial pivoting: |L| <= 1
    assert np.all(r[:, 7] <= 1e-14), r          # backward error of the distributed solve
    assert np.all(r[:, 8] == 0)
    assert np.all(r[:, 9] <= 1e-10), r          # left-looking solve == column-sweep solve (well-conditioned matrix)


@pytest.mark.parametrize("nx,nb", [(24, 64), (40, 128), (50, 512)])
def test_block_cyclic_lu_world1(tmp_path, nx, nb):
    _run(1, nx, nb, tmp_path)


@pytest.mark.parametrize("nx,nb", [(24, 64), (40, 128), (64, 512)])
def test_block_cyclic_lu_world2(tmp_path, nx, nb):
    _run(2, nx, nb, tmp_path)


def _worker_api(rank, world, port, grid, out):
    """The public API on the sharded path: pde_solver_jit after enable_distributed(), called twice (the second
    call used to run out of memory because the first solution pinned its factors), against the single-GPU path."""
    import torch
    import torch.distributed as dist
    import updes_b200 as u
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cloud = u.SquareCloud(Nx=40, Ny=30, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})
        op = lambda x, center, rbf, monomial, fields: u.nodal_laplacian(x, center, rbf, monomial)
        rhs = lambda x, centers, rbf, fields: 0.0
        bcs = {"South": lambda c: 0.0, "West": lambda c: 0.0, "North": lambda c: np.sin(np.pi * c[0]), "East": lambda c: 0.0}
        single = u.pde_solver_jit(op, rhs, cloud, bcs, u.polyharmonic, 1)
        u.enable_distributed(grid=grid)
        sols = []
        for _ in range(2):
            u.clear_cache()
            sols.append(u.pde_solver_jit(op, rhs, cloud, bcs, u.polyharmonic, 1))
        cached = u.pde_solver_jit(op, rhs, cloud, bcs, u.polyharmonic, 1)            # third call: factor cache hit
        u.disable_distributed()
        scale = np.max(np.abs(single.vals))
        res = [float(np.max(np.abs(s.vals - single.vals)) / scale) for s in sols + [cached]]
        allres = [None] * world
        dist.all_gather_object(allres, res)
        if rank == 0:
            np.save(out, np.array(allres))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,grid", [(2, None), (2, (2, 1)), (4, (2, 2))])
def test_public_api_on_the_sharded_path(tmp_path, world, grid):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker_api, args=(world, _free_port(), grid, out), nprocs=world, join=True)
    r = np.load(out)
    assert np.all(r <= 1e-8), r          # same discrete solution as the single-GPU path on every rank, every call
