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
        cloud = u.SquareCloud(Nx=nx, Ny=nx, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})
        M, n = 3, cloud.N + 3
        coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1))
        rows = asm.DeviceRows(cloud, asm.build_operator_rows(cloud, coef))
        layout = ColumnBlockCyclic(n, nb, world)
        be = CudaBackend(layout, rank, gemm_sms_reserved=8)
        be.assemble(rows, "polyharmonic", 1.0, M)
        # assembled blocks == the single-GPU assembly
        Kfull = asm.assemble_system(rows, "polyharmonic", 1.0, M)
        for j in layout.local_blocks(rank):
            w, lc = layout.width(j), layout.local_offset(j)
            assert torch.equal(be.local[:, lc:lc + w], Kfull[:, j * nb:j * nb + w])
        lu = DistributedLU(layout, rank, be).factor()
        xy = cloud.sorted_nodes
        q = np.zeros(n)
        north = np.asarray(cloud.facet_nodes["North"])
        q[north] = np.sin(np.pi * xy[north, 0])
        x = lu.solve(q)
        torch.cuda.synchronize()
        # reference: single-GPU factorisation of the same matrix
        from updes_b200.linalg import LUFactorization
        ref = LUFactorization(Kfull, n).factor()
        xref = ref.solve(torch.as_tensor(q).cuda())
        piv_same = bool(torch.equal(ref.ipiv, be.ipiv))
        err = float((x - xref).abs().max() / xref.abs().max())
        fac_err = 0.0
        for j in layout.local_blocks(rank):
            w, lc = layout.width(j), layout.local_offset(j)
            d = (be.local[:, lc:lc + w] - Kfull[:, j * nb:j * nb + w]).abs().max() / Kfull[:, :n].abs().max()
            fac_err = max(fac_err, float(d))
        res = [None] * world
        dist.all_gather_object(res, (piv_same, err, fac_err, be.zero_pivot()))
        if rank == 0:
            np.save(out, np.array([[float(a), b, c, float(d)] for a, b, c, d in res]))
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
    assert np.all(r[:, 0] == 1.0), "pivot lists differ from the single-GPU factorisation"
    assert np.all(r[:, 1] <= 1e-9), r
    assert np.all(r[:, 2] <= 1e-11), r          # same factors up to GEMM blocking order
    assert np.all(r[:, 3] == 0)


@pytest.mark.parametrize("nx,nb", [(24, 64), (40, 128), (50, 512)])
def test_block_cyclic_lu_world1(tmp_path, nx, nb):
    _run(1, nx, nb, tmp_path)


@pytest.mark.parametrize("nx,nb", [(24, 64), (40, 128), (64, 512)])
def test_block_cyclic_lu_world2(tmp_path, nx, nb):
    _run(2, nx, nb, tmp_path)
