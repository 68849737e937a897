"""Multi-GPU test of the public API with NODAL FIELDS on the sharded path (`rhs_args` used to factor the full
interpolation matrix A on every rank): after enable_distributed() both K and A are sharded (1 x Q or P x Q), the
right-hand side evaluates value / gradient of two nodal fields, the operator takes a field-dependent coefficient, the
cloud has periodic, Dirichlet and Robin facets.  Two solves with different fields (second one: K cached, A cached) must
equal the single-GPU path.  Needs 2 / 4 GPUs; dry-run on the emulated C-ABI + gloo (tests/run_multi_gpu_tests_on_cpu.py)."""
import os
import socket
from functools import partial

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, grid, out):
    import torch
    import torch.distributed as dist
    import updes_b200 as u
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cloud = u.SquareCloud(Nx=14, Ny=12, facet_types={"South": "p1", "North": "p1", "West": "d", "East": "r"})
        DT = 1e-2
        rbf = partial(u.polyharmonic, a=1)

        def op(x, c, r, m, f):
            return u.nodal_value(x, c, r, m) / DT + f[0] * u.nodal_gradient(x, c, r, m)[0] - 0.1 * u.nodal_laplacian(x, c, r, m)

        def rhs(x, centers, rbf, fields):
            return u.value(x, fields[:, 0], centers, rbf) / DT + u.gradient(x, fields[:, 1], centers, rbf)[1]

        rng = np.random.default_rng(3)
        f0, g0, g1 = rng.normal(size=cloud.N), rng.normal(size=cloud.N), rng.normal(size=cloud.N)
        bcs = {"South": lambda p: 0.0, "North": lambda p: 0.0, "West": lambda p: p[1], "East": (lambda p: 1.0, 2.0)}
        solve = lambda k: u.pde_solver(op, rhs, cloud, bcs, rbf, 1, diff_args=[f0], rhs_args=[g0 * (k + 1), g1])
        u.clear_cache()
        singles = [solve(0), solve(1)]
        u.clear_cache()
        u.enable_distributed(grid=grid)
        sharded = [solve(0), solve(1)]
        u.disable_distributed()
        u.clear_cache()
        res = [float(np.max(np.abs(a.vals - b.vals)) / np.max(np.abs(b.vals))) for a, b in zip(sharded, singles)]
        allres = [None] * world
        dist.all_gather_object(allres, res)
        if rank == 0:
            np.save(out, np.array(allres))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,grid", [(2, None), (2, (2, 1)), (4, (2, 2))])
def test_sharded_api_with_nodal_fields(tmp_path, world, grid):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(world, _free_port(), grid, out), nprocs=world, join=True)
    r = np.load(out)
    assert r.shape == (world, 2) and np.all(r <= 1e-8), r
