#!/usr/bin/env python
"""GPU check of the P x Q block-cyclic path (updes_b200/grid2d.py) against the single-GPU product path.

Stand-alone:            python tools/grid2d_gpu_check.py                      (1 x 1 grid on one GPU: every CUDA wrapper of
                                                                                CudaKernels2D, no communication)
Under torchrun:         python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
                            tools/grid2d_gpu_check.py --grid 2x2 [--nx 100 --nb 256]
Prints one JSON line per case on rank 0: pivots identical to the single-GPU LU (row-equilibrated both), assembled tiles
bit-identical, solution difference, backward error, milliseconds."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FACETS = {"South": "n", "West": "d", "North": "d", "East": "d"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="1x1")
    ap.add_argument("--cases", default="30x20:32,50x50:128", help="comma list of NXxNY:nb")
    args = ap.parse_args()
    P, Q = (int(v) for v in args.grid.split("x"))
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29517")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    assert world == P * Q, "world size must equal P * Q"
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", rank=rank, world_size=world)
    import updes_b200 as u
    from updes_b200 import assembly as asm
    from updes_b200.grid2d import BlockCyclic2D, CudaKernels2D, DistributedLU2D
    from updes_b200.linalg import LUFactorization

    for case in args.cases.split(","):
        dims, nb = case.split(":")
        nx, ny = (int(v) for v in dims.split("x"))
        nb = int(nb)
        cloud = u.SquareCloud(Nx=nx, Ny=ny, facet_types=FACETS)
        M = 3
        n = cloud.N + M
        coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1))
        rows = asm.DeviceRows(cloud, asm.build_operator_rows(cloud, coef))
        xy = cloud.sorted_nodes
        q = np.zeros(n)
        north = np.asarray(cloud.facet_nodes["North"])
        q[north] = np.sin(np.pi * xy[north, 0])
        b = torch.as_tensor(q).cuda()
        # single-GPU product path (every rank computes it: small)
        K = asm.assemble_system(rows, "polyharmonic", 1.0, M)
        K0 = K.clone()
        ref = LUFactorization(K, n).factor(equilibrate=True)
        xref = ref.solve(b.clone())
        for stage_u in (False, True):
            layout = BlockCyclic2D(n, nb, P, Q)
            d = DistributedLU2D(layout, rank, CudaKernels2D())
            d.always_stage_u = stage_u
            d.assemble(rows, "polyharmonic", 1.0, M)
            # my tiles against the single-GPU assembly, bit for bit
            npad = layout.nblocks * nb
            Kp = torch.zeros((npad, npad), dtype=torch.float64, device="cuda")
            Kp[:n, :n] = K0[:, :n]
            mine = Kp.view(layout.nblocks, nb, layout.nblocks, nb)[d.p::P, :, d.q::Q, :].reshape(d.mloc, d.ld)
            tiles_equal = bool(torch.equal(mine, d.local))
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            d.equilibrate().factor()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            x = d.solve(b)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            piv_equal = bool(np.array_equal(d.ipiv, ref.ipiv.cpu().numpy().astype(np.int64)))
            r = b - asm.apply_rows(rows, "polyharmonic", 1.0, M, x.view(1, -1))[0]
            knorm = float(K0[:, :n].abs().sum(dim=1).max().item())
            berr = float(r.abs().max().item() / (knorm * x.abs().max().item() + b.abs().max().item()))
            diff = float((x - xref).abs().max().item() / xref.abs().max().item())
            status = d.zero_pivot()
            d.K.check_sweeps()
            if rank == 0:
                print(json.dumps({"grid": args.grid, "cloud": dims, "n": n, "nb": nb, "u12_through_staging": stage_u,
                                  "tiles_bit_identical": tiles_equal, "pivots_equal_single_gpu": piv_equal,
                                  "rel_diff_vs_single_gpu": diff, "backward_error": berr, "zero_pivot": status,
                                  "factor_ms": (t1 - t0) * 1e3, "solve_ms": (t2 - t1) * 1e3}), flush=True)
            d.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
