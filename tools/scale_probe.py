"""Scale probe (GPU box): times assembly / LU / solve of the synthetic Laplace SquareCloud at several
sizes and reports TFLOP/s, GB/s and backward error.  Not a bench line -- exploration only."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import updes_b200 as u
from updes_b200 import assembly as asm, _lib
from updes_b200.linalg import LUFactorization

FACETS = {"South": "n", "West": "d", "North": "d", "East": "d"}


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


def run(nx, detail=False):
    cloud = u.SquareCloud(Nx=nx, Ny=nx, facet_types=FACETS)
    M = 3
    n = cloud.N + M
    coef = np.tile([0.0, 0, 0, 1.0, 1.0], (cloud.Ni, 1))
    rows = asm.DeviceRows(cloud, asm.build_operator_rows(cloud, coef))
    K = asm.assemble_system(rows, "polyharmonic", 1.0, M)          # warm-up + allocation
    torch.cuda.synchronize()
    e0 = ev(); asm.assemble_system(rows, "polyharmonic", 1.0, M, out=K); e1 = ev(); torch.cuda.synchronize()
    t_asm = e0.elapsed_time(e1)
    lu = LUFactorization(K, n)
    if PANEL_VARIANT is not None:
        lu.set_panel_variant(PANEL_VARIANT)
    lu.factor(); torch.cuda.synchronize()                      # warm-up: lazy kernel loading, attribute calls
    asm.assemble_system(rows, "polyharmonic", 1.0, M, out=K)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    l0 = _lib.launch_count()
    e0 = ev(); lu.factor(); e1 = ev(); torch.cuda.synchronize()
    t_lu = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    prof = {k: (round(_lib.profile_read(k)[0], 2), _lib.profile_read(k)[2]) for k in ("gemm", "panel", "swap", "trsm")}
    _lib.profile_enable(False)
    xy = cloud.sorted_nodes
    q = np.zeros(n)
    north = np.asarray(cloud.facet_nodes["North"])
    q[north] = np.sin(np.pi * xy[north, 0])
    b = torch.as_tensor(q).cuda()
    x = b.clone()
    e0 = ev(); lu.solve(x); e1 = ev(); torch.cuda.synchronize()
    t_solve = e0.elapsed_time(e1)
    # backward error, matrix-free
    r = b - asm.apply_rows(rows, "polyharmonic", 1.0, M, x.view(1, -1))[0]
    # ||K||_inf estimate: row sums via apply on |.| is not available matrix-free; use max |K| row-sum bound from a re-assembly
    asm.assemble_system(rows, "polyharmonic", 1.0, M, out=K)
    knorm = K[:, :n].abs().sum(dim=1).max().item()
    berr = r.abs().max().item() / (knorm * x.abs().max().item() + b.abs().max().item())
    own = torch.arange(cloud.N, dtype=torch.int32, device="cuda")
    jphi, jpol = asm.eval_jets("polyharmonic", 1.0, rows.centres, x.view(1, -1), rows.centres, own)
    vals = (jphi[0, :, 0] + jpol[0, :, 0]).cpu().numpy()
    exact = np.sin(np.pi * xy[:, 0]) * np.cosh(np.pi * xy[:, 1]) / np.cosh(np.pi)
    out = dict(nx=nx, n=n, asm_ms=round(t_asm, 3), asm_gbs=round(8.0 * n * n / t_asm * 1e-6, 1), lu_ms=round(t_lu, 2),
               lu_tflops=round(2 / 3 * n ** 3 / t_lu * 1e-9, 2), solve_ms=round(t_solve, 3),
               solve_gbs=round(8.0 * n * n / t_solve * 1e-6, 1), launches=launches, info=lu.zero_pivot(),
               backward_err=berr, max_err_vs_analytic=float(np.max(np.abs(vals - exact))), prof_ms_launches=prof)
    print(json.dumps(out), flush=True)
    del K, lu, rows
    torch.cuda.empty_cache()


PANEL_VARIANT = None

if __name__ == "__main__":
    import os
    if "PANEL_VARIANT" in os.environ:
        PANEL_VARIANT = int(os.environ["PANEL_VARIANT"])
    for nx in [int(a) for a in sys.argv[1:]] or [64, 128]:
        run(nx)
