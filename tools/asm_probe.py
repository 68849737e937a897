"""Assembly kernel alone (for ncu): Laplace rows (closed-form radial Laplacian) and a general-jet operator."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import updes_b200 as u
from updes_b200 import assembly as asm
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 200
cloud = u.SquareCloud(Nx=nx, Ny=nx, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})
for coefrow in ([0, 0, 0, 1.0, 1.0], [1e4, 100.0, 3.0, -0.08, -0.05]):
    rows = asm.DeviceRows(cloud, asm.build_operator_rows(cloud, np.tile(coefrow, (cloud.Ni, 1))))
    K = asm.assemble_system(rows, "polyharmonic", 1.0, 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); asm.assemble_system(rows, "polyharmonic", 1.0, 3, out=K); e1.record(); torch.cuda.synchronize()
    n = cloud.N + 3
    print(coefrow, "ms", e0.elapsed_time(e1), "GB/s", 8.0 * n * n / e0.elapsed_time(e1) * 1e-6)
    del K
