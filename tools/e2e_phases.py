"""Where the wall time of ONE public-API call goes (GPU box): pde_solver_jit on a SquareCloud with numpy inputs, phases
from updes_b200.operators.TRACE (device synchronize at every mark).  usage: e2e_phases.py [nx ...]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import updes_b200 as u
from updes_b200 import operators as ops

out = []
for nx in [int(a) for a in sys.argv[1:]] or [300]:
    cloud = u.SquareCloud(Nx=nx, Ny=nx, facet_types={"South": "n", "West": "d", "North": "d", "East": "d"})
    xy = cloud.sorted_nodes
    north = np.asarray(cloud.facet_nodes["North"])
    op = lambda xx, center, rbf, monomial, fields: u.nodal_laplacian(xx, center, rbf, monomial)
    rhs = lambda xx, centers, rbf, fields: 0.0
    bcs = {"South": np.zeros(len(cloud.facet_nodes["South"])), "West": np.zeros(len(cloud.facet_nodes["West"])),
           "North": np.sin(np.pi * xy[north, 0]), "East": np.zeros(len(cloud.facet_nodes["East"]))}
    for rep in range(3):
        u.clear_cache()
        ops.TRACE = [] if rep else None
        import torch
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ops._trace_t[0] = t0
        sol = u.pde_solver_jit(op, rhs, cloud, bcs, u.polyharmonic, 1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rep:
            out.append({"nx": nx, "n": cloud.N + 3, "total_ms": round(dt * 1e3, 3),
                        "phases_ms": [(k, round(v * 1e3, 3)) for k, v in ops.TRACE]})
    ops.TRACE = None
    u.clear_cache()
print(json.dumps(out, indent=1))
