"""Where the wall time of the small configurations goes (GPU box): every public-API call of config 3's projection
loop and config 1 timed with a device synchronize on both sides, with launch counts and the per-class kernel
milliseconds of the library's profiler.  Exploration tool."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import updes_b200 as u
from updes_b200 import _lib
import configs
from helpers import cloud_from_golden

LOG = []


def timed(name, fn):
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    prof = {k: round(_lib.profile_read(k)[0], 3) for k in ("gemm", "panel", "swap", "trsm", "assemble", "solve")}
    _lib.profile_enable(False)
    LOG.append({"call": name, "wall_ms": round(dt, 3), "launches": _lib.launch_count() - l0, "kernel_ms": prof})
    return out


def wrap(mod, name):
    f = getattr(mod, name)

    def g(*a, **k):
        return timed(name, lambda: f(*a, **k))
    return g


class Proxy:
    def __init__(self, mod, names):
        self._m = mod
        for n in names:
            setattr(self, n, wrap(mod, n))

    def __getattr__(self, k):
        return getattr(self._m, k)


if __name__ == "__main__":
    cv, _ = cloud_from_golden("mesh_msh_cloud_vel.npz")
    cp, _ = cloud_from_golden("mesh_msh_cloud_phi.npz")
    configs.config3_projection_loop(u, cv, cp, nb_iter=1)          # warm (module load, first launches)
    u.clear_cache()
    P = Proxy(u, ["pde_solver_jit_with_bc", "interpolate_field", "gradient_vec"])
    t0 = time.perf_counter()
    configs.config3_projection_loop(P, cv, cp, nb_iter=2)
    total = (time.perf_counter() - t0) * 1e3
    out = {"config3_two_iterations_ms": round(total, 2), "calls": list(LOG)}
    LOG.clear()
    cloud, solve = configs.config1(u)
    solve(); u.clear_cache()
    timed("config1 pde_solver_jit (cold cache)", solve)
    timed("config1 pde_solver_jit (cached factor)", solve)
    out["config1"] = list(LOG)
    print(json.dumps(out))
