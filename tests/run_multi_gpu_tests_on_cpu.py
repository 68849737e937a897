#!/usr/bin/env python
"""Dry run, on the CPU emulation of the C-ABI (tests/cpu_abi_emulation.py) and gloo, of the `-m gpu` tests that spawn
one process per GPU: the 1 x Q driver (tests/test_gpu_distributed.py), the P x Q driver (tests/test_gpu_zx_grid2d.py)
and the public API on the sharded path.  Each spawned rank installs the emulation and then runs the test's own
`_worker` unchanged; the parent applies the test's own assertions (`_run` with mp.spawn redirected).  This checks the
host drivers and the test logic at world sizes the build container has no GPUs for -- not the kernels.

    python tests/run_multi_gpu_tests_on_cpu.py            # every parametrisation, world <= 4
    python tests/run_multi_gpu_tests_on_cpu.py --quick    # smallest case of the three world-size > 1 drivers only
"""
import importlib
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


QUICK = ("test_block_cyclic_lu_world2", "test_sharded_api_with_nodal_fields", "test_grid2d_2x2")


def _emulated(rank, module, fn_name, args):
    import torch
    torch.set_num_threads(2)                        # several ranks share the container's cores
    import cpu_abi_emulation as emu
    emu.install()
    mod = importlib.import_module(module)
    getattr(mod, fn_name)(rank, *args)


def main():
    import torch
    import torch.multiprocessing as mp
    os.environ["UPDES_EMULATED_GPUS"] = "8"
    os.environ.setdefault("OMP_NUM_THREADS", "2")
    import cpu_abi_emulation as emu
    emu.install()                                   # the parent only needs device_count(); workers install their own
    real_spawn = mp.spawn

    def spawn(fn, args=(), nprocs=1, join=True, **kw):
        return real_spawn(_emulated, args=(fn.__module__, fn.__name__, args), nprocs=nprocs, join=join)

    mp.spawn = spawn
    ran = 0
    for modname in ("test_gpu_distributed", "test_gpu_zx_grid2d", "test_gpu_zz_sharded_api_fields"):
        mod = importlib.import_module(modname)
        for name in sorted(n for n in dir(mod) if n.startswith("test_")):
            if "--quick" in sys.argv and name not in QUICK:
                continue
            fn = getattr(mod, name)
            params = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
            cases = params[0].args[1] if params else [()]
            for vals in (cases[:1] if "--quick" in sys.argv else cases):
                vals = vals if isinstance(vals, tuple) else (vals,)
                t0 = time.perf_counter()
                with tempfile.TemporaryDirectory() as td:
                    fn(Path(td), *vals)
                ran += 1
                print("ok   ", modname, name, vals, "%.1f s" % (time.perf_counter() - t0), flush=True)
    print("%d multi-process GPU test cases hold on the emulated ABI + gloo" % ran)


if __name__ == "__main__":
    main()
