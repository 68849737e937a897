#!/usr/bin/env python
"""Dry run of bench.py's GPU arms on the CPU emulation of the C-ABI (tests/cpu_abi_emulation.py): exercises the bench's
own Python -- problem set-up, timed loop, profile legs, e2e calls, timeline, config5 pass, JSON line -- at toy sizes in
a container without a GPU.  The numbers it prints mean nothing.

    python tests/run_bench_on_cpu.py --nx 30 --steps 2 --warmup 1 --cpu-sizes 30x20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
        tests/run_bench_on_cpu.py --gpus 4 --nx 30 --steps 1 --warmup 1 --config5 on --config5-nx 40
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

torch.set_num_threads(2)
import cpu_abi_emulation as emu  # noqa: E402

emu.install()
import bench  # noqa: E402

bench.main()
