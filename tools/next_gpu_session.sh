#!/bin/bash
# First GPU session after round 2 (everything below was written after the round-2 GPU budget was spent; the tests were
# dry-run on the CPU emulation of the C-ABI only: tests/run_gpu_tests_on_cpu.py, tests/run_multi_gpu_tests_on_cpu.py).  1 GPU:   gpurun --timeout 900 -- 'bash tools/next_gpu_session.sh one'
#                      4 GPUs:  gpurun --gpus 4 --timeout 900 -- 'bash tools/next_gpu_session.sh four'
#                      8 GPUs:  gpurun --gpus 8 --timeout 900 -- 'bash tools/next_gpu_session.sh eight'
set -u
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
case "${1:-one}" in
one)
  # CUDA path against the reference-produced golden vectors, explicit assemble_* wrappers, 1x1 grid
  timeout 600 python -m pytest tests/test_gpu_zy_reference_golden.py tests/test_gpu_zx_grid2d.py tests/test_gpu_zz_fullsize_entries.py \
    tests/test_gpu_zzz_jax_adapter.py tests/test_gpu_solver.py -q -m gpu -s > $O/r03_pytest_new.log 2>&1
  tail -5 $O/r03_pytest_new.log
  # ncu launch list of ONE whole default step (about 19 000 launches; per-launch times cold-cache and serialised: shares only)
  timeout 450 ncu --clock-control none --metrics gpu__time_duration.sum -c 25000 --csv --log-file $O/r03_launches_default.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-library --no-small --e2e-steps 0 > $O/r03_launches_default.log 2>&1
  python tools/launch_list_summary.py $O/r03_launches_default.csv "ncu launch list of one whole default step (n = 90 003)" > $O/r03_launches_default_summary.txt 2>&1 || true
  head -20 $O/r03_launches_default_summary.txt
  ;;
four)
  timeout 400 python -m pytest tests/test_gpu_zx_grid2d.py tests/test_gpu_distributed.py tests/test_gpu_zz_sharded_api_fields.py -q -m gpu > $O/r03_pytest_4gpu.log 2>&1
  tail -5 $O/r03_pytest_4gpu.log
  timeout 200 $RUN --nproc-per-node 4 --master-port 29541 tools/grid2d_gpu_check.py --grid 2x2 --cases 50x50:128,100x100:256 > $O/r03_grid2d_2x2.log 2>&1
  tail -4 $O/r03_grid2d_2x2.log
  for grid in 1x4 2x2; do      # same strong-scaling step on both layouts (1x4 = the measured default path)
    extra=""; [ $grid = 2x2 ] && extra="--grid 2x2"
    timeout 300 $RUN --nproc-per-node 4 --master-port 2954${grid:0:1} bench.py --gpus 4 --steps 3 --warmup 2 --e2e-steps 0 --no-timeline $extra \
      > $O/r03_bench_4gpu_$grid.json 2> $O/r03_bench_4gpu_$grid.err
    tail -c 1200 $O/r03_bench_4gpu_$grid.json
  done
  ;;
eight)
  timeout 300 $RUN --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 2 --e2e-steps 0 --no-timeline --config5 off --grid 2x4 \
    > $O/r03_bench_8gpu_2x4.json 2> $O/r03_bench_8gpu_2x4.err
  tail -c 1200 $O/r03_bench_8gpu_2x4.json
  timeout 500 $RUN --nproc-per-node 8 --master-port 29552 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r03_bench_8gpu.json 2> $O/r03_bench_8gpu.err
  tail -c 2500 $O/r03_bench_8gpu.json
  ;;
esac
