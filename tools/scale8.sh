#!/bin/bash
# 8-GPU session (gpurun --gpus 8): block-width A/B on the strong-scaling problem, then the full line with config5.
set -u
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
for nb in 512 1024; do
  timeout 300 $RUN 2952$((nb/512)) bench.py --gpus 8 --steps 3 --warmup 2 --nb $nb --e2e-steps 0 --config5 off --no-timeline \
    > $O/r02_bench_8gpu_nb$nb.json 2> $O/r02_bench_8gpu_nb$nb.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$O/r02_bench_8gpu_nb$nb.json") if l.startswith("{")][-1]
    print("nb=$nb", d["ms_per_step"], d["phase_ms_per_step"], d["roofline"]["achieved"], d["breakdown"])
except Exception as e:
    print("nb=$nb failed", e)
PY
done
NCCL_MAX_CTAS=4 timeout 300 $RUN 29528 bench.py --gpus 8 --steps 3 --warmup 2 --nb 512 --e2e-steps 0 --config5 off --no-timeline \
  > $O/r02_bench_8gpu_nb512_ctas4.json 2> $O/r02_bench_8gpu_nb512_ctas4.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$O/r02_bench_8gpu_nb512_ctas4.json") if l.startswith("{")][-1]
    print("nb=512 NCCL_MAX_CTAS=4", d["ms_per_step"], d["phase_ms_per_step"], d["roofline"]["achieved"])
except Exception as e:
    print("ctas4 failed", e)
PY
timeout 700 $RUN 29530 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r02_bench_8gpu.json 2> $O/r02_bench_8gpu.err
tail -c 2500 $O/r02_bench_8gpu.json
tail -3 $O/r02_bench_8gpu.err
