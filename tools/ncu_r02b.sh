#!/bin/bash
# Round-2 ncu evidence, second pass (after the RED epilogue and the rewritten 128-row TRSM kernel): launch lists +
# full captures of the kernels that changed.  ONE GPU, through gpurun.  Outputs under gpurun_out/.
set -u
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02_launches_default_c400.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-library --no-small --e2e-steps 0 > $O/r02_launches_default.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 30000 --csv --log-file $O/r02_launches_nx91.csv \
  python bench.py --nx 91 --steps 2 --warmup 1 --no-cpu --no-library --no-small --e2e-steps 0 > $O/r02_launches_nx91.log 2>&1
cap() {  # name regex target
  timeout 600 $NCU --set full --import-source on -k regex:$2 -s 1 -c 1 -f -o $O/r02_$1 python tools/ncu_targets.py $3 > $O/r02_ncu_$1.log 2>&1
  ncu -i $O/r02_$1.ncu-rep --page raw --csv > $O/r02_ncu_$1_raw.csv 2>/dev/null
  ncu -i $O/r02_$1.ncu-rep --page details > $O/r02_ncu_$1_details.txt 2>/dev/null
  ncu -i $O/r02_$1.ncu-rep --page source --csv > $O/r02_ncu_$1_source.csv 2>/dev/null
  [ "$1" = dgemm ] || rm -f $O/r02_$1.ncu-rep
}
cap dgemm dgemm_sub_kernel gemm
cap dgemm_k1024 dgemm_sub_kernel gemm1k
cap trsm trsm_base_big_kernel trsm
du -sh $O
