#!/bin/bash
# Round-2 ncu evidence (run on the GPU box through gpurun, ONE GPU).  Outputs under gpurun_out/.
set -u
O=gpurun_out
NCU="ncu --clock-control none"
# 1. launch list of the default bench command (first 400 launches) and of a complete small run (n = 8 284)
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02_launches_default_c400.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-library --no-small --e2e-steps 0 > $O/r02_launches_default.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 30000 --csv --log-file $O/r02_launches_nx91.csv \
  python bench.py --nx 91 --steps 2 --warmup 1 --no-cpu --no-library --no-small --e2e-steps 0 > $O/r02_launches_nx91.log 2>&1
# 2. full captures of one warm launch of each hot kernel
cap() {  # name regex target
  timeout 600 $NCU --set full --import-source on -k regex:$2 -s 1 -c 1 -f -o $O/r02_$1 python tools/ncu_targets.py $3 > $O/r02_ncu_$1.log 2>&1
  ncu -i $O/r02_$1.ncu-rep --page raw --csv > $O/r02_ncu_$1_raw.csv 2>/dev/null
  ncu -i $O/r02_$1.ncu-rep --page details > $O/r02_ncu_$1_details.txt 2>/dev/null
  ncu -i $O/r02_$1.ncu-rep --page source --csv > $O/r02_ncu_$1_source.csv 2>/dev/null
  [ "$1" = dgemm ] || rm -f $O/r02_$1.ncu-rep      # gpurun_out is capped at 64 MiB: keep one report, the extracted pages of all
}
cap dgemm dgemm_sub_kernel gemm
cap dgemm_k1024 dgemm_sub_kernel gemm1k
cap sweep tri_sweep2_kernel sweep
cap panel lu_panel2_kernel panel
cap trsm trsm_base_big_kernel trsm
# assembly: two launches of interest (Laplace path = 1st group, general jet = 2nd group; 3 reps each, 2 kernels per rep)
timeout 600 $NCU --set full --import-source on -k regex:assemble_phi_kernel -c 8 -f -o $O/r02_assemble python tools/ncu_targets.py asm > $O/r02_ncu_assemble.log 2>&1
ncu -i $O/r02_assemble.ncu-rep --page raw --csv > $O/r02_ncu_assemble_raw.csv 2>/dev/null
ncu -i $O/r02_assemble.ncu-rep --page details > $O/r02_ncu_assemble_details.txt 2>/dev/null
rm -f $O/r02_assemble.ncu-rep
du -sh $O
ls -la $O | tail -30
