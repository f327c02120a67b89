#!/bin/bash
# ncu captures (run under gpurun, one GPU).  Numbers printed under ncu are NOT bench values.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:k_box_tma -s 1 -c 1 -f -o gpurun_out/prof_blur $B --batch 32 > gpurun_out/ncu_blur.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c3.csv $B --workload c3 --batch 64 > gpurun_out/ncu_l3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c4.csv $B --workload c4 --batch 8 > gpurun_out/ncu_l4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lbp_scan -s 1 -c 1 -f -o gpurun_out/prof_lbp $B --workload c4 --batch 4 > gpurun_out/ncu_lbp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fast_score -s 1 -c 1 -f -o gpurun_out/prof_fast $B --workload c3 --batch 32 > gpurun_out/ncu_fast.log 2>&1
ls -la gpurun_out/*.ncu-rep
