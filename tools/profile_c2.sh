#!/bin/bash
# ncu captures for the c2 kernels (run under gpurun; one GPU).  Numbers printed by bench.py under
# ncu are NOT bench values.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --batch 32 --no-e2e --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c2.csv $B > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_box_tma -s 1 -c 1 -f -o gpurun_out/prof_blur $B > gpurun_out/ncu_blur.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stencil3_tma -s 1 -c 1 -f -o gpurun_out/prof_sobel $B > gpurun_out/ncu_sobel.log 2>&1
ls -la gpurun_out/*.ncu-rep
