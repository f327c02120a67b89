#!/bin/bash
# One pass of the round's reference measurements (run under gpurun, one GPU); results in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/rec_bench_c2.json 2> gpurun_out/rec_c2.err
timeout 900 python bench.py --workload c3 --steps 30 --warmup 5 > gpurun_out/rec_bench_c3.json 2> gpurun_out/rec_c3.err
timeout 1200 python bench.py --workload c4 --steps 5 --warmup 2 > gpurun_out/rec_bench_c4.json 2> gpurun_out/rec_c4.err
timeout 900 python bench.py --workload c5 --steps 5 --warmup 2 --batch 1024 > gpurun_out/rec_bench_c5.json 2> gpurun_out/rec_c5.err
timeout 600 python bench.py --workload ops --steps 30 --warmup 5 --no-cpu > gpurun_out/rec_bench_ops.json 2> gpurun_out/rec_ops.err
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c3.csv $B --workload c3 --batch 64 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c4.csv $B --workload c4 --batch 32 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lbp_scan2 -s 1 -c 1 -f -o gpurun_out/prof_lbp $B --workload c4 --batch 32 > gpurun_out/ncu_lbp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fast_score_tiled -s 1 -c 1 -f -o gpurun_out/prof_fast $B --workload c3 --batch 32 > gpurun_out/ncu_fast.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_integral_bands -s 1 -c 1 -f -o gpurun_out/prof_integral $B --workload c4 --batch 32 > gpurun_out/ncu_int.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_orb_brief -s 1 -c 1 -f -o gpurun_out/prof_describe $B --workload c3 --batch 32 > gpurun_out/ncu_desc.log 2>&1
for f in c2 c3 c4 c5; do python - <<PY
import json
d=json.load(open("gpurun_out/rec_bench_$f.json"))
print("$f", "value %.4g %s" % (d["value"], d["unit"]), "ms/step %.3f" % d["ms_per_step"], "clocks", d["clocks"], "e2e", (d["e2e"] or {}).get("value"), "cpu", (d["cpu_baseline"] or {}).get("value"), (d["cpu_baseline"] or {}).get("cores"))
print("   ", {k: (round(v["ms"],3), round(v["frac"],3)) for k,v in d["kernels"].items()})
PY
done
