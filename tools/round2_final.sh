#!/bin/bash
# Final verification of round 2 on one B200 (gpurun), trimmed to the GPU-minutes left: parity suite, smoke, the driver's
# bench line + reference arm, the per-op table, c5, the ncu launch list of the bench command and the ncu --set full
# captures of the kernels that changed since tools/round2_verify.sh last ran (k_box_mid, k_resize_tiled).
mkdir -p gpurun_out
TAG=${1:-r02f}
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench (default line: c2 + configs c3/c4 + shard + cpu baselines)"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo rc=$?
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; echo rc=$?
echo "== ops / c5"
timeout 600 python bench.py --workload ops --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_ops.json 2>/dev/null
timeout 600 python bench.py --workload c5 --steps 3 --warmup 2 --no-cpu > gpurun_out/${TAG}_bench_c5.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_default.json"))
print("c2  value %.4g %s ms/step %.3f e2e %.4g dropin %.4g roofline %s %.3f cpu %.4g (%d cores) variants %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["e2e_dropin"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d.get("ms_per_step_variants")))
for k,v in d["configs"].items(): print("%s  value %.4g %s ms/step %.3f  cpu %.4g  clocks %s" % (k, v["value"], v["unit"], v["ms_per_step"], v["cpu_baseline"]["value"], v["clocks"]["sm_mhz"]), {kk: round(vv["ms"],3) for kk,vv in v["kernels"].items()})
s=d["shard"]; print("shard", s["ms"], "incl %.1f excl %.1f overlapped %.1f Mpix/s" % (s["mpix_s_including_collectives"], s["mpix_s_excluding_collectives"], s["overlapped"]["mpix_s"]))
r=json.load(open("gpurun_out/${TAG}_bench_reference.json")); print("reference arm  value %.4g %s  cores %s" % (r["value"], r["unit"], r["cpu_baseline"]["cores"]))
try:
    e=json.load(open("gpurun_out/${TAG}_bench_c5.json")); print("c5     value %.4g %s ms/step %.3f" % (e["value"], e["unit"], e["ms_per_step"]), {kk: round(vv["ms"],3) for kk,vv in e["kernels"].items()})
except Exception as ex: print("c5 FAILED", ex)
e=json.load(open("gpurun_out/${TAG}_bench_ops.json"))
for k,v in e["kernels"].items(): print("%-28s %8.3f ms  %7.1f GB/s  frac %.3f" % (k, v["ms"], v["achieved_gbs"], v["frac"]))
PY
echo "== ncu launch list of the bench command (kernel share of the step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --shard-frames 64 --shard-steps 1 > /dev/null 2>&1; echo rc=$?
echo "== ncu --set full captures"
cap() { name=$1; regex=$2; shift 2; timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip 1 -c 1 -f -o gpurun_out/${TAG}_$name "$@" > /dev/null 2>&1; echo "$name rc=$?"; }
cap box_mid_blur15 "k_box_mid" python tools/prof_r2.py blur15 32
cap box_mid_adaptive15 "k_box_mid" python tools/prof_r2.py box15 32
cap resize_tiled "k_resize_tiled" python tools/prof_r2.py resize_odd 32
ls -la gpurun_out/${TAG}_*.ncu-rep
echo "== k_box_mid compile-time variants (A/B): unroll 2, IMAD packing"
AB_ONLY=1 bash tools/ab_box.sh bmu2 bmpi 2>&1 | tail -20
