#!/bin/bash
# end-of-round verification on one B200: parity suite, smoke, every bench workload, A/B leftovers, ncu.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
b() { timeout 600 python bench.py "$@" 2>gpurun_out/final.err; }
echo "== c2"; b > gpurun_out/final_c2.json; tail -c 300 gpurun_out/final.err | tail -1
echo "== c4"; b --workload c4 --steps 5 --warmup 3 > gpurun_out/final_c4.json
echo "== c3"; b --workload c3 --steps 10 --warmup 3 > gpurun_out/final_c3.json
echo "== c5"; b --workload c5 --steps 3 --warmup 3 --no-cpu > gpurun_out/final_c5.json
echo "== ops"; b --workload ops --steps 10 --warmup 3 --no-cpu > gpurun_out/final_ops.json
echo "== match"; b --workload match --steps 10 --warmup 3 > gpurun_out/final_match.json
echo "== tmatch"; b --workload tmatch --steps 5 --warmup 3 --no-cpu > gpurun_out/final_tmatch.json
python - <<'PY'
import json
for wl in ("c2","c4","c3","c5","match","tmatch"):
    try:
        d=json.load(open("gpurun_out/final_%s.json"%wl))
        print("%-6s value %.4g %s  ms/step %.3f  e2e %s  roofline %s frac %.3f  cpu %s clocks %s" % (wl, d["value"], d["unit"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), d["roofline"]["kernel"], d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value"), d["clocks"]))
    except Exception as e:
        print(wl, "FAILED", e)
d=json.load(open("gpurun_out/final_ops.json"))
for k,v in d["kernels"].items(): print("%-28s %8.3f ms  %7.1f GB/s  frac %.3f" % (k, v["ms"], v["achieved_gbs"], v["frac"]))
PY
echo "== resize A/B (GSB_RS_PAIRS=0)"
GS_B200_LIB=$PWD/grayskull_b200/libv_rs0.so timeout 300 python bench.py --workload ops --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); v=d['kernels']['gs_resize_to_half']; print('gather-only resize: %.3f ms frac %.3f' % (v['ms'], v['frac']))"
echo "== LBP tuning (c4, 32 frames)"
t() { name=$1; shift; env "$@" timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --batch 32 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-22s %.3e windows/s  lbp %.2f ms' % ('$name', d['value'], d['kernels']['gs_lbp_detect']['ms']))"; }
t default X=1
t flat4 GS_B200_LBP_FLAT=4
t flat6 GS_B200_LBP_FLAT=6
t flat12 GS_B200_LBP_FLAT=12
t cuts_every GS_B200_LBP_CUTS=1,2,3,4,5,6,7
t cuts_1_2_4 GS_B200_LBP_CUTS=1,2,4
t cuts_2_4 GS_B200_LBP_CUTS=2,4
t tile120 GS_B200_LBP_TILE_KB=120
echo "== ncu"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/final_launches_c4.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --workload c4 --batch 8 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_lbp_scan3 -s 15 -c 15 -f -o gpurun_out/prof_lbp3w python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --workload c4 --batch 4 > gpurun_out/ncu_lbp3w.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
