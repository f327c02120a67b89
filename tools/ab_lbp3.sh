#!/bin/bash
# A/B the LBP scan: k_lbp_scan2 (global gathers) vs k_lbp_scan3 (TMA-staged tiles): tile budgets, flat-mode threshold
run() {
  ok=$(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "lbp or c4 or c5 or golden" 2>&1 | tail -1)
  timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --batch 32 --no-cpu > gpurun_out/ab3_$1.json 2>gpurun_out/ab3.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab3_$1.json"))
    print("%-16s %.3e windows/s  lbp %.2f ms | tests: $ok" % ("$1", d["value"], d["kernels"]["gs_lbp_detect"]["ms"]))
except Exception as e:
    print("$1 failed", e, open("gpurun_out/ab3.err").read()[-600:], "| tests: $ok")
PY
}
run v2
export GS_B200_LBP_TMA=1
run v3_default
GS_B200_LBP_FLAT=0 run v3_flat0
GS_B200_LBP_FLAT=128 run v3_flat128
GS_B200_LBP_FLAT=512 run v3_flat512
GS_B200_LBP_TILE_KB=88 run v3_tile88
