#!/bin/bash
# A/B the LBP scan variants (GS_B200_LIB selects the library)
for v in "$@"; do
  export GS_B200_LIB=$PWD/grayskull_b200/$v.so
  ok=$(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "lbp_vs or c4" 2>&1 | tail -1)
  timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --batch 32 --no-cpu > gpurun_out/ab_$v.json 2>gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$v.json"))
print("%-20s %.3e windows/s  lbp %.2f ms | tests: $ok" % ("$v", d["value"], d["kernels"]["gs_lbp_detect"]["ms"]))
PY
done
