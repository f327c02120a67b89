#!/bin/bash
# A/B the gs_match_orb kernel variants (GS_B200_LIB selects the library)
for v in "$@"; do
  export GS_B200_LIB=$PWD/grayskull_b200/$v.so
  ok=$(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "match" 2>&1 | tail -1)
  timeout 300 python bench.py --workload match --steps 10 --warmup 3 --batch 1024 --no-cpu > gpurun_out/abm_$v.json 2>gpurun_out/abm.err
  python - <<PY
import json
d=json.load(open("gpurun_out/abm_$v.json"))
print("%-22s %.1f Gcmp/s  %.3f ms/step | tests: $ok" % ("$v", d["value"], d["ms_per_step"]))
PY
done
