#!/bin/bash
# A/B the blur kernel variants built by `python -m grayskull_b200.build --define=... --out=libvar_X.so`
for v in libgrayskull_b200 libvar_nototimad; do
  export GS_B200_LIB=$PWD/grayskull_b200/$v.so
  ok=$(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "stencils or blur_constant" 2>&1 | tail -1)
  timeout 300 python bench.py --steps 60 --warmup 10 --no-cpu --no-e2e > gpurun_out/ab_$v.json 2>gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$v.json"))
k=d["kernels"]["gs_blur_r5"]
print("%-18s blur %.3f ms  frac %.3f | sobel %.3f ms | sm %s MHz | tests: $ok" % ("$v", k["ms"], k["frac"], d["kernels"]["gs_sobel"]["ms"], d["clocks"]["sm_mhz"]))
PY
done
