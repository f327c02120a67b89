#!/bin/bash
# second A/B round for k_lbp_scan3: threads per CTA x flat-mode threshold (bench only; parity was green for all modes)
run() {
  timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --batch 32 --no-cpu > gpurun_out/ab3_$1.json 2>gpurun_out/ab3.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab3_$1.json"))
    print("%-16s %.3e windows/s  lbp %.2f ms" % ("$1", d["value"], d["kernels"]["gs_lbp_detect"]["ms"]))
except Exception as e:
    print("$1 failed", e, open("gpurun_out/ab3.err").read()[-600:])
PY
}
export GS_B200_LBP_TMA=1
export GS_B200_LIB=$PWD/grayskull_b200/libv_t256.so
GS_B200_LBP_FLAT=0 run t256_flat0
GS_B200_LBP_FLAT=64 run t256_flat64
GS_B200_LBP_FLAT=128 run t256_flat128
GS_B200_LBP_FLAT=128 GS_B200_LBP_TILE_KB=120 run t256_f128_120k
export GS_B200_LIB=$PWD/grayskull_b200/libv_t384.so
GS_B200_LBP_FLAT=96 run t384_flat96
ok=$(GS_B200_LBP_FLAT=64 GS_B200_LIB=$PWD/grayskull_b200/libv_t256.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "lbp or c4 or c5 or golden" 2>&1 | tail -1)
echo "parity t256 flat64: $ok"
