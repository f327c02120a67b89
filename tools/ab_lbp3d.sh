#!/bin/bash
# k_lbp_scan3 with warp-autonomous survivor lists: parity first, then flat threshold / tile budget / threads
timeout 300 python tools/dbg_lbp.py 2>&1 | tail -1
ok=$(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "lbp or c4 or c5 or golden" 2>&1 | tail -1)
echo "parity (warp-autonomous): $ok"
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
run w_t512_flat16
GS_B200_LBP_FLAT=0 run w_t512_flat0
GS_B200_LBP_FLAT=8 run w_t512_flat8
GS_B200_LBP_FLAT=32 run w_t512_flat32
GS_B200_LBP_TILE_KB=140 run w_t512_f16_140k
GS_B200_LIB=$PWD/grayskull_b200/libv_t256.so run w_t256_flat16
GS_B200_LBP_TMA=0 run v2
