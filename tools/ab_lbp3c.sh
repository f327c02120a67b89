#!/bin/bash
# k_lbp_scan3 with column-parity planes: parity first, then threads x flat threshold x tile budget
export GS_B200_LBP_TMA=1
timeout 300 python tools/dbg_lbp.py 2>&1 | tail -1
ok=$(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "lbp or c4 or c5 or golden" 2>&1 | tail -1)
echo "parity (TMA planes): $ok"
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
run p_t256_flat64
GS_B200_LBP_FLAT=0 run p_t256_flat0
GS_B200_LBP_FLAT=128 run p_t256_flat128
GS_B200_LBP_TILE_KB=72 run p_t256_f64_72k
export GS_B200_LIB=$PWD/grayskull_b200/libv_t512.so
GS_B200_LBP_FLAT=128 run p_t512_flat128
unset GS_B200_LIB; unset GS_B200_LBP_TMA
run v2
