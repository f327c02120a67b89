#!/bin/bash
# k_lbp_scan3: per-warp mask stores (no final CTA barrier) vs the previous build; 1024-thread CTAs
ok=$(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "lbp or c4 or c5 or golden" 2>&1 | tail -1)
echo "parity (no final barrier): $ok"
run() {
  timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --batch 32 --no-cpu 2>gpurun_out/ab3.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-18s %.3e windows/s  lbp %.2f ms' % ('$1', d['value'], d['kernels']['gs_lbp_detect']['ms']))"
}
run nobarrier_t512
GS_B200_LIB=$PWD/grayskull_b200/libv_prev.so run prev_t512
GS_B200_LIB=$PWD/grayskull_b200/libv_t1024.so run nobarrier_t1024
GS_B200_LIB=$PWD/grayskull_b200/libv_t1024.so GS_B200_LBP_TILE_KB=150 run nobar_t1024_150k
ok=$(GS_B200_LIB=$PWD/grayskull_b200/libv_t1024.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "lbp or c4" 2>&1 | tail -1)
echo "parity t1024: $ok"
