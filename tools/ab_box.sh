#!/bin/bash
# k_box_mid (transposed rolling sums) vs k_box_wide (prefix scan) on the per-op table; parity first
set -o pipefail
[ -z "$AB_ONLY" ] && timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "wide_radius or stencils_vs_oracle or fused_blur_sobel" 2>&1 | tail -5
run() { timeout 300 python bench.py --workload ops --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read())
for k in ('gs_blur_r7','gs_blur_r9','gs_blur_r15','gs_blur_r31','gs_adaptive_threshold_r5','gs_adaptive_threshold_r15'):
    print('%-8s %-28s %.3f ms  %.3f' % ('$1', k, d['kernels'][k]['ms'], d['kernels'][k]['frac']))"; }
run mid
for v in "$@"; do [ -f grayskull_b200/libv_$v.so ] && GS_B200_LIB=$PWD/grayskull_b200/libv_$v.so run $v; done
[ -n "$AB_WIDE" ] && GS_B200_BOX=wide run wide
