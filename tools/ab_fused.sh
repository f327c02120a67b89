#!/bin/bash
# A/B of the fused blur->sobel kernel variants + one ncu capture (run on the GPU box)
mkdir -p gpurun_out
run() { timeout 300 python bench.py --workload ops --steps 10 --warmup 3 --no-cpu --batch 64 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']
print('%-10s' % '$1', ' '.join('%s %.3f ms %.3f' % (n, k[n]['ms'], k[n]['frac']) for n in ('gs_blur_sobel_r5','gs_blur_r5','gs_sobel') if n in k))"; }
run default
for v in bsu6 bsu3; do if [ -f grayskull_b200/libv_$v.so ]; then GS_B200_LIB=$PWD/grayskull_b200/libv_$v.so run $v; fi; done
