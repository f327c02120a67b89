#!/bin/bash
# LBP A/B on 32 UHD frames (c4 workload): which scales run the 1024-thread / 224 KB tile form
c4() { timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --batch 32 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-22s c4 %.3e windows/s  lbp %.2f ms' % ('$1', d['value'], d['kernels']['gs_lbp_detect']['ms']))"; }
c4 auto_rows32
GS_B200_LBP_BIG=0 c4 big_never
GS_B200_LBP_BIG=1 c4 big_always
GS_B200_LBP_BIG_ROWS=17 c4 auto_rows17
GS_B200_LBP_BIG_ROWS=9 c4 auto_rows9
