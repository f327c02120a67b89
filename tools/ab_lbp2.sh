#!/bin/bash
# LBP A/B, round 2b: pooled survivor lists (GW warps per list) and the carry-chain LBP code
c4() { timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --batch 32 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-22s c4 %.3e windows/s  lbp %.2f ms' % ('$1', d['value'], d['kernels']['gs_lbp_detect']['ms']))"; }
c4 default
for v in carry gw2 gw4 gw4c; do
  if [ -f grayskull_b200/libv_$v.so ]; then
    export GS_B200_LIB=$PWD/grayskull_b200/libv_$v.so
    echo "$v parity: $(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k 'lbp or c4 or c5 or golden' 2>&1 | tail -1)"
    c4 $v
    GS_B200_LBP_FLAT=32 c4 ${v}_flat32
    unset GS_B200_LIB
  fi
done
