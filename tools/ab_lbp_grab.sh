#!/bin/bash
# k_lbp_scan3: dynamic slot hand-out (GS_B200_LBP_GRAB slots per hand-out from a per-CTA counter, 0 = static slot = warp + k * nwarps)
# parity first (every LBP test under each setting that is a candidate default), then 32 UHD frames of the c4 workload
c4() { timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --batch 32 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-22s c4 %.3e windows/s  lbp %.2f ms' % ('$1', d['value'], d['kernels']['gs_lbp_detect']['ms']))"; }
for g in 1 2; do echo "parity grab=$g"; GS_B200_LBP_GRAB=$g timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lbp or c4 or c5 or smoke" 2>&1 | tail -2; done
c4 static
for g in 1 2 4 8; do GS_B200_LBP_GRAB=$g c4 grab$g; done
GS_B200_LBP_GRAB=1 GS_B200_LBP_FLAT=16 c4 grab1_flat16
GS_B200_LBP_GRAB=2 GS_B200_LBP_FLAT=16 c4 grab2_flat16
