#!/bin/bash
# First GPU call of the next round: A/B the compile-time experiments written (but not measured) at the end of round 1.
# Build the variants in the authoring container first (they travel with the snapshot):
#   python -m grayskull_b200.build --define=GSB_LBP_FSHIFT=1 --out=libv_fshift.so
#   python -m grayskull_b200.build --define=GSB_RS_PIPE=1    --out=libv_rspipe.so
# then:  gpurun --timeout 900 -- tools/next_round_ab.sh
mkdir -p gpurun_out
echo "== new parity test of round 1's last commit"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "frame_chunks" 2>&1 | tail -1
c4() { timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --batch 32 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-12s c4 %.3e windows/s  lbp %.2f ms' % ('$1', d['value'], d['kernels']['gs_lbp_detect']['ms']))"; }
rs() { timeout 300 python bench.py --workload ops --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); v=d['kernels']['gs_resize_to_half']; print('%-12s resize %.3f ms frac %.3f' % ('$1', v['ms'], v['frac']))"; }
c4 default
rs default
if [ -f grayskull_b200/libv_fshift.so ]; then
  export GS_B200_LIB=$PWD/grayskull_b200/libv_fshift.so
  echo "fshift parity: $(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k 'lbp or c4 or c5 or golden' 2>&1 | tail -1)"
  c4 fshift
  GS_B200_LBP_TMA=0 c4 fshift_scan2
fi
if [ -f grayskull_b200/libv_rspipe.so ]; then
  export GS_B200_LIB=$PWD/grayskull_b200/libv_rspipe.so
  echo "rspipe parity: $(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k 'stencils or golden' 2>&1 | tail -1)"
  rs rspipe
fi
