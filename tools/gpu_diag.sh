#!/bin/bash
# GPU-box diagnosis helper: every GPU test in its own process (a CUDA fault poisons the context),
# first with the generic kernels forced, then the TMA kernels under compute-sanitizer.
mkdir -p gpurun_out
TESTS=$(python -m pytest tests/test_gpu_parity.py --collect-only -q -m gpu 2>/dev/null | grep "::")
echo "== generic kernels forced =="
for t in $TESTS; do
  GS_B200_FORCE_GENERIC=1 timeout 600 python -m pytest "$t" -q -x --timeout 500 -p no:cacheprovider > gpurun_out/diag_tmp.log 2>&1
  rc=$?
  echo "rc=$rc $t"
  if [ $rc -ne 0 ]; then grep -E "^E |Error|error|assert" gpurun_out/diag_tmp.log | head -12; fi
done
echo "== TMA kernels under compute-sanitizer =="
cat > /tmp/tma_min.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import grayskull_b200 as g
from grayskull_b200 import api
g.lib().gs_b200_set_device(0)
which = sys.argv[1]
src = torch.randint(0, 256, (2, 128, 256), dtype=torch.uint8, device="cuda")
if which == "sobel":
    out = api.sobel_batch(src)
elif which == "blur":
    out = api.blur_batch(src, 5)
torch.cuda.synchronize()
print(which, "ok", int(out.sum()))
PY
for k in sobel blur; do
  timeout 300 compute-sanitizer --tool memcheck python /tmp/tma_min.py $k > gpurun_out/sanitizer_$k.log 2>&1
  echo "sanitizer $k rc=$?"; grep -vE "^$" gpurun_out/sanitizer_$k.log | head -40
done
