#!/bin/bash
# compute-sanitizer over the GPU parity suite (SURVEY.md section 5 hook).  Run on the GPU box:
#   gpurun --timeout 1500 -- tools/sanitize.sh [memcheck|racecheck|synccheck ...]
# Writes gpurun_out/sanitizer_<tool>.log (full) and gpurun_out/sanitizer_summary.txt.
# The full-size configs (c2 4096^2, c3 FHD, c4 UHD, last-frame) run under memcheck only at their reduced batches;
# racecheck / synccheck (shared-memory hazards, barrier misuse: the TMA + mbarrier kernels) run on
# the small-image tests, which take every kernel of the library.
mkdir -p gpurun_out
TOOLS=${@:-memcheck racecheck synccheck}
SMALL='golden or stencils or fast or orb or lbp or integral or match or histogram or filter or resize or blobs or perspective or radius'
: > gpurun_out/sanitizer_summary.txt
for tool in $TOOLS; do
  sel="-m gpu"
  [ "$tool" != memcheck ] && sel="-m gpu -k \"($SMALL) and not c2 and not c3 and not c4 and not c5 and not overlay and not cli and not 4096 and not 1080p and not 2160p\""
  log=gpurun_out/sanitizer_$tool.log
  t0=$(date +%s)
  eval timeout 1300 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
      python -m pytest tests/test_gpu_parity.py $sel -q -x --timeout 1200 -p no:cacheprovider > $log 2>&1
  rc=$?
  t1=$(date +%s)
  {
    echo "== compute-sanitizer --tool $tool   rc=$rc   $((t1 - t0)) s"
    grep -E "passed|failed|error" $log | tail -2
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Misaligned|Uninitialized" $log | sort | uniq -c | head -20
  } >> gpurun_out/sanitizer_summary.txt
done
cat gpurun_out/sanitizer_summary.txt
