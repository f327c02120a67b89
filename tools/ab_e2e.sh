#!/bin/bash
# sweep the c2 end-to-end leg (pinned host -> H2D -> blur -> sobel -> D2H) over chunk size / stream count
for cfg in "16 2" "16 3" "16 4" "8 4" "32 3" "4 6"; do
  set -- $cfg
  timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu --e2e-chunk $1 --e2e-streams $2 --e2e-frames 96 > gpurun_out/e2e_$1_$2.json 2>gpurun_out/e2e.err
  python - <<PY
import json
d=json.load(open("gpurun_out/e2e_$1_$2.json"))
print("chunk %2d streams %d: e2e %.0f Mpix/s  (device-resident %.3e)" % ($1, $2, d["e2e"]["value"], d["value"]))
PY
done
