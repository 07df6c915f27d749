#!/bin/bash
# Round 2, pass O: is the text -> wav e2e line sensitive to the dependent launch of the streamed kernel?  (r02n final: e2e 219.1 ms
# against 211.2 ms from a ready cond, with ONE warm-up call of the e2e path, i.e. with the graph capture inside the timed region.)
mkdir -p gpurun_out
NB="--steps 3 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline --no-alt-mode --no-campnet --no-train --no-kernel-timing"
for pdl in 0 1 0 1; do
  FSE_STREAM_PDL=$pdl python bench.py $NB 2>/dev/null | tail -1 > gpurun_out/r02o_bench_pdl$pdl.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r02o_bench_pdl$pdl.json"))
print("pdl=$pdl value %.2f ms  e2e(text) %.2f ms  e2e(from cond) %.2f ms  launches %d  clocks %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["from_cond"]["ms_per_step"], d["gpu_launches"], d["clocks"]["sm_mhz"]))
PY
done 2>&1 | tee gpurun_out/r02o_e2e_pdl_ab.log
