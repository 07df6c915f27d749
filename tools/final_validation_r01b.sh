#!/bin/bash
# Second validation pass of round 1 (after the condition encoder and CampNet landed): everything the updated numbers are quoted from.
# usage (from the repo root): gpurun --timeout 160 -- 'bash tools/final_validation_r01b.sh'
mkdir -p gpurun_out
timeout 100 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -2 | tee gpurun_out/r01b_gpu_tests.log
timeout 100 python bench.py --steps 3 --warmup 3 2>gpurun_out/r01b_bench.err | tail -1 > gpurun_out/r01b_bench_n1.json
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/r01b_smoke.log
timeout 60 ncu --set full --clock-control none --import-source on -k regex:camp_attention_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_r01_attention_tc \
    python tools/campnet_bench.py --iters 1 --warmup 1 > /dev/null 2>&1
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r01b_bench_n1.json"))
    print(d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("cpu_baseline"), d.get("clocks"), d.get("gpu_launches"), d["breakdown"].get("cond_encoder_plus_mel_encoder_ms"))
except Exception as e:
    print("bench unreadable:", e)
PY
ls -la gpurun_out/prof_r01_attention_tc.ncu-rep 2>/dev/null
