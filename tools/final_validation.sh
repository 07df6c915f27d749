#!/bin/bash
# One GPU-box pass that produces everything the round's numbers are quoted from (outputs under gpurun_out/).
# usage (from the repo root): gpurun --timeout 2400 -- 'bash tools/final_validation.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -2 | tee gpurun_out/r01_final_gpu_tests.log
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/r01_bench_n1_final.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r01_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_final.csv \
    python bench.py --steps 1 --warmup 1 --timesteps 3 --no-cpu-baseline --no-e2e --no-kernel-timing > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:denoiser_stream_kernel -s 1 -c 1 -f -o gpurun_out/prof_r01_stream_final \
    python bench.py --steps 1 --warmup 1 --timesteps 2 --no-vocoder --no-cpu-baseline --no-e2e --no-kernel-timing > /dev/null 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r01_bench_n1_final.json", "gpurun_out/r01_bench_reference_arm.json"):
    try:
        d = json.load(open(f))
        print(f, d.get("value"), d.get("ms_per_step"), d.get("e2e"), (d.get("roofline") or {}).get("frac"), d.get("cpu_baseline"), d.get("clocks"), d.get("gpu_launches"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
