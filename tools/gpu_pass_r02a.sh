#!/bin/bash
# Round 2, pass A: first GPU contact of the tf32 tensor-core mode + the benchmark-shape parity tests + bench + ncu of the stream kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_cond_encoder.py -m gpu -q -x --timeout 200 -rP -k "tf32 or ragged or dilation or whole_model" 2>&1 | grep -E "margin|passed|failed|Error|error|assert" | head -60 > gpurun_out/r02a_tf32_small.log
tail -3 gpurun_out/r02a_tf32_small.log
timeout 900 python -m pytest tests/test_gpu_zzz_bench_config.py -m gpu -q --timeout 300 -rP 2>&1 | grep -E "margin|passed|failed|Error|error|assert|FAILED" | head -60 > gpurun_out/r02a_bench_config.log
cat gpurun_out/r02a_bench_config.log
timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 > gpurun_out/r02a_gpu_tests.log
tail -4 gpurun_out/r02a_gpu_tests.log
timeout 500 python bench.py --steps 5 --warmup 3 2>gpurun_out/r02a_bench.err | tail -1 > gpurun_out/r02a_bench_n1.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02a_bench_n1.json"))
    print("bench:", d["dtype"], d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
    print(" breakdown:", d.get("breakdown"))
    a = d.get("alt_mode") or {}
    print(" alt:", a.get("mode"), a.get("value"), a.get("ms_per_step"), (a.get("roofline") or {}).get("frac"), a.get("error"))
    print(" eager:", d.get("eager_gpu_baseline")); print(" cpu:", d.get("cpu_baseline"))
except Exception as e:
    print("bench unreadable:", e); print(open("gpurun_out/r02a_bench.err").read()[-2000:])
PY
# ncu: one launch of the streamed kernel in each tensor-core mode (full set), at the bench shape
for m in tc_tf32 tc_bf16; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:denoiser_stream_kernel -s 5 -c 1 -f -o gpurun_out/prof_r02_stream_$m \
      python bench.py --mode $m --steps 1 --warmup 1 --timesteps 6 --no-vocoder --no-e2e --no-cpu-baseline --no-eager-gpu-baseline --no-alt-mode --no-kernel-timing > /dev/null 2> gpurun_out/r02a_ncu_$m.err
done
ls -la gpurun_out/*.ncu-rep
