#!/bin/bash
# Quick iteration pass: parity of the denoiser in every mode + determinism + bench (both tensor-core modes) [+ ncu of the stream kernel with NCU=1]
mkdir -p gpurun_out
tag=${1:-quick}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zzz_bench_config.py tests/test_gpu_conv_gemm.py -m gpu -q -x --timeout 300 -rP -k "not campnet" 2>&1 | grep -E "margin|passed|failed|Error|error|assert|FAILED" | head -40 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_tests.log
timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-gpu-baseline 2>gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("bench:", d["dtype"], round(d["value"]), round(d["ms_per_step"], 2), "e2e", round((d.get("e2e") or {}).get("value") or 0), "frac", (d.get("roofline") or {}).get("frac"), "launch_us", (d.get("roofline") or {}).get("avg_launch_us"))
    print(" breakdown:", json.dumps(d.get("breakdown")))
    a = d.get("alt_mode") or {}
    print(" alt:", a.get("mode"), a.get("value"), a.get("ms_per_step"), (a.get("roofline") or {}).get("frac"), a.get("error"), json.dumps(a.get("breakdown")))
    print(" clocks:", d.get("clocks"))
except Exception as e:
    print("bench unreadable:", e); print(open("gpurun_out/${tag}_bench.err").read()[-3000:])
PY
if [ -n "$NCU" ]; then
for m in $NCU; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:denoiser_stream_kernel -s 5 -c 1 -f -o gpurun_out/prof_${tag}_stream_$m \
      python bench.py --mode $m --steps 1 --warmup 1 --timesteps 6 --no-vocoder --no-e2e --no-cpu-baseline --no-eager-gpu-baseline --no-alt-mode --no-kernel-timing > /dev/null 2> gpurun_out/${tag}_ncu_$m.err
done
fi
