#!/bin/bash
# Round 2, final pass after the dependent-launch change: smoke(), the whole GPU suite with margins, the default bench line, the
# reference arm, the ncu launch list of one step, fresh `ncu --set full` captures of the streamed kernel in both modes (summarised on
# the box; bench.py reads roofline.traffic from them), memcheck over the fixture tests.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02n_final_smoke.log 2>&1; tail -3 gpurun_out/r02n_final_smoke.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -rP 2>&1 | grep -E "margin|passed|failed|Error|error|assert|FAILED" > gpurun_out/r02n_final_gpu_tests.log
tail -2 gpurun_out/r02n_final_gpu_tests.log | cut -c1-300
python bench.py 2> gpurun_out/r02n_final_bench_n1.err | tail -1 > gpurun_out/r02n_final_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02n_final_bench_reference_arm.json
FSE_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02n_final_launches.csv \
    python bench.py --steps 1 --warmup 1 --timesteps 3 --no-cpu-baseline --no-eager-gpu-baseline --no-e2e --no-kernel-timing --no-alt-mode --no-campnet --no-train > /dev/null 2>&1
NB="--steps 1 --warmup 1 --timesteps 6 --no-vocoder --no-e2e --no-cpu-baseline --no-eager-gpu-baseline --no-alt-mode --no-kernel-timing --no-campnet --no-train"
for m in tc_tf32 tc_bf16; do
  FSE_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:denoiser_stream_kernel -s 5 -c 1 -f \
      -o gpurun_out/prof_r02n_stream_$m python bench.py --mode $m $NB > gpurun_out/r02n_ncu_stream_$m.log 2>&1
  python tools/ncu_traffic.py gpurun_out/prof_r02n_stream_$m.ncu-rep --md gpurun_out/r02n_ncu_stream_$m.md | tail -3
done
bash tools/sanitizer.sh memcheck 300 | tail -4
python - <<'PY'
import json
for f in ("gpurun_out/r02n_final_bench_n1.json", "gpurun_out/r02n_final_bench_reference_arm.json"):
    try:
        d = json.load(open(f))
        print(f, d.get("dtype"), d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("clocks"), d.get("gpu_launches"))
        for k in ("alt_mode", "campnet", "train_step"):
            r = d.get(k) or {}
            print("  ", k, r.get("value"), r.get("ms_per_step") or r.get("ms_per_forward"), r.get("error"))
        print("   breakdown", d.get("breakdown"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
