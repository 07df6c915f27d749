#!/bin/bash
# Round 2, pass G: GPU suite with 192-wide GEMM tiles, CampNet timing, then ncu --set full captures summarised ON the box
# (tools/ncu_traffic.py -> markdown + raw csv; only the stream kernel's .ncu-rep files travel back: gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -rP 2>&1 | grep -E "margin|passed|failed|Error|error|assert|FAILED" > gpurun_out/r02g_gpu_tests.log
tail -3 gpurun_out/r02g_gpu_tests.log | cut -c1-300
python tools/campnet_bench.py > gpurun_out/r02g_campnet_bench.json 2> gpurun_out/r02g_campnet_bench.err; tail -1 gpurun_out/r02g_campnet_bench.json | cut -c1-400
FSE_BN192=0 python tools/campnet_bench.py 2>/dev/null | tail -1 | cut -c1-200
export FSE_GRAPH=0
NB="--steps 1 --warmup 1 --timesteps 6 --no-vocoder --no-e2e --no-cpu-baseline --no-eager-gpu-baseline --no-alt-mode --no-kernel-timing --no-campnet --no-train"
for m in tc_tf32 tc_bf16; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:denoiser_stream_kernel -s 5 -c 1 -f \
      -o gpurun_out/prof_r02g_stream_$m python bench.py --mode $m $NB > gpurun_out/r02g_ncu_stream_$m.log 2>&1
  timeout 400 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:EpiIn|EpiSkip|EpiOut" -s 18 -c 3 -f \
      -o /tmp/prof_r02g_proj_$m python bench.py --mode $m $NB > gpurun_out/r02g_ncu_proj_$m.log 2>&1
  python tools/ncu_traffic.py /tmp/prof_r02g_proj_$m.ncu-rep --md gpurun_out/r02g_ncu_denoiser_projections_$m.md > /dev/null
  ncu -i /tmp/prof_r02g_proj_$m.ncu-rep --page raw --csv > gpurun_out/r02g_ncu_denoiser_projections_$m.raw.csv 2>/dev/null
done
timeout 400 ncu --set full --clock-control none --kernel-name-base demangled -k regex:EpiResAdd -s 39 -c 1 -f -o /tmp/prof_r02g_voc_resadd_tf32 \
    python tools/voc_launches.py tc_tf32 > gpurun_out/r02g_ncu_voc_resadd.log 2>&1
python tools/ncu_traffic.py /tmp/prof_r02g_voc_resadd_tf32.ncu-rep --md gpurun_out/r02g_ncu_voc_resadd_stage2_tf32.md > /dev/null
ncu -i /tmp/prof_r02g_voc_resadd_tf32.ncu-rep --page raw --csv > gpurun_out/r02g_ncu_voc_resadd_stage2_tf32.raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -20; du -sh gpurun_out
