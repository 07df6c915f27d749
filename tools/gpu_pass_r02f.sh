#!/bin/bash
# Round 2, pass F: training tests (model-level training branch), ncu --set full captures of the four denoiser launches of one
# DiffNet evaluation in both tensor-core modes and of a stage-2 residual conv of the vocoder (kernel names matched demangled).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zzzz_train.py tests/test_gpu_zz_campnet.py -m gpu -q --timeout 300 -rP 2>&1 | grep -E "margin|passed|failed|Error|error|assert|FAILED" > gpurun_out/r02f_train_tests.log
tail -4 gpurun_out/r02f_train_tests.log | cut -c1-400
export FSE_GRAPH=0
NB="--steps 1 --warmup 1 --timesteps 6 --no-vocoder --no-e2e --no-cpu-baseline --no-eager-gpu-baseline --no-alt-mode --no-kernel-timing --no-campnet --no-train"
for m in tc_tf32 tc_bf16; do
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:denoiser_stream_kernel|EpiIn|EpiSkip|EpiOut" -s 24 -c 4 -f \
      -o gpurun_out/prof_r02f_denoiser_$m python bench.py --mode $m $NB > gpurun_out/r02f_ncu_denoiser_$m.log 2>&1
  tail -3 gpurun_out/r02f_ncu_denoiser_$m.log
done
timeout 400 ncu --set full --clock-control none --kernel-name-base demangled -k regex:EpiResAdd -s 39 -c 1 -f -o gpurun_out/prof_r02f_voc_resadd_tf32 \
    python tools/voc_launches.py tc_tf32 > gpurun_out/r02f_ncu_voc_resadd.log 2>&1
tail -3 gpurun_out/r02f_ncu_voc_resadd.log
ls -la gpurun_out/*.ncu-rep
