#!/bin/bash
# Round 2, pass E: GPU suite after the vectorised LayerNorm, then `ncu --set full` captures (one launch each, unthrottled clocks) of
# the current stream kernel in both tensor-core modes, the three non-stream denoiser launches, a fused conv pair and a residual conv
# of the vocoder, and the CampNet attention kernel.  Summaries: tools/ncu_traffic.py -> profiles/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -rP 2>&1 | grep -E "margin|passed|failed|Error|error|assert|FAILED" > gpurun_out/r02e_gpu_tests.log
tail -3 gpurun_out/r02e_gpu_tests.log
python tools/campnet_bench.py > gpurun_out/r02e_campnet_bench.json 2> gpurun_out/r02e_campnet_bench.err; tail -2 gpurun_out/r02e_campnet_bench.json | cut -c1-600
export FSE_GRAPH=0
NB="--steps 1 --warmup 1 --timesteps 6 --no-vocoder --no-e2e --no-cpu-baseline --no-eager-gpu-baseline --no-alt-mode --no-kernel-timing --no-campnet --no-train"
for m in tc_tf32 tc_bf16; do
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:denoiser_stream_kernel|EpiIn|EpiSkip|EpiOut" -s 24 -c 4 -f \
      -o gpurun_out/prof_r02e_denoiser_$m python bench.py --mode $m $NB > /dev/null 2> gpurun_out/r02e_ncu_denoiser_$m.err
done
# vocoder, tf32: fused pairs (C = 32) of the second forward: k = 3 (first), k = 11 (7th); residual convs: stage 2 k = 7 (13th of 27)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:resblock_pair_kernel -s 9 -c 7 -f -o gpurun_out/prof_r02e_voc_pair_tf32 \
    python tools/voc_launches.py tc_tf32 > /dev/null 2> gpurun_out/r02e_ncu_voc_pair.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:EpiResAdd -s 39 -c 1 -f -o gpurun_out/prof_r02e_voc_resadd_tf32 \
    python tools/voc_launches.py tc_tf32 > /dev/null 2> gpurun_out/r02e_ncu_voc_resadd.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:camp_attention_tc2_kernel -s 14 -c 1 -f -o gpurun_out/prof_r02e_camp_attention_tc2 \
    python tools/campnet_bench.py --iters 1 --warmup 1 > /dev/null 2> gpurun_out/r02e_ncu_camp.err
ls -la gpurun_out/*.ncu-rep
