#!/bin/bash
# Round 2, pass N: A/B of the streamed kernel's L2 eviction priorities (FSE_STREAM_L2HINT) and of its programmatic dependent launch
# with alternating publication counters (FSE_STREAM_PDL): timing by CUDA events over graph replays, DRAM bytes by ncu, results must
# be bit-identical; then the GPU suite with both switched on.  (As run; the L2-hint switch and its sweep tool were removed afterwards —
# no effect on time or DRAM bytes, profiles/r02n_* — and the dependent launch became the default; tools/sample_ab.py is the A/B tool now.)
mkdir -p gpurun_out
L=gpurun_out/r02n_l2hint_pdl.log
: > $L
timeout 300 python tools/l2hint_sweep.py tc_tf32 0,1,2,3,7,11,19,23,31 >> $L 2>&1
timeout 200 python tools/l2hint_sweep.py tc_bf16 0,3,7,19 >> $L 2>&1
FSE_STREAM_PDL=1 timeout 200 python tools/l2hint_sweep.py tc_tf32 0,3,7 >> $L 2>&1
FSE_STREAM_PDL=1 timeout 200 python tools/l2hint_sweep.py tc_bf16 0,3 >> $L 2>&1
cat $L
for m in 0 3 7 19; do
  FSE_GRAPH=0 timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
      -k regex:denoiser_stream_kernel -s 2 -c 2 --csv --log-file gpurun_out/r02n_ncu_dram_mask$m.csv python tools/l2hint_sweep.py tc_tf32 $m --ncu > /dev/null 2>&1
  tail -8 gpurun_out/r02n_ncu_dram_mask$m.csv | cut -c1-300
done
python tools/step_gaps.py tc_tf32 > gpurun_out/r02n_step_gaps.log 2>&1
FSE_STREAM_PDL=1 python tools/step_gaps.py tc_tf32 >> gpurun_out/r02n_step_gaps.log 2>&1
cat gpurun_out/r02n_step_gaps.log
FSE_STREAM_PDL=1 FSE_STREAM_L2HINT=3 timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -5 | tee gpurun_out/r02n_gpu_tests_pdl_hint.log
