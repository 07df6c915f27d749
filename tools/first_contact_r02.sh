#!/bin/bash
# What round 1 left unmeasured (its GPU budget ended first).  One GPU-box pass, outputs under gpurun_out/:
#   usage (from the repo root): gpurun --timeout 900 -- 'bash tools/first_contact_r02.sh'
mkdir -p gpurun_out
# 1. first GPU contact of the region-surgery kernels, SpecDenoiserInferB200.forward_model(sample) and the mel front-end
FSE_TEST_UNVERIFIED=1 timeout 200 python -m pytest tests/test_gpu_zz_cond_encoder.py -m gpu -q --timeout 150 \
    -k "region_surgery or edit_forward or mel_frontend" 2>&1 | tail -30 | tee gpurun_out/r02_first_contact.log
# 2. the whole GPU suite + smoke (default attention = two-tile tcgen05 schedule)
timeout 200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -3 | tee gpurun_out/r02_gpu_tests.log
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/r02_smoke.log
# 3. bench line with the eager-PyTorch GPU baseline (north_star's "reference single-GPU PyTorch" figure)
timeout 300 python bench.py --steps 3 --warmup 3 --eager-gpu-baseline 2>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_n1_eager.json
# 4. ncu of the two-tile attention kernel + CampNet launch list
timeout 120 ncu --set full --clock-control none --import-source on -k regex:camp_attention_tc2_kernel -s 3 -c 1 -f -o gpurun_out/prof_r02_attention_tc2 \
    python tools/campnet_bench.py --iters 1 --warmup 1 > /dev/null 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_campnet_launches_tc2.csv \
    python tools/campnet_bench.py --iters 1 --warmup 1 > /dev/null 2>&1
# 5. sanitizers over the small tests
bash tools/sanitizer.sh memcheck
bash tools/sanitizer.sh racecheck
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_n1_eager.json"))
    print("bench:", d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("eager_gpu_baseline"))
except Exception as e:
    print("bench unreadable:", e)
PY
