#!/bin/bash
# compute-sanitizer passes over the kernels' GPU tests (race detection / sanitizers).  Slow (10-50x): run on
# the small fixture tests only.  usage (on the GPU box, from the repo root):  bash tools/sanitizer.sh [memcheck|racecheck|synccheck|initcheck] [seconds]
tool=${1:-memcheck}
limit=${2:-600}
mkdir -p gpurun_out
sel='denoise_step_vs_reference_fixture or hifigan_vs_reference_fixture or mel_encoder_vs_reference or test_tc_matches_simt_bf16_on_ragged_shapes'
timeout "$limit" compute-sanitizer --tool "$tool" --target-processes all --error-exitcode 3 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$sel" -p no:cacheprovider \
    > "gpurun_out/sanitizer_${tool}.log" 2>&1
echo "compute-sanitizer $tool rc=$? (124 = time limit)" | tee -a "gpurun_out/sanitizer_${tool}.log"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Race|hazard|Error" "gpurun_out/sanitizer_${tool}.log" | sort | uniq -c | sort -rn | head -12
