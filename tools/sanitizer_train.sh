#!/bin/bash
# compute-sanitizer over the training kernels' GPU tests (fse_wgrad*, fse_mel_loss_*, the native DiffNet chain on the small fixture).
# usage (GPU box, repo root): bash tools/sanitizer_train.sh [memcheck|synccheck|racecheck] [seconds]
tool=${1:-memcheck}
limit=${2:-400}
mkdir -p gpurun_out
sel='weight_gradient_gemm or mel_loss_kernels_vs_reference_fixture or edge_shapes or (gradients_vs_reference and tc_bf16)'
timeout "$limit" compute-sanitizer --tool "$tool" --target-processes all --error-exitcode 3 \
    python -m pytest tests/test_gpu_zzzz_train.py -m gpu -q -x -k "$sel" -p no:cacheprovider \
    > "gpurun_out/r02m_sanitizer_train_${tool}.log" 2>&1
echo "compute-sanitizer $tool rc=$? (124 = time limit)" | tee -a "gpurun_out/r02m_sanitizer_train_${tool}.log"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Race|hazard|Error" "gpurun_out/r02m_sanitizer_train_${tool}.log" | sort | uniq -c | sort -rn | head -12
