#!/bin/bash
# compute-sanitizer passes over the kernels' GPU tests (race detection / sanitizers).  Slow (10-50x): run on
# the small fixture tests only.  usage (on the GPU box, from the repo root):  bash tools/sanitizer.sh [memcheck|racecheck|synccheck|initcheck]
tool=${1:-memcheck}
mkdir -p gpurun_out
sel='fixture or integer_ops or conv_gemm_tc_vs_simt or two_tile'
timeout 1500 compute-sanitizer --tool "$tool" --target-processes all --error-exitcode 3 \
    python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_zz_cond_encoder.py tests/test_gpu_zz_campnet.py -m gpu -q -x -k "$sel" \
    > "gpurun_out/sanitizer_${tool}.log" 2>&1
echo "compute-sanitizer $tool rc=$?" | tee -a "gpurun_out/sanitizer_${tool}.log"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Race|hazard" "gpurun_out/sanitizer_${tool}.log" | tail -20
