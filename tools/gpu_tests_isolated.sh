#!/bin/bash
# Run every GPU test in its own process (a device-side trap poisons the CUDA context of the process),
# logging to gpurun_out/.  Usage (on the GPU box): bash tools/gpu_tests_isolated.sh [pytest file ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
files=("$@"); [ ${#files[@]} -eq 0 ] && files=(tests/test_gpu_conv_gemm.py tests/test_gpu_parity.py)
nvidia-smi -L | tee gpurun_out/isolated.log
python -m pytest "${files[@]}" -m gpu --collect-only -q 2>/dev/null | grep "::" > gpurun_out/ids.txt
pass=0; fail=0
while read -r id; do
  out=$(timeout 600 python -m pytest "$id" -q -x --no-header -p no:cacheprovider 2>&1)
  rc=$?
  if [ $rc -eq 0 ]; then pass=$((pass+1)); echo "PASS $id" | tee -a gpurun_out/isolated.log
  else fail=$((fail+1)); echo "FAIL($rc) $id" | tee -a gpurun_out/isolated.log; echo "$out" | tail -25 | tee -a gpurun_out/isolated.log; fi
done < gpurun_out/ids.txt
echo "passed=$pass failed=$fail" | tee -a gpurun_out/isolated.log
