#!/bin/bash
# ncu --set full (with source) of the two backward kernels of the training chain, one launch each; the .ncu-rep files travel back.
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:EpiGateBwd|EpiDh" -s 44 -c 2 -f -o gpurun_out/prof_r02l_train_bwd python tools/train_profile.py --top 1 > gpurun_out/r02l_ncu_train.log 2>&1
python tools/ncu_traffic.py gpurun_out/prof_r02l_train_bwd.ncu-rep --md gpurun_out/r02l_ncu_train_backward_kernels.md > /dev/null
ls -la gpurun_out/*.ncu-rep
