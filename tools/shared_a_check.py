"""Which descriptor base-offset convention makes row-shifted (shared-A) descriptors read the right rows?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import test_gpu_conv_gemm as t

for case in t.SHARED_CASES:
    B, T, C0, offs, N, BN, KB = case
    A, W, ref = t._operands(B, T, C0, offs, N, KB)
    for mode in (1, 2):   # 1 = base-offset 0 (correct on B200), 2 = base-offset (addr>>7)&7 (wrong)
        try:
            out = t.run("tc_bf16", A, W, B, T, C0, offs, N, BN, KB, shared_a=mode)
            err = np.abs(out - ref)
            print(f"case C0={C0} offs={offs[:3]}.. KB={KB} N={N} shared_a={mode}: max err {np.nanmax(err):.3e} nan={np.isnan(out).sum()}")
        except Exception as e:
            print("case", case, "mode", mode, "FAILED", e)
            raise
