"""Phase timing of one conv_gemm tile on the GPU (clock64 stamps of CTA 0) for the denoiser GEMM shapes."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from speech_editing_toolkit_b200 import _lib

def run(name, B, T, C0, offs, N, BN, KB=64, reps=3, shared_a=0):
    nkb = (C0 + KB - 1) // KB
    Kp = len(offs) * nkb * KB
    A = torch.randn(B, T, C0, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, Kp, device="cuda") / 30).to(torch.bfloat16)
    out = torch.empty(B * T, N, device="cuda")
    arr = (C.c_int32 * len(offs))(*offs)
    for r in range(reps):
        dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib().fse_debug_conv_gemm(0, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(out.data_ptr()),
                                                  B, T, C0, len(offs), arr, N, BN, KB, C.c_void_p(torch.cuda.current_stream().cuda_stream),
                                                  C.c_void_p(dbg.data_ptr()), shared_a))
        e1.record()
        torch.cuda.synchronize()
        d = dbg.cpu().numpy()
    t0 = d[0]
    tiles = B * ((T + 127) // 128) * (N // BN)
    print(f"{name}: tiles={tiles} kernel={e0.elapsed_time(e1)*1e3:.1f}us  cycles from start: setup={d[1]-t0} loads_issued={d[2]-t0} first_kb_landed={d[3]-t0} "
          f"mma_issued={d[4]-t0} acc_ready={d[5]-t0} epi_done={[int(x-t0) for x in d[6:14]]} end={d[14]-t0}")
    print("   warp2 chunks (start, after tmem ld, after transpose, after apply):", [[int(x - t0) for x in d[16 + 4 * i:20 + 4 * i]] for i in range(4)])

if __name__ == "__main__":
    run("gate 1 tile/CTA (64 tiles)", 4, 1024, 256, [-1, 0, 1, 0], 512, 256)     # K=1024 ~ 960 of the real gate GEMM
    run("gate full (512 tiles)", 32, 1024, 256, [-1, 0, 1, 0], 512, 256)
    run("res 1 tile/CTA", 4, 1024, 256, [0], 512, 256)
    run("res full", 32, 1024, 256, [0], 512, 256)
    run("gate shared-A full", 32, 1024, 256, [-1, 0, 1], 512, 256, shared_a=1)
    run("gate separate-A full (3 taps)", 32, 1024, 256, [-1, 0, 1], 512, 256, shared_a=0)
