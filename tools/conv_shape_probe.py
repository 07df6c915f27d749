"""Kernel time of the conv-as-GEMM kernel on the HiFi-GAN resblock shapes, per schedule (CUDA events, warm)."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speech_editing_toolkit_b200 import _lib


def taps(k, d):
    return [(j - (k - 1) // 2) * d for j in range(k)]


def run(name, B, T, C0, offs, N, BN, KB, sched, MT, reps=5):
    nkb = (C0 + KB - 1) // KB
    Kp = len(offs) * nkb * KB
    A = torch.randn(B, T, C0, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, Kp, device="cuda") / 30).to(torch.bfloat16)
    out = torch.empty(B * T, N, device="cuda")
    arr = (C.c_int32 * len(offs))(*offs)
    ts = []
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib().fse_debug_conv_gemm(0, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(out.data_ptr()),
                                                  B, T, C0, len(offs), arr, N, BN, KB, C.c_void_p(torch.cuda.current_stream().cuda_stream),
                                                  None, sched + 16 * (MT - 1)))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = min(ts[1:])
    rows = B * T
    flops = 2.0 * rows * N * C0 * len(offs)
    byts = rows * (C0 * 2 + N * 4)
    print(f"{name:46s} sched={sched} MT={MT}: {t:8.1f} us  {flops / t / 1e6:7.1f} TFLOP/s  {byts / t / 1e3:6.0f} GB/s (algorithmic)  "
          f"{t * 1.9e3 / (rows / 128 / 148) / len(offs):6.0f} cyc per (tap, 128-row sub-tile) per SM")


if __name__ == "__main__":
    R = 1 << 21   # rows
    for k, d in ((3, 1), (11, 5)):
        for sched, MT in ((0, 1), (0, 8), (1, 1), (1, 8)):
            run(f"stage4 C=32 k={k} d={d}", 8, R // 8, 32, taps(k, d), 32, 32, 32, sched, MT)
    for sched, MT in ((0, 4), (1, 4), (1, 1)):
        run("stage3 C=64 k=11 d=5", 8, R // 8, 64, taps(11, 5), 64, 64, 64, sched, MT)
    for sched, MT in ((0, 2), (1, 2)):
        run("stage2 C=128 k=11 d=5", 8, R // 16, 128, taps(11, 5), 128, 128, 64, sched, MT)
    for sched, MT in ((0, 1), (1, 1)):
        run("stage1 C=256 k=11 d=5", 8, R // 64, 256, taps(11, 5), 256, 256, 64, sched, MT)
