"""First-contact / regression probe of fse_wgrad (csrc/wgrad.cu): per-shape error against float64 in both tensor-core modes, and timing at the
training shape.  usage: python tools/wgrad_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_editing_toolkit_b200 import train  # noqa: E402


def ref(P, Q, offs):
    P64, Q64 = P.double(), Q.double()
    T = P.shape[1]
    out = torch.zeros(P.shape[2], Q.shape[2], len(offs), dtype=torch.float64, device=P.device)
    for j, off in enumerate(offs):
        lo, hi = max(0, -off), min(T, T - off)
        if hi > lo:
            out[:, :, j] = torch.einsum("btm,btn->mn", P64[:, lo:hi], Q64[:, lo + off:hi + off])
    return out


def main():
    for mode in ("tc_bf16", "tc_tf32"):
        wg = train.WeightGradGemm(mode)
        dt = torch.bfloat16 if mode == "tc_bf16" else torch.float32
        gen = torch.Generator().manual_seed(3)

        def operand(*shape):
            x = torch.randn(*shape, generator=gen)
            if mode == "tc_tf32":
                x = (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
            return x.to(dt).cuda()
        cases = [("one chunk", operand(1, 64, 128), operand(1, 64, 64), (0,)),
                 ("M=128 N=256", operand(1, 64, 128), operand(1, 64, 256), (0,)),
                 ("K=4 chunks", operand(1, 256, 128), operand(1, 256, 256), (0,)),
                 ("B=3 ragged taps", operand(3, 203, 512), operand(3, 203, 256), (-1, 0, 1)),
                 ("narrow N", operand(1, 160, 256), operand(1, 160, 80), (0,)),
                 ("narrow M", operand(2, 96, 80), operand(2, 96, 256), (0,)),
                 ("flat rows", operand(1, 4099, 512), operand(1, 4099, 256), (0,))]
        for name, P, Q, offs in cases:
            out = torch.full((P.shape[2], Q.shape[2], len(offs)), float("nan"), device="cuda")
            wg(P, Q, out, offs)
            torch.cuda.synchronize()
            want = ref(P, Q, offs)
            d = (out.double() - want).abs()
            print(f"{mode} {name:18s} max|want| {float(want.abs().max()):9.3f} max err {float(d.nan_to_num(1e9).max()):.3e} nan {int(out.isnan().sum())} "
                  f"zeros {int((out == 0).sum())}/{out.numel()} out[0,:3,0] {out[0, :3, 0].tolist()} want {want[0, :3, 0].tolist()}")
        # timing at the training shape: conv taps of one layer
        B, T = 32, 1024
        P, Q = operand(B, T, 512), operand(B, T, 256)
        out = torch.empty(512, 256, 3, device="cuda")
        for _ in range(3):
            wg(P, Q, out, (-1, 0, 1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            wg(P, Q, out, (-1, 0, 1))
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"{mode} conv wgrad 32x1024, 512 x 256 x 3 taps: {us:.1f} us = {2 * B * T * 512 * 768 / us / 1e6:.0f} TFLOP/s")


if __name__ == "__main__":
    main()
