"""CampNet mask-predict forward at BASELINE.json configs[3] (batch 64, egs/campnet.yaml shapes) on one B200: CUDA-event timing of
CampNetB200.forward with inputs resident, the attention kernel's share from per-kernel events is read from the ncu launch list
(profiles/).  Usage: python tools/campnet_bench.py [--batch 64 --frames 1024 --mode tc_bf16 --iters 5]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speech_editing_toolkit_b200 import synth                      # noqa: E402
from speech_editing_toolkit_b200.modules import CampNetB200        # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--frames", type=int, default=1024)
ap.add_argument("--mode", default="tc_bf16")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
B, T = a.batch, a.frames
net = CampNetB200(80, 100, dict(hidden_size=192, dec_ffn_kernel_size=9, audio_num_mel_bins=80, b200_mode=a.mode)).cuda()
net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.campnet_state_dict(1234, 80).items()}, strict=False)
b = synth.synthetic_campnet_batch(1, B, T, vocab=80)
txt, mels, m = (torch.from_numpy(b[k]).cuda() for k in ("txt_tokens", "mels", "time_mel_masks"))
for _ in range(a.warmup):
    out = net(txt, mels=mels, time_mel_masks=m, infer=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    out = net(txt, mels=mels, time_mel_masks=m, infer=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
Tt, H = txt.shape[1], 192
# algorithmic FLOPs of one forward (2 x MACs): projections + FFN + ConvBlocks + MelEncoder per frame / token, attention per item
enc_tok = 3 * (2 * H * 3 * H + 2 * H * H + 2 * H * 4 * H * 9 + 2 * 4 * H * H)
dec_frm = 6 * (2 * H * 3 * H + 2 * H * H + 2 * H * H + 2 * H * H + 2 * H * 4 * H * 9 + 2 * 4 * H * H)
fine_frm = 10 * (2 * H * 2 * H * 5 + 2 * 2 * H * H) + 2 * H * H * 3
mel_enc = 2 * (2 * 80 * H + 2 * 2 * H * H) + 2 * 2 * H * 80
gemm = B * Tt * (enc_tok + 6 * 2 * H * 2 * H) + B * T * (dec_frm + fine_frm + mel_enc)
attn = B * (3 * 4 * Tt * Tt * H + 6 * 4 * T * T * H + 6 * 4 * T * Tt * H)
print(json.dumps({"metric": "CampNet mask-predict forward", "mode": a.mode, "batch": B, "frames": T, "tokens": Tt, "ms_per_forward": ms,
                  "mel_frames_per_s": B * T / (ms / 1e3), "launches": net.engine().last_launches,
                  "gemm_tflop": gemm / 1e12, "attention_tflop": attn / 1e12, "achieved_tflops": (gemm + attn) / (ms / 1e3) / 1e12,
                  "finite": bool(torch.isfinite(out["mel_out_fine"]).all().item())}))
