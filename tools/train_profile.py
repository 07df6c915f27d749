"""Per-kernel device time of one training step (train.train_step, BASELINE configs[4] shape) from torch.profiler — where the step's
time goes between the native data-path kernels (fse::*), the library weight-gradient GEMMs and the torch glue.
usage: python tools/train_profile.py [--mode tc_bf16] [--B 32] [--T 1024] [--top 40]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_editing_toolkit_b200 import schedule, synth, train  # noqa: E402
from speech_editing_toolkit_b200.modules import DiffNetB200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="tc_bf16")
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--T", type=int, default=1024)
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B, T = args.B, args.T
    hp = dict(audio_num_mel_bins=80, hidden_size=192, residual_layers=20, residual_channels=256, dilation_cycle_length=1, b200_mode=args.mode)
    net = DiffNetB200(80, hp).to(dev).train()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.denoiser_state_dict(1234).items()})
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, betas=(0.9, 0.98), weight_decay=0.0, fused=True)
    sched = {k: torch.from_numpy(v).to(dev) for k, v in schedule.diffusion_buffers(100).items()}
    batch = synth.synthetic_edit_batch(1234, B, T)
    data = {"ref_mels": torch.from_numpy(batch["ref_mels"]).to(dev), "time_mel_masks": torch.from_numpy(batch["time_mel_masks"]).to(dev),
            "cond": torch.from_numpy(synth.synthetic_cond(1234, B, T)).to(dev)}
    for _ in range(3):
        train.train_step(net, sched, data, opt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        train.train_step(net, sched, data, opt)
    e1.record()
    torch.cuda.synchronize()
    print(f"train_step {args.mode} {B} x {T}: {e0.elapsed_time(e1) / 5:.2f} ms / step (CUDA events, 5 steps)")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        train.train_step(net, sched, data, opt)
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    print(f"kernels: {len(rows)} kinds, {sum(r[2] for r in rows)} launches, {tot / 1e3:.2f} ms of device time")
    for k, us, n in rows[:args.top]:
        print(f"{us / 1e3:9.3f} ms {100 * us / tot:5.1f} % x{n:<4d} {k[:150]}")


if __name__ == "__main__":
    main()
