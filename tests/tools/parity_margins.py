"""Print the measured parity margins of the tensor-core path (relative mel-L1) against the reference fixtures and
the bf16-operand oracle — the numbers the tolerances in tests/test_gpu_parity.py are set from."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from conftest import golden, rel_l1
from oracle import fluentspeech_oracle as O
from speech_editing_toolkit_b200 import schedule, synth
from speech_editing_toolkit_b200.engine import Denoiser

sd = synth.denoiser_state_dict(1234)
g = golden("diffnet_step.npz")
B, T = int(g["B"]), int(g["T"])
cond = synth.synthetic_cond(int(g["seed"]), B, T)
d = Denoiser(mode="tc_bf16"); d.load_state_dict(sd)
x0 = d.denoise_step(torch.from_numpy(g["x"]).cuda(), torch.from_numpy(cond).cuda(), torch.from_numpy(g["t"]).cuda()).cpu().numpy()
ob = O.diffnet_forward(sd, g["x"], g["t"], cond.transpose(0, 2, 1), gemm_dtype="bf16")
print(f"denoise step : vs fp32 reference fixture {rel_l1(x0, g['x0']):.3e} (tol 2e-2) | vs bf16 oracle {rel_l1(x0, ob):.3e} (tol 8e-3)")
g = golden("sample_c1.npz")
seed, B, T, S = int(g["seed"]), int(g["B"]), int(g["T"]), int(g["S"])
b = schedule.diffusion_buffers(S)
d.set_schedule(b["posterior_mean_coef1"], b["posterior_mean_coef2"], b["posterior_log_variance_clipped"])
mel = d.sample(torch.from_numpy(synth.synthetic_cond(seed, B, T)).cuda(), torch.from_numpy(synth.synthetic_noise(seed, S, B, T)).cuda()).cpu().numpy()
print(f"sample C1    : vs fp32 reference fixture {rel_l1(mel, g['mel_out']):.3e} (tol 2e-2)")
