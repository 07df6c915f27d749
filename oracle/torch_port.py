"""CPU port of the reference path in plain torch functional ops (fp32, torch's own oneDNN/ATen CPU
kernels — the same library calls the reference's CPU path ends in).  The functions are device-agnostic: on CUDA tensors the
same code is the reference's eager GPU path (cuDNN / cuBLAS, TF32 convolutions by default), which `bench.py
--eager-gpu-baseline` times as the "reference single-GPU PyTorch" figure of BASELINE.json's north_star.

TEST / BASELINE INFRASTRUCTURE — NOT PRODUCT CODE.  Used by bench.py's `cpu_baseline` leg and
`--impl reference` arm on the GPU box, where the reference tree itself cannot travel; validated
against the unmodified reference in tests/test_torch_port.py (build container).  Citations are
file:line of the reference tree.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def to_torch(sd) -> Dict[str, torch.Tensor]:
    return {k: torch.as_tensor(v).float() for k, v in sd.items()}


def sinusoidal_pos_emb(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffnet.py:34-46"""
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, device=t.device) * -emb)
    emb = t[:, None] * emb[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


@torch.no_grad()
def diffnet_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, t: torch.Tensor, cond: torch.Tensor,
                    dilation_cycle_length: int = 1) -> torch.Tensor:
    """DiffNet.forward (diffnet.py:110-132): x[B,M,T], t[B], cond[B,H,T] -> x0[B,M,T]."""
    C = p["input_projection.weight"].shape[0]
    n_layers = sum(1 for k in p if k.endswith("dilated_conv.weight"))
    h = F.relu(F.conv1d(x, p["input_projection.weight"], p["input_projection.bias"]))
    e = sinusoidal_pos_emb(t.float(), C)
    e = F.linear(e, p["mlp.0.weight"], p["mlp.0.bias"])
    e = e * torch.tanh(F.softplus(e))                                   # Mish (diffnet.py:14-16)
    temb = F.linear(e, p["mlp.2.weight"], p["mlp.2.bias"])
    skips = []
    for n in range(n_layers):
        pre = f"residual_layers.{n}."
        dil = 2 ** (n % dilation_cycle_length)
        d = F.linear(temb, p[pre + "diffusion_projection.weight"], p[pre + "diffusion_projection.bias"]).unsqueeze(-1)
        c = F.conv1d(cond, p[pre + "conditioner_projection.weight"], p[pre + "conditioner_projection.bias"])
        y = F.conv1d(h + d, p[pre + "dilated_conv.weight"], p[pre + "dilated_conv.bias"], padding=dil, dilation=dil) + c
        gate, filt = torch.chunk(y, 2, dim=1)
        y = torch.sigmoid(gate) * torch.tanh(filt)
        y = F.conv1d(y, p[pre + "output_projection.weight"], p[pre + "output_projection.bias"])
        res, skip = torch.chunk(y, 2, dim=1)
        h = (h + res) / math.sqrt(2.0)
        skips.append(skip)
    s = torch.sum(torch.stack(skips), dim=0) / math.sqrt(n_layers)      # diffnet.py:128 (incl. the stack)
    s = F.relu(F.conv1d(s, p["skip_projection.weight"], p["skip_projection.bias"]))
    return F.conv1d(s, p["output_projection.weight"], p["output_projection.bias"])


@torch.no_grad()
def sample_loop(p, sched: Dict[str, torch.Tensor], cond: torch.Tensor, timesteps: int, noise=None, steps=None):
    """GaussianDiffusion.forward infer branch (spec_denoiser.py:177-185) with p_sample (:103-108) and
    q_posterior_sample (:95-101).  `steps` bounds the number of iterations actually run (for timing)."""
    B, H, T = cond.shape
    M = p["output_projection.weight"].shape[0]
    dev = cond.device
    x = noise[0] if noise is not None else torch.randn(B, M, T, device=dev)
    it = list(reversed(range(timesteps)))
    if steps is not None:
        it = it[:steps]
    for k, i in enumerate(it):
        t = torch.full((B,), i, dtype=torch.long, device=dev)
        x0 = diffnet_forward(p, x, t, cond)
        c1 = sched["posterior_mean_coef1"][t][:, None, None]
        c2 = sched["posterior_mean_coef2"][t][:, None, None]
        lv = sched["posterior_log_variance_clipped"][t][:, None, None]
        z = noise[1 + k] if noise is not None else torch.randn_like(x)
        nz = (1 - (t == 0).float())[:, None, None]
        x = c1 * x0 + c2 * x + nz * (0.5 * lv).exp() * z
    return x.transpose(1, 2)


def fold_weight_norm(p: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in p.items():
        if k.endswith(".weight_v"):
            g = p[k[:-2] + "_g"]
            norm = v.flatten(1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
            out[k[:-9] + ".weight"] = v * (g / norm)
        elif not k.endswith(".weight_g"):
            out[k] = v
    return out


@torch.no_grad()
def hifigan_forward(p: Dict[str, torch.Tensor], cfg: dict, mel: torch.Tensor) -> torch.Tensor:
    """HifiGanGenerator.forward (hifigan.py:126-142); mel[B,80,T] -> wav[B,1,T*hop].  The reference keeps
    weight-norm at inference and re-normalises every call (vocoder_infer/hifigan.py:13-21), so the fold is
    inside the timed function here too."""
    w = fold_weight_norm(p)
    nk = len(cfg["resblock_kernel_sizes"])
    x = F.conv1d(mel, w["conv_pre.weight"], w["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, 0.1)
        x = F.conv_transpose1d(x, w[f"ups.{i}.weight"], w[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (rk, rd) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            pre = f"resblocks.{i * nk + j}."
            y = x
            for m, d in enumerate(rd):
                yt = F.leaky_relu(y, 0.1)
                yt = F.conv1d(yt, w[pre + f"convs1.{m}.weight"], w[pre + f"convs1.{m}.bias"], dilation=d, padding=(rk * d - d) // 2)
                yt = F.leaky_relu(yt, 0.1)
                yt = F.conv1d(yt, w[pre + f"convs2.{m}.weight"], w[pre + f"convs2.{m}.bias"], padding=(rk - 1) // 2)
                y = yt + y
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)
    x = F.conv1d(x, w["conv_post.weight"], w["conv_post.bias"], padding=3)
    return torch.tanh(x)
