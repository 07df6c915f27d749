"""CPU restatement (torch, fp32) of the native training path of csrc/denoiser_train.cuh: the forward that keeps what the backward
needs and the activation-gradient chain, buffer for buffer, behind the interface of speech_editing_toolkit_b200.train.DiffNetTrainer.

TEST INFRASTRUCTURE ONLY.  tests/test_train_oracle.py plugs it under train.DiffNetFunction (which forms the weight gradients from the
buffers) and compares every gradient with torch.autograd through the reference DiffNet / its fixture
(tests/golden/diffnet_train.npz) — so the algebra of the native design and the caller-side weight-gradient code are pinned on the
CPU, and the GPU test only has to show that the kernels compute these same buffers.  Reference: diffnet.py:60-132.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F


class TorchTrainer:
    def __init__(self, n_mels=80, hidden=192, channels=256, layers=20, dilation_cycle_length=1, mode="simt_f32"):
        self.cfg = SimpleNamespace(n_mels=n_mels, hidden=hidden, channels=channels, layers=layers, dilation_cycle_length=dilation_cycle_length)
        self.mode, self.es, self.op_dtype = mode, 4, torch.float32
        self.p, self.buf = None, None

    def load(self, named):
        self.p = {k: v.detach().float() for k, v in named.items()}

    def forward(self, x_t, cond_bth, d):
        p, L, C = self.p, self.cfg.layers, self.cfg.channels
        B, M, T = x_t.shape
        N = B * T
        x_rows = x_t.transpose(1, 2).reshape(N, M)
        h = F.relu(x_rows @ p["input_projection.weight"][:, :, 0].t() + p["input_projection.bias"])
        h0 = h.clone()
        S = torch.zeros(N, C)
        hin_all, sg_all, tf_all, u_all = [], [], [], []
        cond = cond_bth.reshape(N, -1)
        for l in range(L):
            pre = f"residual_layers.{l}."
            dil = 2 ** (l % self.cfg.dilation_cycle_length)
            hin = (h.view(B, T, C) + d[l][:, None, :])
            wdc = p[pre + "dilated_conv.weight"]
            y = torch.zeros(B, T, 2 * C)
            for j, off in enumerate((-dil, 0, dil)):
                lo, hi = max(0, -off), min(T, T - off)
                y[:, lo:hi] += hin[:, lo + off:hi + off] @ wdc[:, :, j].t()
            y = y.reshape(N, 2 * C) + cond @ p[pre + "conditioner_projection.weight"][:, :, 0].t() + p[pre + "dilated_conv.bias"] + p[pre + "conditioner_projection.bias"]
            sg, tf = torch.sigmoid(y[:, :C]), torch.tanh(y[:, C:])
            u = sg * tf
            o = u @ p[pre + "output_projection.weight"][:, :, 0].t() + p[pre + "output_projection.bias"]
            h = (h + o[:, :C]) / math.sqrt(2.0)
            S = S + o[:, C:]
            hin_all.append(hin); sg_all.append(sg); tf_all.append(tf); u_all.append(u)
        s = S / math.sqrt(L)
        r = F.relu(s @ p["skip_projection.weight"][:, :, 0].t() + p["skip_projection.bias"])
        x0 = r @ p["output_projection.weight"][:, :, 0].t() + p["output_projection.bias"]
        self.buf = dict(x_rows=x_rows, h0=h0, hin=torch.stack(hin_all), sg=sg_all, tf=tf_all, u=torch.stack(u_all), s=s, r=r, cond=cond_bth, B=B, T=T)
        return x0.view(B, T, M).transpose(1, 2).contiguous()

    def backward(self, dx0):
        p, L, C, b = self.p, self.cfg.layers, self.cfg.channels, self.buf
        B, T = b["B"], b["T"]
        N = B * T
        dx_rows = dx0.transpose(1, 2).reshape(N, -1)
        dz = (dx_rows @ p["output_projection.weight"][:, :, 0]) * (b["r"] > 0)
        dS = (dz @ p["skip_projection.weight"][:, :, 0]) / math.sqrt(L)
        dh = torch.zeros(N, C)
        dres = [None] * L
        dres[L - 1] = torch.zeros(N, C)
        dy_all = torch.zeros(B, T, L * 2 * C)
        for l in range(L - 1, -1, -1):
            pre = f"residual_layers.{l}."
            dil = 2 ** (l % self.cfg.dilation_cycle_length)
            du = torch.cat([dres[l], dS], 1) @ p[pre + "output_projection.weight"][:, :, 0]
            sg, tf = b["sg"][l], b["tf"][l]
            dy = torch.cat([du * tf * sg * (1 - sg), du * sg * (1 - tf * tf)], 1).view(B, T, 2 * C)
            dy_all[:, :, l * 2 * C:(l + 1) * 2 * C] = dy
            wdc = p[pre + "dilated_conv.weight"]
            dhin = torch.zeros(B, T, C)
            for j, off in enumerate((-dil, 0, dil)):          # forward: y[t] += W_j hin[t + off]  =>  dhin[t + off] += W_j^T dy[t]
                lo, hi = max(0, -off), min(T, T - off)
                dhin[:, lo + off:hi + off] += dy[:, lo:hi] @ wdc[:, :, j]
            dh = dh / math.sqrt(2.0) + dhin.reshape(N, C)
            if l > 0:
                dres[l - 1] = dh / math.sqrt(2.0)
        dcond = torch.zeros(B, T, self.cfg.hidden)
        for l in range(L):
            dcond += dy_all[:, :, l * 2 * C:(l + 1) * 2 * C] @ p[f"residual_layers.{l}.conditioner_projection.weight"][:, :, 0]
        b.update(dx_rows=dx_rows, dz=dz, dS=dS, dh0=dh, dres=torch.stack(dres), dy=dy_all)
        return dcond

    def views(self):
        b = self.buf
        return {k: b[k] for k in ("x_rows", "h0", "hin", "u", "s", "r", "cond", "dx_rows", "dz", "dS", "dh0", "dres", "dy")}


def mel_losses_torch(mel_out: torch.Tensor, target: torch.Tensor, lambdas=(("l1", 0.5), ("ssim", 0.5))):
    """Torch restatement of SpeechBaseTask.add_mel_loss with `mel_losses: l1:0.5|ssim:0.5` (tasks/tts/speech_base.py:219-257: l1_loss,
    ssim_loss, weights_nonzero_speech) over utils/metrics/ssim.py:12-44 (window, _ssim); its autograd gradient is the expected value
    of fse_mel_loss_backward.  Pinned to the reference's own methods by tests/test_train_oracle.py."""
    weights = (target.abs().sum(-1, keepdim=True) > 0).to(mel_out.dtype).repeat(1, 1, target.shape[-1])
    out = {}
    for name, lam in lambdas:
        if name == "l1":
            out["l1"] = (F.l1_loss(mel_out, target, reduction="none") * weights).sum() / weights.sum() * lam
        elif name == "ssim":
            a, b = mel_out[:, None] + 6.0, target[:, None] + 6.0
            k = torch.tensor([math.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)], device=a.device)
            k = (k / k.sum())[:, None]
            win = (k @ k.t())[None, None].to(a.dtype)
            mu1, mu2 = F.conv2d(a, win, padding=5), F.conv2d(b, win, padding=5)
            s1 = F.conv2d(a * a, win, padding=5) - mu1 * mu1
            s2 = F.conv2d(b * b, win, padding=5) - mu2 * mu2
            s12 = F.conv2d(a * b, win, padding=5) - mu1 * mu2
            ssim_map = ((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s1 + s2 + 9e-4))
            out["ssim"] = ((1 - ssim_map.mean(1)) * weights).sum() / weights.sum() * lam
    return out


def mel_loss_native_algebra(mel_out, target, lam_l1=0.5, lam_ssim=0.5):
    """numpy (float64) restatement of csrc/mel_loss.cu, field for field: the forward's three derivative fields and the backward's
    window pass — pins the kernel's closed-form gradient to the reference's autograd gradient on the CPU (tests/test_train_oracle.py).
    Returns (lam_l1 * l1, lam_ssim * ssim, d(sum of both) / d mel_out)."""
    import numpy as np
    from scipy.signal import correlate2d
    x, y = np.asarray(mel_out, np.float64), np.asarray(target, np.float64)
    B, T, M = x.shape
    g = np.array([math.exp(-(i - 5) ** 2 / (2 * 1.5 ** 2)) for i in range(11)])
    g = g / g.sum()
    W = np.outer(g, g)
    w = (np.abs(y).sum(-1) != 0).astype(np.float64)                       # [B, T]
    wsum = w.sum() * M
    C1, C2 = 1e-4, 9e-4
    conv = lambda img: correlate2d(img, W, mode="same", boundary="fill", fillvalue=0.0)
    l1 = (np.abs(x - y) * w[:, :, None]).sum() / wsum
    ssim_sum, grad = 0.0, np.zeros_like(x)
    for i in range(B):
        a, b = x[i] + 6.0, y[i] + 6.0
        mu1, mu2, eaa, ebb, eab = conv(a), conv(b), conv(a * a), conv(b * b), conv(a * b)
        A1, A2 = 2 * mu1 * mu2 + C1, 2 * (eab - mu1 * mu2) + C2
        B1, B2 = mu1 ** 2 + mu2 ** 2 + C1, (eaa - mu1 ** 2) + (ebb - mu2 ** 2) + C2
        S = A1 * A2 / (B1 * B2)
        ssim_sum += ((1 - S) * w[i][:, None]).sum()
        up = -w[i][:, None] / wsum
        Gmu = up * (2 * mu2 * (A2 - A1) / (B1 * B2) - 2 * mu1 * S * (1 / B1 - 1 / B2))
        Gaa = up * (-S / B2)
        Gab = up * (2 * A1 / (B1 * B2))
        grad[i] = lam_ssim * (conv(Gmu) + 2 * a * conv(Gaa) + b * conv(Gab)) + lam_l1 * np.sign(x[i] - y[i]) * w[i][:, None] / wsum
    return lam_l1 * l1, lam_ssim * ssim_sum / wsum, grad
