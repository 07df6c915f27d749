"""CPU oracle for the FluentSpeech spec_denoiser sampling path + HiFi-GAN generator forward.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs may import this module.  The shipped
path (speech_editing_toolkit_b200/) never imports it and fails loudly without its CUDA library.

This is a plain numpy restatement of the reference algorithm (reference = Zain-Jiang/
Speech-Editing-Toolkit @ a8d5bf33, all file:line citations relative to that tree).  The
reference ships NO tests / golden vectors for this path, so the oracle is
pinned against the reference ITSELF: oracle/make_golden.py imports the unmodified reference
modules in the build container, runs them on seeded inputs and writes tests/golden/*.npz;
tests/test_oracle_golden.py checks this file against those fixtures (and, when
/root/reference is present, against the live reference).  Parity status: PINNED to the
reference's own outputs (generated fixtures), not to upstream golden vectors (none exist).

Arithmetic: float32 everywhere (schedule in float64 then cast, as the reference does).
`gemm_dtype="bf16"` additionally rounds the OPERANDS of every tensor-core contraction to
bfloat16 (round-to-nearest-even, fp32 accumulation) — it restates the arithmetic contract of
the sm_100a kernels so device results can be checked tightly; "f32" is the reference contract.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round float32 -> bfloat16 (RNE) -> float32, matching __float2bfloat16_rn."""
    x = np.ascontiguousarray(x, dtype=F32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = rounded.astype(np.uint32).view(F32)
    return np.where(np.isnan(x), x, out).astype(F32)


def _q(x, gemm_dtype):
    return bf16_round(x) if gemm_dtype == "bf16" else x


def sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(F32)


def mish(x):
    """diffnet.py:14-16  x * tanh(softplus(x))."""
    x64 = x.astype(np.float64)
    return (x64 * np.tanh(np.logaddexp(0.0, x64))).astype(F32)


def leaky_relu(x, slope):
    return np.where(x >= 0, x, x * F32(slope)).astype(F32)


def conv1d(x, w, b, dilation=1, padding=0, gemm_dtype="f32"):
    """torch.nn.Conv1d (stride 1).  x[B,Ci,T], w[Co,Ci,K], b[Co] -> [B,Co,T']."""
    B, Ci, T = x.shape
    Co, _, K = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (padding, padding)))
    To = T + 2 * padding - dilation * (K - 1)
    xq = _q(xp, gemm_dtype)
    wq = _q(w, gemm_dtype)
    out = np.zeros((B, Co, To), dtype=F32)
    for k in range(K):
        seg = xq[:, :, k * dilation:k * dilation + To]          # [B,Ci,To]
        out += np.einsum("oc,bct->bot", wq[:, :, k], seg, optimize=True).astype(F32)
    if b is not None:
        out += b[None, :, None]
    return out.astype(F32)


def conv_transpose1d(x, w, b, stride, padding, gemm_dtype="f32"):
    """torch.nn.ConvTranspose1d.  x[B,Ci,T], w[Ci,Co,K] -> [B,Co,(T-1)*stride-2*padding+K]."""
    B, Ci, T = x.shape
    _, Co, K = w.shape
    Lfull = (T - 1) * stride + K
    xq = _q(x, gemm_dtype)
    wq = _q(w, gemm_dtype)
    full = np.zeros((B, Co, Lfull), dtype=F32)
    for k in range(K):
        contrib = np.einsum("co,bct->bot", wq[:, :, k], xq, optimize=True).astype(F32)   # [B,Co,T]
        full[:, :, k:k + (T - 1) * stride + 1:stride] += contrib
    out = full[:, :, padding:Lfull - padding]
    if b is not None:
        out = out + b[None, :, None]
    return out.astype(F32)


def linear(x, w, b, gemm_dtype="f32"):
    y = _q(x, gemm_dtype) @ _q(w, gemm_dtype).T
    return (y + b).astype(F32)


# --------------------------------------------------------------------------------------
# noise schedule + posterior buffers
# --------------------------------------------------------------------------------------
def vpsde_beta_t(t, T, min_beta, max_beta):
    """diffusion_utils.py:16-18."""
    t_coef = (2 * t - 1) / (T ** 2)
    return 1.0 - np.exp(-min_beta / T - 0.5 * (max_beta - min_beta) * t_coef)


def make_schedule(timesteps: int, min_beta: float = 0.1, max_beta: float = 40.0) -> Dict[str, np.ndarray]:
    """GaussianDiffusion.__init__ buffers (spec_denoiser.py:26-69) for schedule_type 'vpsde'
    (diffusion_utils.py:36-38): S+1 betas, float64 math, cast to float32 at the end."""
    n = timesteps + 1
    betas = np.array([vpsde_beta_t(t, n, min_beta, max_beta) for t in range(1, n + 1)], dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    sched = {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": np.log(np.maximum(post_var, 1e-20)),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }
    return {k: v.astype(F32) for k, v in sched.items()}


# --------------------------------------------------------------------------------------
# DiffNet denoiser (diffnet.py:34-132)
# --------------------------------------------------------------------------------------
def sinusoidal_pos_emb(t: np.ndarray, dim: int) -> np.ndarray:
    """diffnet.py:34-46: cat(sin(t*f), cos(t*f)), f_j = exp(-j ln(1e4)/(dim/2-1)).
    torch computes exp/sin/cos in float32; so do we."""
    half = dim // 2
    c = F32(math.log(10000) / (half - 1))
    f = np.exp(np.arange(half, dtype=F32) * -c).astype(F32)
    e = t.astype(F32)[:, None] * f[None, :]
    return np.concatenate([np.sin(e), np.cos(e)], axis=-1).astype(F32)


def diffnet_forward(p: Dict[str, np.ndarray], x: np.ndarray, t: np.ndarray, cond: np.ndarray,
                    dilation_cycle_length: int = 1, gemm_dtype: str = "f32",
                    return_layers: bool = False):
    """DiffNet.forward (diffnet.py:110-132) with ResidualBlock.forward (:68-81).

    p: state_dict of `denoise_fn` (keys as in the reference).  x[B,80,T] (the reference's
    spec[:,0]), t[B] int, cond[B,H,T].  Returns x0[B,80,T].

    gemm_dtype="bf16": operands of the big contractions (input/conditioner/dilated/output/
    skip projections) are rounded to bf16; the timestep-embedding path (mlp,
    diffusion_projection) stays fp32, exactly like the device kernels."""
    C = p["input_projection.weight"].shape[0]
    n_layers = sum(1 for k in p if k.endswith("dilated_conv.weight"))
    h = conv1d(x, p["input_projection.weight"], p["input_projection.bias"], gemm_dtype=gemm_dtype)
    h = np.maximum(h, 0).astype(F32)                                            # :118-120
    e = sinusoidal_pos_emb(t, C)                                                # :121
    temb = linear(mish(linear(e, p["mlp.0.weight"], p["mlp.0.bias"])), p["mlp.2.weight"], p["mlp.2.bias"])  # :122
    skip_sum = np.zeros_like(h)
    layers = []
    us = []
    for n in range(n_layers):
        pre = f"residual_layers.{n}."
        dil = 2 ** (n % dilation_cycle_length)                                   # :103
        d = linear(temb, p[pre + "diffusion_projection.weight"], p[pre + "diffusion_projection.bias"])  # :69
        if gemm_dtype == "bf16":
            # device contract: bf16(h) goes through the tensor cores, the (h+d) shift is applied
            # as an exact fp32 bias with the zero-padding edge correction
            y = conv1d(h, p[pre + "dilated_conv.weight"], p[pre + "dilated_conv.bias"], dilation=dil,
                       padding=dil, gemm_dtype="bf16")
            ones = np.ones((h.shape[0], 1, h.shape[2]), dtype=F32)
            dmap = d[:, :, None] * ones
            y = y + conv1d(dmap, p[pre + "dilated_conv.weight"], None, dilation=dil, padding=dil)
        else:
            y = conv1d(h + d[:, :, None], p[pre + "dilated_conv.weight"], p[pre + "dilated_conv.bias"],
                       dilation=dil, padding=dil)                                # :71,74 (pad after +d)
        y = y + conv1d(cond, p[pre + "conditioner_projection.weight"], p[pre + "conditioner_projection.bias"],
                       gemm_dtype=gemm_dtype)                                    # :70,74
        gate, filt = y[:, :C], y[:, C:]                                          # :76 chunk: gate first
        u = (sigmoid(gate) * np.tanh(filt)).astype(F32)                          # :77
        o = conv1d(u, p[pre + "output_projection.weight"], p[pre + "output_projection.bias"], gemm_dtype=gemm_dtype)
        res, skip = o[:, :C], o[:, C:]                                           # :80
        h = ((h + res) / F32(math.sqrt(2.0))).astype(F32)                        # :81
        skip_sum = (skip_sum + skip).astype(F32)
        us.append(u)
        if return_layers:
            layers.append((h.copy(), skip.copy()))
    if gemm_dtype == "bf16":
        # device contract: skip_projection(sum_l skip_l / sqrt(L)) is ONE contraction over the concatenated
        # gate outputs with the folded weight W_skip W_op,l[C:] / sqrt(L), rounded to bf16 after folding
        inv = 1.0 / math.sqrt(n_layers)
        Ws = p["skip_projection.weight"][:, :, 0].astype(np.float64)
        acc = np.zeros_like(h)
        bsum = np.zeros((C,), dtype=np.float64)
        for n in range(n_layers):
            Wo = p[f"residual_layers.{n}.output_projection.weight"][C:, :, 0].astype(np.float64)
            Wc = (Ws @ Wo * inv).astype(F32)
            acc += conv1d(us[n], Wc[:, :, None], None, gemm_dtype="bf16")
            bsum += p[f"residual_layers.{n}.output_projection.bias"][C:]
        bc = (p["skip_projection.bias"] + (Ws @ bsum) * inv).astype(F32)
        r = np.maximum(acc + bc[None, :, None], 0)
    else:
        s = (skip_sum / F32(math.sqrt(n_layers))).astype(F32)                    # :128
        r = np.maximum(conv1d(s, p["skip_projection.weight"], p["skip_projection.bias"]), 0)
    x0 = conv1d(r.astype(F32), p["output_projection.weight"], p["output_projection.bias"], gemm_dtype=gemm_dtype)
    if return_layers:
        return x0, layers
    return x0


# --------------------------------------------------------------------------------------
# posterior step + sampling loop (spec_denoiser.py:86-108, 177-185)
# --------------------------------------------------------------------------------------
def posterior_sample(sched, x0, x_t, t, z):
    """q_posterior_sample (spec_denoiser.py:95-101): x_{t-1} = c1[t] x0 + c2[t] x_t + [t!=0] exp(.5 logvar[t]) z."""
    c1 = sched["posterior_mean_coef1"][t][:, None, None]
    c2 = sched["posterior_mean_coef2"][t][:, None, None]
    lv = sched["posterior_log_variance_clipped"][t][:, None, None]
    nz = (1.0 - (t == 0).astype(F32))[:, None, None]
    mean = (c1 * x0 + c2 * x_t).astype(F32)
    return (mean + nz * np.exp(F32(0.5) * lv).astype(F32) * z).astype(F32)


def sample_loop(p, sched, cond, noise, timesteps, dilation_cycle_length=1, gemm_dtype="f32", trace=False):
    """GaussianDiffusion.forward(infer=True) loop (spec_denoiser.py:177-185) with INJECTED noise.

    cond[B,H,T]; noise[S+1,B,80,T]: noise[0] is x_S, noise[1+k] is the z drawn at the k-th
    iteration (t = S-1-k), including the wasted draw at t=0 (diffusion_utils.py:65-68).
    Returns mel_out[B,T,80] (= x[:,0].transpose(1,2), :183-184)."""
    B = cond.shape[0]
    x = noise[0].astype(F32)
    xs = []
    for k, i in enumerate(reversed(range(timesteps))):
        t = np.full((B,), i, dtype=np.int64)
        x0 = diffnet_forward(p, x, t, cond, dilation_cycle_length, gemm_dtype)
        x = posterior_sample(sched, x0, x, t, noise[1 + k])
        if trace:
            xs.append(x.copy())
    mel = np.ascontiguousarray(np.transpose(x, (0, 2, 1)))
    return (mel, xs) if trace else mel


def composite(mel_out, ref_mels, mask):
    """tasks/speech_editing/spec_denoiser.py:53,84: mel_out*m + ref*(1-m), m in {0,1} [B,T,1]."""
    return (mel_out * mask + ref_mels * (1 - mask)).astype(F32)


# --------------------------------------------------------------------------------------
# MelEncoder (modules/speech_editing/commons/mel_encoder.py:3-19)
# --------------------------------------------------------------------------------------
def mel_encoder_forward(p, x):
    h = np.maximum(linear(x, p["encoder.0.weight"], p["encoder.0.bias"]), 0)
    h = np.maximum(linear(h, p["encoder.2.weight"], p["encoder.2.bias"]), 0)
    return linear(h, p["fc_out.weight"], p["fc_out.bias"])


# --------------------------------------------------------------------------------------
# HiFi-GAN generator (modules/vocoder/hifigan/hifigan.py:27-64, 101-142)
# --------------------------------------------------------------------------------------
HIFIGAN_V1 = dict(
    upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
    resblock="1", resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
)


def weight_norm_fold(v: np.ndarray, g: np.ndarray) -> np.ndarray:
    """torch.nn.utils.weight_norm, dim=0: w = g * v / ||v||, norm over all dims but 0
    (dim 0 = C_out for Conv1d, C_in for ConvTranspose1d — SURVEY.md appendix A)."""
    norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
    return (v * (g / norm)).astype(F32)


def _wn(p, name):
    if name + ".weight" in p:
        return p[name + ".weight"]
    return weight_norm_fold(p[name + ".weight_v"], p[name + ".weight_g"])


def hifigan_forward(p: Dict[str, np.ndarray], cfg: dict, mel: np.ndarray, gemm_dtype: str = "f32") -> np.ndarray:
    """HifiGanGenerator.forward (hifigan.py:126-142), ResBlock1.forward (:51-58) / ResBlock2.forward (:80-85).
    mel[B,80,T] -> wav[B,1,T*prod(rates)]."""
    rb2 = str(cfg["resblock"]) == "2"
    nk = len(cfg["resblock_kernel_sizes"])
    x = conv1d(mel, _wn(p, "conv_pre"), p["conv_pre.bias"], padding=3, gemm_dtype=gemm_dtype)
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = leaky_relu(x, 0.1)
        x = conv_transpose1d(x, _wn(p, f"ups.{i}"), p[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2,
                             gemm_dtype=gemm_dtype)
        xs = None
        for j, (rk, rd) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            pre = f"resblocks.{i * nk + j}."
            y = x
            for m, d in enumerate(rd):
                yt = leaky_relu(y, 0.1)
                if rb2:                                   # ResBlock2: x = c(lrelu(x)) + x  (:81-84)
                    yt = conv1d(yt, _wn(p, pre + f"convs.{m}"), p[pre + f"convs.{m}.bias"], dilation=d, padding=(rk * d - d) // 2,
                                gemm_dtype=gemm_dtype)
                    y = (yt + y).astype(F32)
                    continue
                yt = conv1d(yt, _wn(p, pre + f"convs1.{m}"), p[pre + f"convs1.{m}.bias"], dilation=d,
                            padding=(rk * d - d) // 2, gemm_dtype=gemm_dtype)
                yt = leaky_relu(yt, 0.1)
                yt = conv1d(yt, _wn(p, pre + f"convs2.{m}"), p[pre + f"convs2.{m}.bias"], dilation=1,
                            padding=(rk - 1) // 2, gemm_dtype=gemm_dtype)
                y = (yt + y).astype(F32)
            xs = y if xs is None else (xs + y).astype(F32)
        x = (xs / F32(nk)).astype(F32)
    x = leaky_relu(x, 0.01)                      # F.leaky_relu default slope (:138)
    x = conv1d(x, _wn(p, "conv_post"), p["conv_post.bias"], padding=3)   # fp32 on device too (C_out=1)
    return np.tanh(x).astype(F32)


# --------------------------------------------------------------------------------------
# integer / index ops of the condition path that must be bit-exact
# --------------------------------------------------------------------------------------
def f0_to_coarse(f0, f0_bin=256, f0_max=900.0, f0_min=50.0):
    """utils/audio/pitch/utils.py:17-28."""
    f0_mel_min = 1127 * np.log(1 + f0_min / 700)
    f0_mel_max = 1127 * np.log(1 + f0_max / 700)
    f0 = np.asarray(f0, dtype=F32)
    f0_mel = (1127 * np.log(1 + f0.astype(F32) / F32(700))).astype(F32)
    nz = f0_mel > 0
    f0_mel = np.where(nz, (f0_mel - F32(f0_mel_min)) * F32(f0_bin - 2) / F32(f0_mel_max - f0_mel_min) + 1, f0_mel).astype(F32)
    f0_mel = np.where(f0_mel <= 1, F32(1), f0_mel)
    f0_mel = np.where(f0_mel > f0_bin - 1, F32(f0_bin - 1), f0_mel)
    return (f0_mel + F32(0.5)).astype(np.int64)


def mel2token_to_dur(mel2token, T_txt):
    """utils/audio/align.py:71-90: scatter_add of ones into T_txt+1 buckets, bucket 0 (padding) dropped."""
    mel2token = np.asarray(mel2token, dtype=np.int64)
    B = mel2token.shape[0]
    dur = np.zeros((B, T_txt + 1), dtype=np.int64)
    for b in range(B):
        np.add.at(dur[b], mel2token[b], 1)
    return dur[:, 1:]


def expand_states(h, mel2token):
    """modules/tts/commons/align_ops.py:21-25: pad a zero row in front, gather by 1-based index."""
    hp = np.pad(h, ((0, 0), (1, 0), (0, 0)))
    return np.take_along_axis(hp, np.asarray(mel2token, dtype=np.int64)[..., None].repeat(h.shape[-1], -1), axis=1)
