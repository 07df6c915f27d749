"""CPU oracle of the FluentSpeech condition encoder: FastSpeech.forward(skip_decoder=True) and the pieces the
inference script calls on their own (encoder, forward_style_embed, forward_dur, LengthRegulator).

TEST INFRASTRUCTURE — NOT PRODUCT CODE (same rules as oracle/fluentspeech_oracle.py: only tests/, smoke() and
bench.py's CPU legs may import it).

numpy restatement of (all file:line relative to the reference tree, Zain-Jiang/Speech-Editing-Toolkit @ a8d5bf33):
  modules/speech_editing/spec_denoiser/fs.py:83-189           FastSpeech.forward / forward_style_embed / forward_dur / forward_pitch
  modules/commons/conv.py:24-139                              ResidualBlock, ConvBlocks, TextConvEncoder
  modules/commons/nar_tts_modules.py:8-100                    DurationPredictor, LengthRegulator, PitchPredictor
  modules/commons/layers.py:5-24                              channel LayerNorm
  modules/tts/commons/align_ops.py:15-25                      clip_mel2token_to_multiple, expand_states
  utils/audio/align.py:71-90, utils/audio/pitch/utils.py:17-28,71-82   mel2token_to_dur, f0_to_coarse, denorm_f0

Parity status: PINNED to outputs of the unmodified reference (tests/golden/cond_encoder.npz, written by
oracle/make_golden.py cond_encoder); the reference ships no tests or golden vectors of its own.

`gemm_dtype="bf16"` rounds the operands of every contraction the kernels run on tensor cores (all Conv1d layers) to
bfloat16; LayerNorm, GELU, the embedding adds, the two predictor heads (192 -> 1 / 2) and spk_embed_proj stay fp32,
as on the device.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
from scipy.special import erf

from .fluentspeech_oracle import F32, conv1d, expand_states, f0_to_coarse, mel2token_to_dur

DEFAULT_HP = dict(hidden_size=192, enc_dilations=[1, 1, 1, 1], enc_kernel_size=5, layers_in_block=2, enc_post_net_kernel=3,
                  dur_predictor_layers=3, dur_predictor_kernel=5, predictor_kernel=5, pitch_predictor_layers=5,
                  use_pitch_embed=True, use_uv=True, pitch_type="frame", use_spk_embed=True, frames_multiple=1)


def layer_norm_c(x, w, b, eps=1e-5):
    """layers.py:5-24 (LayerNorm over the channel axis of a [B,C,T] tensor); biased variance, eps inside the sqrt."""
    x64 = x.astype(np.float64)
    mu = x64.mean(axis=1, keepdims=True)
    var = ((x64 - mu) ** 2).mean(axis=1, keepdims=True)
    y = (x64 - mu) / np.sqrt(var + eps)
    return (y * w[None, :, None] + b[None, :, None]).astype(F32)


def gelu(x):
    """torch.nn.GELU() (exact, erf form)."""
    x64 = x.astype(np.float64)
    return (0.5 * x64 * (1.0 + erf(x64 / math.sqrt(2.0)))).astype(F32)


def softplus(x):
    """torch.nn.Softplus(beta=1, threshold=20)."""
    x64 = x.astype(np.float64)
    return np.where(x64 > 20.0, x64, np.log1p(np.exp(np.minimum(x64, 20.0)))).astype(F32)


def _nonpadding_c(x):
    """conv.py:58,104: (x.abs().sum(1) > 0) of a [B,C,T] tensor -> [B,1,T] float."""
    return (np.abs(x).sum(axis=1, keepdims=True) > 0).astype(F32)


def text_encoder(p: Dict[str, np.ndarray], txt: np.ndarray, hp: Optional[dict] = None, gemm_dtype: str = "f32") -> np.ndarray:
    """TextConvEncoder.forward (conv.py:130-139) -> ConvBlocks.forward (:99-116) -> ResidualBlock.forward (:57-65).
    txt[B,Tt] int64 -> encoder_out[B,Tt,H]."""
    hp = {**DEFAULT_HP, **(hp or {})}
    H, k = hp["hidden_size"], hp["enc_kernel_size"]
    x = (F32(math.sqrt(H)) * p["encoder.embed_tokens.weight"][txt]).astype(F32)       # embed_scale * embed_tokens(txt)
    x = x.transpose(0, 2, 1)                                                           # [B,H,Tt]
    nonpad0 = _nonpadding_c(x)
    for i, d in enumerate(hp["enc_dilations"]):
        nonpad = _nonpadding_c(x)                                                      # from the block's own input
        for j in range(hp["layers_in_block"]):
            pre = f"encoder.res_blocks.{i}.blocks.{j}."
            y = layer_norm_c(x, p[pre + "0.weight"], p[pre + "0.bias"])
            y = conv1d(y, p[pre + "1.weight"], p[pre + "1.bias"], dilation=d, padding=d * (k - 1) // 2, gemm_dtype=gemm_dtype)
            y = gelu((y * F32(k ** -0.5)).astype(F32))
            y = conv1d(y, p[pre + "4.weight"], p[pre + "4.bias"], gemm_dtype=gemm_dtype)
            x = ((x + y) * nonpad).astype(F32)
    x = (x * nonpad0).astype(F32)
    x = (layer_norm_c(x, p["encoder.last_norm.weight"], p["encoder.last_norm.bias"]) * nonpad0).astype(F32)
    kp = hp["enc_post_net_kernel"]
    x = (conv1d(x, p["encoder.post_net1.weight"], p["encoder.post_net1.bias"], padding=kp // 2, gemm_dtype=gemm_dtype) * nonpad0)
    return x.transpose(0, 2, 1).astype(F32)


def style_embed(p, spk_embed):
    """fs.py:114-121 with use_spk_embed: spk_embed_proj(spk_embed)[:, None, :] (fp32 on the device too)."""
    return (spk_embed.astype(np.float64) @ p["spk_embed_proj.weight"].astype(np.float64).T
            + p["spk_embed_proj.bias"]).astype(F32)[:, None, :]


def masked_dur_gt(mel2ph, time_mel_masks, txt):
    """fs.py:136-138: mel2token_to_dur(mel2ph * (1 - mask).long(), Tt) * nonpadding  (integer, bit-exact)."""
    keep = (1 - time_mel_masks.reshape(mel2ph.shape)).astype(np.int64)
    return mel2token_to_dur(mel2ph * keep, txt.shape[1]) * (txt != 0).astype(np.int64)


def _predictor_stack(p, prefix, n_layers, k, x_bct, keep=None, gemm_dtype="f32"):
    for i in range(n_layers):
        x_bct = conv1d(x_bct, p[f"{prefix}.conv.{i}.0.weight"], p[f"{prefix}.conv.{i}.0.bias"], padding=k // 2, gemm_dtype=gemm_dtype)
        x_bct = np.maximum(x_bct, 0)
        x_bct = layer_norm_c(x_bct, p[f"{prefix}.conv.{i}.2.weight"], p[f"{prefix}.conv.{i}.2.bias"])
        if keep is not None:
            x_bct = (x_bct * keep[:, None, :]).astype(F32)
    return x_bct


def duration_predictor(p, dur_inp, src_padding, hp=None, gemm_dtype="f32"):
    """DurationPredictor.forward (nar_tts_modules.py:24-34): dur_inp[B,Tt,H], src_padding[B,Tt] bool -> dur[B,Tt]."""
    hp = {**DEFAULT_HP, **(hp or {})}
    keep = (1 - src_padding.astype(F32)).astype(F32)
    x = _predictor_stack(p, "dur_predictor", hp["dur_predictor_layers"], hp["dur_predictor_kernel"], dur_inp.transpose(0, 2, 1),
                         keep, gemm_dtype)
    y = x.transpose(0, 2, 1).astype(np.float64) @ p["dur_predictor.linear.0.weight"].astype(np.float64).T + p["dur_predictor.linear.0.bias"]
    return (softplus(y.astype(F32)) * keep[:, :, None])[..., 0].astype(F32)


def length_regulator(dur, dur_padding=None):
    """LengthRegulator.forward (nar_tts_modules.py:42-72), alpha = 1: round half to even, 1-based token index per frame."""
    d = np.rint(dur.astype(F32)).astype(np.int64)
    if dur_padding is not None:
        d = d * (1 - dur_padding.astype(np.int64))
    B, Tt = d.shape
    T = int(d.sum(-1).max())
    out = np.zeros((B, T), dtype=np.int64)
    for b in range(B):
        pos = 0
        for i in range(Tt):
            out[b, pos:pos + d[b, i]] = i + 1
            pos += int(d[b, i])
    return out


def denorm_f0(f0, uv, pitch_padding=None, lo=50.0, hi=900.0):
    """utils/audio/pitch/utils.py:71-82 with pitch_norm='log'."""
    y = np.clip(np.exp2(f0.astype(np.float64)).astype(F32), F32(lo), F32(hi)).astype(F32)
    if uv is not None:
        y = np.where(uv > 0, F32(0), y)
    if pitch_padding is not None:
        y = np.where(pitch_padding, F32(0), y)
    return y.astype(F32)


def pitch_predictor(p, x, hp=None, gemm_dtype="f32"):
    """PitchPredictor.forward (nar_tts_modules.py:90-100): x[B,T,H] -> [B,T,2]."""
    hp = {**DEFAULT_HP, **(hp or {})}
    y = _predictor_stack(p, "pitch_predictor", hp["pitch_predictor_layers"], hp["predictor_kernel"], x.transpose(0, 2, 1), None, gemm_dtype)
    out = y.transpose(0, 2, 1).astype(np.float64) @ p["pitch_predictor.linear.weight"].astype(np.float64).T + p["pitch_predictor.linear.bias"]
    return out.astype(F32)


def fastspeech_forward(p, txt, time_mel_masks, mel2ph, spk_embed, f0, uv, hp=None, use_pred_mel2ph=False, use_pred_pitch=False,
                       gemm_dtype="f32"):
    """FastSpeech.forward(..., skip_decoder=True) (fs.py:83-105).  time_mel_masks[B,T,1] (or [B,T]) float 0/1.
    Returns the reference's `ret` dict (decoder_inp, dur, mel2ph, pitch_pred, f0_denorm, f0_denorm_pred) plus the
    intermediates the tests pin (encoder_out, style_embed, masked_dur_gt, masked_gt_pitch, pitch)."""
    hp = {**DEFAULT_HP, **(hp or {})}
    ret = {}
    B, Tt = txt.shape
    enc = text_encoder(p, txt, hp, gemm_dtype)
    src_nonpad = (txt > 0).astype(F32)[:, :, None]
    style = style_embed(p, spk_embed) if hp["use_spk_embed"] else F32(0)
    dur_inp = ((enc + style) * src_nonpad).astype(F32)
    m = time_mel_masks.reshape(mel2ph.shape[0], -1).astype(F32)
    # forward_dur (fs.py:123-151); the predictor_grad mix (:145-146) is the identity in the forward value
    mdur = masked_dur_gt(mel2ph, m, txt)
    dur_inp = (dur_inp + p["dur_embed.weight"][mdur]).astype(F32)
    src_padding = txt == 0
    ret["dur"] = duration_predictor(p, dur_inp, src_padding, hp, gemm_dtype)
    if use_pred_mel2ph:
        mel2ph = length_regulator(ret["dur"], src_padding)
    fm = hp["frames_multiple"]
    mel2ph = mel2ph[:, :mel2ph.shape[1] // fm * fm]                                  # clip_mel2token_to_multiple
    ret["mel2ph"] = mel2ph
    tgt_nonpad = (mel2ph > 0).astype(F32)[:, :, None]
    dec = expand_states(enc, mel2ph)
    ret.update(encoder_out=enc, style_embed=style, masked_dur_gt=mdur)
    if hp["use_pitch_embed"]:
        use_uv = hp["pitch_type"] == "frame" and hp["use_uv"]
        pitch_inp = ((dec + style) * tgt_nonpad).astype(F32)
        pad = mel2ph == 0
        mf0 = (f0 * (1 - m)).astype(F32)
        muv = (uv * (1 - m)).astype(F32)
        gt_pitch = f0_to_coarse(denorm_f0(mf0, muv if use_uv else None, pad))
        pitch_inp = (pitch_inp + p["pitch_embed.weight"][gt_pitch]).astype(F32)
        ret["pitch_pred"] = pp = pitch_predictor(p, pitch_inp, hp, gemm_dtype)
        if use_pred_pitch:
            pad2 = None
            res_f0 = (f0 * (1 - m) + pp[:, :, 0] * m).astype(F32)
            res_uv = (uv * (1 - m) + (pp[:, :, 1] > 0).astype(F32) * m).astype(F32)
        else:
            pad2, res_f0, res_uv = pad, f0, uv
        ret["f0_denorm"] = fd = denorm_f0(res_f0, res_uv if use_uv else None, pad2)
        pitch = f0_to_coarse(fd)
        ret["f0_denorm_pred"] = denorm_f0(pp[:, :, 0], (pp[:, :, 1] > 0) if use_uv else None, pad2)
        dec = (dec + p["pitch_embed.weight"][pitch]).astype(F32)
        ret.update(masked_gt_pitch=gt_pitch, pitch=pitch)
    ret["decoder_inp"] = ((dec + style) * tgt_nonpad).astype(F32)
    return ret
