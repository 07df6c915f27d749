"""CPU oracle of the CampNet mask-predict forward (BASELINE.json configs[3]).

TEST INFRASTRUCTURE — NOT PRODUCT CODE (same rules as oracle/fluentspeech_oracle.py).

numpy restatement of (file:line relative to the reference tree, Zain-Jiang/Speech-Editing-Toolkit @ a8d5bf33):
  modules/speech_editing/campnet/campnet.py:40-69            CampNet.forward / run_text_encoder / run_decoder
  modules/speech_editing/commons/transformer.py:14-71        SinusoidalPositionalEmbedding (+ utils/nn/seq_utils.py:6-18 make_positions)
  modules/speech_editing/commons/transformer.py:74-110       TransformerFFNLayer (conv k9 'SAME' / 'LEFT', * k^-0.5, GELU, Linear)
  modules/speech_editing/commons/transformer.py:138-419      MultiheadAttention (2 heads x 96, bias=False; q scaled by d^-0.5;
                                                             fp32 softmax; padded keys masked)
  modules/speech_editing/commons/transformer.py:489-608      EncSALayer / DecSALayer (pre-LN residual blocks)
  modules/speech_editing/commons/transformer.py:639-812      FFTBlocks / TransformerEncoder / TransformerDecoder
  modules/commons/conv.py:68-116                             ConvBlocks (decoder_fine)
  modules/speech_editing/commons/mel_encoder.py:3-19         MelEncoder

Parity status: PINNED to outputs of the unmodified reference CampNet (tests/golden/campnet.npz, written by
oracle/make_golden.py campnet); the reference ships no tests or golden vectors of its own.

`gemm_dtype="bf16"` rounds the operands of every contraction the kernels run on tensor cores (the q/k/v/out projections, the
FFN convolutions, the MelEncoder and ConvBlocks layers, the two 192 -> 80 output projections) to bfloat16; the attention
contractions QK^T / PV follow `attn_dtype`.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

from .cond_encoder_oracle import gelu, layer_norm_c
from .fluentspeech_oracle import F32, _q, conv1d, linear

HEADS = 2


def sinusoid_table(n: int, dim: int, padding_idx: int = 0) -> np.ndarray:
    """SinusoidalPositionalEmbedding.get_embedding (transformer.py:33-49): [sin | cos], padding row zeroed (fp32 arithmetic)."""
    half = dim // 2
    step = F32(math.log(10000) / (half - 1))
    freq = np.exp(np.arange(half, dtype=F32) * -step).astype(F32)
    ang = (np.arange(n, dtype=F32)[:, None] * freq[None, :]).astype(F32)
    tab = np.concatenate([np.sin(ang), np.cos(ang)], axis=1).astype(F32)
    tab[padding_idx] = 0
    return tab


def make_positions(x, padding_idx=0):
    """utils/nn/seq_utils.py:6-18: non-padding symbols numbered 1, 2, ...; padding keeps padding_idx."""
    mask = (x != padding_idx).astype(np.int64)
    return np.cumsum(mask, axis=1) * mask + padding_idx


def layer_norm_last(x, w, b, eps=1e-5):
    return layer_norm_c(x.transpose(0, 2, 1), w, b, eps).transpose(0, 2, 1)


def attention(q_in, kv_in, w_in, w_out, key_pad, gemm_dtype="f32", attn_dtype="f32", need_weights=False):
    """MultiheadAttention.forward with bias=False (transformer.py:205-419): q_in[B,Tq,C], kv_in[B,Tk,C], key_pad[B,Tk] bool or None.
    Returns out[B,Tq,C] (and the head-averaged probabilities [B,Tq,Tk] when need_weights)."""
    B, Tq, C = q_in.shape
    d = C // HEADS
    zero = np.zeros(C, dtype=F32)
    q = (linear(q_in, w_in[:C], zero, gemm_dtype) * F32(d ** -0.5)).astype(F32)
    k = linear(kv_in, w_in[C:2 * C], zero, gemm_dtype)
    v = linear(kv_in, w_in[2 * C:], zero, gemm_dtype)
    out = np.zeros_like(q)
    probs = np.zeros((B, Tq, kv_in.shape[1]), dtype=F32)
    for h in range(HEADS):
        sl = slice(h * d, (h + 1) * d)
        s = np.einsum("bqd,bkd->bqk", _q(q[..., sl], attn_dtype), _q(k[..., sl], attn_dtype), optimize=True).astype(F32)
        if key_pad is not None:
            s = np.where(key_pad[:, None, :], F32(-1e8), s)
        s64 = s.astype(np.float64)
        p = np.exp(s64 - s64.max(-1, keepdims=True))
        p = (p / p.sum(-1, keepdims=True)).astype(F32)
        probs += p / HEADS
        out[..., sl] = np.einsum("bqk,bkd->bqd", _q(p, attn_dtype), _q(v[..., sl], attn_dtype), optimize=True).astype(F32)
    y = linear(out, w_out, zero, gemm_dtype)
    return (y, probs) if need_weights else y


def ffn(p, pre, x, k, left, gemm_dtype):
    """TransformerFFNLayer.forward (transformer.py:92-110): conv k (SAME or LEFT padding) * k^-0.5 -> GELU -> Linear."""
    pre = pre + "ffn."
    w = p.get(pre + "ffn_1.weight", p.get(pre + "ffn_1.1.weight"))
    b = p.get(pre + "ffn_1.bias", p.get(pre + "ffn_1.1.bias"))
    xc = x.transpose(0, 2, 1)
    if left:
        xc = np.pad(xc, ((0, 0), (0, 0), (k - 1, 0)))
        y = conv1d(xc, w, b, gemm_dtype=gemm_dtype)
    else:
        y = conv1d(xc, w, b, padding=k // 2, gemm_dtype=gemm_dtype)
    y = gelu((y * F32(k ** -0.5)).astype(F32)).transpose(0, 2, 1)
    return linear(y, p[pre + "ffn_2.weight"], p[pre + "ffn_2.bias"], gemm_dtype)


def text_encoder(p, txt, k=9, layers=3, gemm_dtype="f32", attn_dtype="f32"):
    """TransformerEncoder.forward (transformer.py:729-747) + FFTBlocks.forward (:664-691) with use_pos_embed=False."""
    C = p["encoder.embed_tokens.weight"].shape[1]
    pad = txt == 0
    tab = sinusoid_table(max(2000, txt.shape[1] + 2), C)
    x = (F32(math.sqrt(C)) * p["encoder.embed_tokens.weight"][txt] + tab[make_positions(txt)]).astype(F32)
    keep = (1 - pad.astype(F32))[:, :, None]
    x = x * keep
    for i in range(layers):
        pre = f"encoder.layers.{i}.op."
        y = layer_norm_last(x, p[pre + "layer_norm1.weight"], p[pre + "layer_norm1.bias"])
        y = attention(y, y, p[pre + "self_attn.in_proj_weight"], p[pre + "self_attn.out_proj.weight"], pad, gemm_dtype, attn_dtype)
        x = ((x + y) * keep).astype(F32)
        y = layer_norm_last(x, p[pre + "layer_norm2.weight"], p[pre + "layer_norm2.bias"])
        x = ((x + ffn(p, pre, y, k, False, gemm_dtype)) * keep).astype(F32)
    return (layer_norm_last(x, p["encoder.layer_norm.weight"], p["encoder.layer_norm.bias"]) * keep).astype(F32)


def coarse_decoder(p, x, enc, k=9, layers=6, gemm_dtype="f32", attn_dtype="f32"):
    """TransformerDecoder.forward (transformer.py:777-812) + DecSALayer.forward (:548-608)."""
    C = x.shape[-1]
    enc_pad = np.abs(enc).sum(-1) == 0
    pad = np.abs(x).sum(-1) == 0
    keep = (1 - pad.astype(F32))[:, :, None]
    tab = sinusoid_table(max(2000, x.shape[1] + 2), C)
    x = (x + p["decoder_coarse.pos_embed_alpha"].astype(F32) * tab[make_positions(x[..., 0])]).astype(F32)
    x = x * keep
    attn0 = None
    for i in range(layers):
        pre = f"decoder_coarse.layers.{i}.op."
        y = layer_norm_last(x, p[pre + "layer_norm1.weight"], p[pre + "layer_norm1.bias"])
        y = attention(y, y, p[pre + "self_attn.in_proj_weight"], p[pre + "self_attn.out_proj.weight"], None, gemm_dtype, attn_dtype)
        x = (x + y).astype(F32)
        y = layer_norm_last(x, p[pre + "layer_norm2.weight"], p[pre + "layer_norm2.bias"])
        y, w = attention(y, enc, p[pre + "encoder_attn.in_proj_weight"], p[pre + "encoder_attn.out_proj.weight"], enc_pad, gemm_dtype,
                         attn_dtype, need_weights=True)
        if i == 0:
            attn0 = w
        x = (x + y).astype(F32)
        y = layer_norm_last(x, p[pre + "layer_norm3.weight"], p[pre + "layer_norm3.bias"])
        x = ((x + ffn(p, pre, y, k, True, gemm_dtype)) * keep).astype(F32)
    x = (layer_norm_last(x, p["decoder_coarse.layer_norm.weight"], p["decoder_coarse.layer_norm.bias"]) * keep).astype(F32)
    return x, attn0


def conv_blocks(p, prefix, x, n_blocks=5, k=5, layers_in_block=2, post_k=3, gemm_dtype="f32"):
    """ConvBlocks.forward (conv.py:99-116) on x[B,T,C] with data-derived nonpadding (decoder_fine)."""
    x = x.transpose(0, 2, 1)
    nonpad0 = (np.abs(x).sum(1, keepdims=True) > 0).astype(F32)
    for i in range(n_blocks):
        nonpad = (np.abs(x).sum(1, keepdims=True) > 0).astype(F32)
        for j in range(layers_in_block):
            pre = f"{prefix}res_blocks.{i}.blocks.{j}."
            y = layer_norm_c(x, p[pre + "0.weight"], p[pre + "0.bias"])
            y = conv1d(y, p[pre + "1.weight"], p[pre + "1.bias"], padding=(k - 1) // 2, gemm_dtype=gemm_dtype)
            y = gelu((y * F32(k ** -0.5)).astype(F32))
            y = conv1d(y, p[pre + "4.weight"], p[pre + "4.bias"], gemm_dtype=gemm_dtype)
            x = ((x + y) * nonpad).astype(F32)
    x = (x * nonpad0).astype(F32)
    x = (layer_norm_c(x, p[prefix + "last_norm.weight"], p[prefix + "last_norm.bias"]) * nonpad0).astype(F32)
    x = conv1d(x, p[prefix + "post_net1.weight"], p[prefix + "post_net1.bias"], padding=post_k // 2, gemm_dtype=gemm_dtype) * nonpad0
    return x.transpose(0, 2, 1).astype(F32)


def campnet_forward(p: Dict[str, np.ndarray], txt, mels, time_mel_masks, k=9, gemm_dtype="f32", attn_dtype="f32"):
    """CampNet.forward (campnet.py:40-69): txt[B,Tt] int64, mels[B,T,80], time_mel_masks[B,T,1] (0/1) ->
    dict(mel_out_coarse, mel_out_fine [B,T,80], attn [B,T,Tt], encoder_out [B,Tt,C])."""
    m = time_mel_masks.reshape(mels.shape[0], mels.shape[1], 1).astype(F32)
    src_keep = (txt > 0).astype(F32)[:, :, None]
    enc = (text_encoder(p, txt, k, gemm_dtype=gemm_dtype, attn_dtype=attn_dtype) * src_keep * src_keep).astype(F32)
    nonpad = (np.abs(mels).sum(-1) > 0).astype(F32)[:, :, None]
    me = {kk[len("mel_encoder."):]: v for kk, v in p.items() if kk.startswith("mel_encoder.")}
    zero80 = np.zeros(mels.shape[-1], dtype=F32)

    def mel_enc(x):
        h = np.maximum(linear(x, me["encoder.0.weight"], me["encoder.0.bias"], gemm_dtype), 0)
        h = np.maximum(linear(h, me["encoder.2.weight"], me["encoder.2.bias"], gemm_dtype), 0)
        return linear(h, me["fc_out.weight"], me["fc_out.bias"], gemm_dtype)

    x = (mels * (1 - m) + p["mask_emb"].reshape(1, 1, -1) * m).astype(F32)
    x = (mel_enc(x) * nonpad).astype(F32)
    x, attn = coarse_decoder(p, x, enc, k, gemm_dtype=gemm_dtype, attn_dtype=attn_dtype)
    x = (x * nonpad).astype(F32)
    coarse = (linear(x, p["mel_out_coarse.weight"], zero80, gemm_dtype) * nonpad).astype(F32)
    mel_coarse = (mels * (1 - m) + coarse * m).astype(F32)
    y = (mel_enc(mel_coarse) * nonpad).astype(F32)
    y = (conv_blocks(p, "decoder_fine.", y, gemm_dtype=gemm_dtype) * nonpad).astype(F32)
    fine = (linear(y, p["mel_out_fine.weight"], zero80, gemm_dtype) * nonpad).astype(F32)
    fine = (mel_coarse + fine * m).astype(F32)
    return {"mel_out_coarse": coarse, "mel_out_fine": fine, "attn": attn, "encoder_out": enc}
