"""Seeded synthetic weights and inputs of the benchmark / parity workloads.

Pretrained FluentSpeech / HiFi-GAN checkpoints are Google-Drive artefacts that are not available
offline, so parity and throughput are measured on seeded random weights with the reference's
state_dict layout.  numpy.random.RandomState is used because its stream is frozen across numpy
versions (fixtures under tests/golden/ depend on it).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

F32 = np.float32


def denoiser_state_dict(seed: int = 1234, n_mels: int = 80, hidden: int = 192, channels: int = 256,
                        layers: int = 20) -> Dict[str, np.ndarray]:
    """Random `denoise_fn.*` state_dict with the reference's keys/shapes (diffnet.py:84-108).

    Conv weights are kaiming-normal like the reference's Conv1d helper (diffnet.py:49-52); every bias is
    randomised and the zero-initialised output_projection (diffnet.py:108) gets N(0, 0.05) so that parity
    tests are not vacuous (SURVEY.md §8a zero-init trap)."""
    rs = np.random.RandomState(seed)
    C, H, M = channels, hidden, n_mels
    sd: Dict[str, np.ndarray] = {}

    def conv(name, co, ci, k, std=None):
        std = math.sqrt(2.0 / (ci * k)) if std is None else std
        sd[name + ".weight"] = (rs.standard_normal((co, ci, k)) * std).astype(F32)
        sd[name + ".bias"] = (rs.standard_normal((co,)) * 0.05).astype(F32)

    def lin(name, co, ci):
        b = 1.0 / math.sqrt(ci)
        sd[name + ".weight"] = rs.uniform(-b, b, (co, ci)).astype(F32)
        sd[name + ".bias"] = rs.uniform(-b, b, (co,)).astype(F32)

    conv("input_projection", C, M, 1)
    lin("mlp.0", 4 * C, C)
    lin("mlp.2", C, 4 * C)
    for n in range(layers):
        p = f"residual_layers.{n}."
        conv(p + "dilated_conv", 2 * C, C, 3)
        lin(p + "diffusion_projection", C, C)
        conv(p + "conditioner_projection", 2 * C, H, 1)
        conv(p + "output_projection", 2 * C, C, 1)
    conv("skip_projection", C, C, 1)
    conv("output_projection", M, C, 1, std=0.05)
    return sd


def hifigan_state_dict(seed: int = 1234, config: dict = None, n_mels: int = 80) -> Dict[str, np.ndarray]:
    """Random HifiGanGenerator state_dict (weight_g / weight_v / bias per conv; hifigan.py:101-124).
    weight_g is drawn independently of ||weight_v|| so that the weight-norm fold is exercised."""
    from .engine import HIFIGAN_V1
    cfg = dict(HIFIGAN_V1 if config is None else config)
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}

    def wn(name, shape, fan_in, gain=1.0):
        v = (rs.standard_normal(shape) * gain / math.sqrt(fan_in)).astype(F32)
        norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
        sd[name + ".weight_v"] = v
        sd[name + ".weight_g"] = (norm * rs.uniform(0.8, 1.2, norm.shape)).astype(F32)

    c0 = cfg["upsample_initial_channel"]
    wn("conv_pre", (c0, n_mels, 7), n_mels * 7)
    sd["conv_pre.bias"] = (rs.standard_normal((c0,)) * 0.02).astype(F32)
    nk = len(cfg["resblock_kernel_sizes"])
    ch = c0
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        cin, ch = ch, c0 // (2 ** (i + 1))
        wn(f"ups.{i}", (cin, ch, k), cin * k / u, gain=1.4)           # ConvTranspose1d weight is [C_in, C_out, k]
        sd[f"ups.{i}.bias"] = (rs.standard_normal((ch,)) * 0.02).astype(F32)
        rb2 = str(cfg.get("resblock", "1")) == "2"
        for j, rk in enumerate(cfg["resblock_kernel_sizes"]):
            for m in range(2 if rb2 else 3):
                for part in (("convs",) if rb2 else ("convs1", "convs2")):
                    name = f"resblocks.{i * nk + j}.{part}.{m}"
                    wn(name, (ch, ch, rk), ch * rk, gain=0.7)
                    sd[name + ".bias"] = (rs.standard_normal((ch,)) * 0.02).astype(F32)
    wn("conv_post", (1, ch, 7), ch * 7, gain=0.12)
    sd["conv_post.bias"] = (rs.standard_normal((1,)) * 0.02).astype(F32)
    return sd


def mel_encoder_state_dict(seed: int = 1234, n_mels: int = 80, hidden: int = 192) -> Dict[str, np.ndarray]:
    """Random MelEncoder state_dict (mel_encoder.py:4-13): encoder.0 / encoder.2 / fc_out Linear layers."""
    rs = np.random.RandomState(seed + 29)
    sd: Dict[str, np.ndarray] = {}
    for name, (n, c) in (("encoder.0", (hidden, n_mels)), ("encoder.2", (hidden, hidden)), ("fc_out", (hidden, hidden))):
        sd[name + ".weight"] = (rs.standard_normal((n, c)) / math.sqrt(c)).astype(F32)
        sd[name + ".bias"] = (rs.standard_normal((n,)) * 0.1).astype(F32)
    return sd


def synthetic_ref_and_mask(seed: int, B: int, T: int, n_mels: int = 80):
    """ref_mels ~ clip(N(-3, 1.5), -6, 1.5) and a contiguous 30 % edit span per item."""
    rs = np.random.RandomState(seed + 31)
    ref = np.clip(rs.standard_normal((B, T, n_mels)) * 1.5 - 3.0, -6.0, 1.5).astype(F32)
    mask = np.zeros((B, T, 1), dtype=F32)
    for b in range(B):
        n = max(1, int(round(0.3 * T)))
        s0 = int(rs.randint(0, T - n + 1))
        mask[b, s0:s0 + n] = 1.0
    return ref, mask


def synthetic_cond(seed: int, B: int, T: int, hidden: int = 192) -> np.ndarray:
    """Stand-in for the condition-encoder output decoder_inp[B,T,H] (spec_denoiser.py:159-167)."""
    rs = np.random.RandomState(seed + 17)
    return rs.standard_normal((B, T, hidden)).astype(F32)


def synthetic_noise(seed: int, S: int, B: int, T: int, n_mels: int = 80) -> np.ndarray:
    """Injected N(0,1) draws [(S+1),B,M,T]: x_S then one per iteration (spec_denoiser.py:98,180)."""
    rs = np.random.RandomState(seed + 29)
    return rs.standard_normal((S + 1, B, n_mels, T)).astype(F32)


def synthetic_edit_batch(seed: int, B: int, T: int, n_mels: int = 80, vocab: int = 80, frames_per_phone: int = 8):
    """Synthetic (text-token, mel-frame) editing batch of SURVEY.md §8d: uniform 8 frames/phone alignment,
    log-mel reference in [-6, 1.5], a contiguous 30 % phone span masked for re-synthesis."""
    rs = np.random.RandomState(seed + 41)
    Tt = max(T // frames_per_phone, 1)
    txt = rs.randint(3, vocab, size=(B, Tt)).astype(np.int64)
    mel2ph = (np.arange(T)[None, :] // frames_per_phone + 1).clip(max=Tt).repeat(B, 0).astype(np.int64)
    ref = np.clip(rs.standard_normal((B, T, n_mels)) * 1.5 - 3.0, -6.0, 1.5).astype(F32)
    mask = np.zeros((B, T), dtype=F32)
    span = max(int(round(0.3 * Tt)), 1)
    for b in range(B):
        s = rs.randint(0, Tt - span + 1)
        ph = mel2ph[b]
        mask[b] = ((ph > s) & (ph <= s + span)).astype(F32)
    f0 = rs.uniform(6.5, 8.5, (B, T)).astype(F32)
    uv = (rs.uniform(size=(B, T)) < 0.3).astype(F32)
    spk = (rs.standard_normal((B, 256)) / 16).astype(F32)
    return dict(txt_tokens=txt, mel2ph=mel2ph, ref_mels=ref, time_mel_masks=mask, f0=f0, uv=uv, spk_embed=spk)


FS_DEFAULTS = dict(hidden_size=192, enc_dilations=[1, 1, 1, 1], enc_kernel_size=5, layers_in_block=2, enc_post_net_kernel=3,
                   dur_predictor_layers=3, dur_predictor_kernel=5, predictor_kernel=5, pitch_predictor_layers=5,
                   use_pitch_embed=True, use_uv=True, pitch_type="frame", use_spk_embed=True, frames_multiple=1)


def fastspeech_state_dict(seed: int = 1234, vocab: int = 80, hp: dict = None) -> Dict[str, np.ndarray]:
    """Random `fs.*` state_dict (keys without the `fs.` prefix) of the condition encoder as the hot path uses it
    (fs.py:49-82 with encoder_type 'conv'): TextConvEncoder, spk_embed_proj, dur_embed / dur_predictor, pitch_embed /
    pitch_predictor.  The unused `decoder.*` / `mel_out.*` entries (skip_decoder=True) are not generated.
    LayerNorm affine parameters and all biases are randomised so that parity checks are not vacuous; the predictor heads
    are biased so that durations round to 1-4 frames and predicted log2-f0 lands in the voiced range."""
    hp = {**FS_DEFAULTS, **(hp or {})}
    rs = np.random.RandomState(seed + 53)
    H = hp["hidden_size"]
    sd: Dict[str, np.ndarray] = {}

    def conv(name, co, ci, k, gain=1.0):
        sd[name + ".weight"] = (rs.standard_normal((co, ci, k)) * (gain / math.sqrt(ci * k))).astype(F32)
        sd[name + ".bias"] = (rs.standard_normal((co,)) * 0.1).astype(F32)

    def ln(name, c):
        sd[name + ".weight"] = (1.0 + 0.1 * rs.standard_normal((c,))).astype(F32)
        sd[name + ".bias"] = (0.1 * rs.standard_normal((c,))).astype(F32)

    def emb(name, n, c):
        w = (rs.standard_normal((n, c)) * c ** -0.5).astype(F32)
        w[0] = 0.0                                                      # padding_idx = 0 (layers.py:45-50)
        sd[name + ".weight"] = w

    k = hp["enc_kernel_size"]
    for i in range(len(hp["enc_dilations"])):
        for j in range(hp["layers_in_block"]):
            p = f"encoder.res_blocks.{i}.blocks.{j}."
            ln(p + "0", H)
            conv(p + "1", 2 * H, H, k, gain=1.6)
            conv(p + "4", H, 2 * H, 1)
    ln("encoder.last_norm", H)
    conv("encoder.post_net1", H, H, hp["enc_post_net_kernel"])
    emb("encoder.embed_tokens", vocab, H)
    sd["spk_embed_proj.weight"] = (rs.standard_normal((H, 256)) / 16.0).astype(F32)
    sd["spk_embed_proj.bias"] = (rs.standard_normal((H,)) * 0.1).astype(F32)
    emb("dur_embed", 2000, H)
    for i in range(hp["dur_predictor_layers"]):
        conv(f"dur_predictor.conv.{i}.0", H, H, hp["dur_predictor_kernel"], gain=1.4)
        ln(f"dur_predictor.conv.{i}.2", H)
    sd["dur_predictor.linear.0.weight"] = (rs.standard_normal((1, H)) * (0.8 / math.sqrt(H))).astype(F32)
    sd["dur_predictor.linear.0.bias"] = np.array([2.0], dtype=F32)
    if hp["use_pitch_embed"]:
        # rows vary smoothly with the pitch bin (a random walk), as a trained table does: a bin that flips by one under
        # bf16 arithmetic then moves the condition by a few percent of a row, not by a whole independent row
        walk = np.cumsum(rs.standard_normal((300, H)) * (0.15 * H ** -0.5), axis=0) + rs.standard_normal((1, H)) * H ** -0.5
        walk[0] = 0.0
        sd["pitch_embed.weight"] = walk.astype(F32)
        for i in range(hp["pitch_predictor_layers"]):
            conv(f"pitch_predictor.conv.{i}.0", H, H, hp["predictor_kernel"], gain=1.4)
            ln(f"pitch_predictor.conv.{i}.2", H)
        sd["pitch_predictor.linear.weight"] = (rs.standard_normal((2, H)) * (0.7 / math.sqrt(H))).astype(F32)
        sd["pitch_predictor.linear.bias"] = np.array([7.5, 0.0], dtype=F32)
    return sd


def pad_edit_batch(batch: dict, item: int, n_tokens: int, frames_per_phone: int = 8) -> dict:
    """Turn the tail of one item into padding (txt token 0, mel2ph 0) as collated ragged batches have it
    (tasks/speech_editing/dataset_utils.py:148-170): the last `n_tokens` phones and their frames."""
    out = {k: v.copy() for k, v in batch.items()}
    Tt = out["txt_tokens"].shape[1]
    keep = Tt - n_tokens
    out["txt_tokens"][item, keep:] = 0
    pad = out["mel2ph"][item] > keep
    out["mel2ph"][item, pad] = 0
    for k in ("time_mel_masks", "f0", "uv"):
        out[k][item, pad] = 0
    out["ref_mels"][item, pad] = 0
    return out


def campnet_state_dict(seed: int = 1234, vocab: int = 80, hidden: int = 192, n_mels: int = 80, enc_layers: int = 3, dec_layers: int = 6,
                       ffn_kernel: int = 9, fine_blocks: int = 5) -> Dict[str, np.ndarray]:
    """Random CampNet state_dict (modules/speech_editing/campnet/campnet.py:14-38) with the reference's keys/shapes for every
    tensor the forward uses: encoder (embed_tokens + 3 EncSALayers + layer_norm), mel_encoder, decoder_coarse (6 DecSALayers),
    decoder_fine (ConvBlocks), mel_out_coarse / mel_out_fine, mask_emb.  Unused entries (encoder.pre_net.*, mel_out.*, the
    _float_tensor buffers) are not generated."""
    rs = np.random.RandomState(seed + 97)
    C = hidden
    sd: Dict[str, np.ndarray] = {}

    def mat(name, co, ci, gain=1.0, bias=True):
        sd[name + ".weight"] = (rs.standard_normal((co, ci)) * (gain / math.sqrt(ci))).astype(F32)
        if bias:
            sd[name + ".bias"] = (rs.standard_normal((co,)) * 0.1).astype(F32)

    def conv(name, co, ci, k, gain=1.0):
        sd[name + ".weight"] = (rs.standard_normal((co, ci, k)) * (gain / math.sqrt(ci * k))).astype(F32)
        sd[name + ".bias"] = (rs.standard_normal((co,)) * 0.1).astype(F32)

    def ln(name):
        sd[name + ".weight"] = (1.0 + 0.1 * rs.standard_normal((C,))).astype(F32)
        sd[name + ".bias"] = (0.1 * rs.standard_normal((C,))).astype(F32)

    def attn(name):
        sd[name + ".in_proj_weight"] = (rs.standard_normal((3 * C, C)) * (1.5 / math.sqrt(C))).astype(F32)
        sd[name + ".out_proj.weight"] = (rs.standard_normal((C, C)) / math.sqrt(C)).astype(F32)

    sd["mask_emb"] = (rs.standard_normal((1, 1, n_mels)) * 0.5).astype(F32)
    w = (rs.standard_normal((vocab, C)) * C ** -0.5).astype(F32)
    w[0] = 0.0
    sd["encoder.embed_tokens.weight"] = w
    for i in range(enc_layers):
        p = f"encoder.layers.{i}.op."
        ln(p + "layer_norm1"); attn(p + "self_attn"); ln(p + "layer_norm2")
        conv(p + "ffn.ffn_1", 4 * C, C, ffn_kernel, gain=1.5)
        mat(p + "ffn.ffn_2", C, 4 * C)
    ln("encoder.layer_norm")
    for name, (n, c) in (("mel_encoder.encoder.0", (C, n_mels)), ("mel_encoder.encoder.2", (C, C)), ("mel_encoder.fc_out", (C, C))):
        mat(name, n, c)
    sd["decoder_coarse.pos_embed_alpha"] = np.array([0.8], dtype=F32)
    for i in range(dec_layers):
        p = f"decoder_coarse.layers.{i}.op."
        ln(p + "layer_norm1"); attn(p + "self_attn"); ln(p + "layer_norm2"); attn(p + "encoder_attn"); ln(p + "layer_norm3")
        conv(p + "ffn.ffn_1.1", 4 * C, C, ffn_kernel, gain=1.5)           # 'LEFT' padding: Sequential(ConstantPad1d, Conv1d)
        mat(p + "ffn.ffn_2", C, 4 * C)
    ln("decoder_coarse.layer_norm")
    for i in range(fine_blocks):
        for j in range(2):
            p = f"decoder_fine.res_blocks.{i}.blocks.{j}."
            ln(p + "0")
            conv(p + "1", 2 * C, C, 5, gain=1.6)
            conv(p + "4", C, 2 * C, 1)
    ln("decoder_fine.last_norm")
    conv("decoder_fine.post_net1", C, C, 3)
    mat("mel_out_coarse", n_mels, C, bias=False)
    mat("mel_out_fine", n_mels, C, bias=False)
    return sd


def synthetic_campnet_batch(seed: int, B: int, T: int, vocab: int = 80, frames_per_phone: int = 8, pad_items=()):
    """(txt_tokens, mels, time_mel_masks) of the CampNet mask-predict forward (tasks/speech_editing/campnet.py:52-80);
    `pad_items` = [(item, n_tokens)] turns tails into padding (token 0 / all-zero mel frames)."""
    b = synthetic_edit_batch(seed, B, T, vocab=vocab, frames_per_phone=frames_per_phone)
    for item, n in pad_items:
        b = pad_edit_batch(b, item, n, frames_per_phone)
    return dict(txt_tokens=b["txt_tokens"], mels=b["ref_mels"], time_mel_masks=b["time_mel_masks"][:, :, None].copy())


def synthetic_edit_item(seed: int, n_words: int = 9, n_mels: int = 80, vocab: int = 80, edit_span=(3, 4), new_span_phones=(2, 3, 1)):
    """One utterance as inference/tts/spec_denoiser.py::preprocess_input produces it (:151-196), synthetic: words of 1-4
    phones, phones of 2-7 frames (1-based ph2word / mel2ph / mel2word), an edit that replaces the words `edit_span` (1-based,
    inclusive) by len(new_span_phones) new words with that many phones each."""
    rs = np.random.RandomState(seed + 71)
    phones_per_word = rs.randint(1, 5, size=n_words)
    ph2word = np.repeat(np.arange(1, n_words + 1), phones_per_word).astype(np.int64)
    Tp = len(ph2word)
    dur = rs.randint(2, 8, size=Tp).astype(np.int64)
    mel2ph = np.repeat(np.arange(1, Tp + 1), dur).astype(np.int64)
    mel2word = ph2word[mel2ph - 1]
    T = len(mel2ph)
    w0, w1 = edit_span
    head_words, tail_words = w0 - 1, n_words - w1
    new_ppw = np.concatenate([phones_per_word[:head_words], np.asarray(new_span_phones), phones_per_word[w1:]])
    edited_ph2word = np.repeat(np.arange(1, len(new_ppw) + 1), new_ppw).astype(np.int64)
    c0, c1 = w0, w0 + len(new_span_phones) - 1
    ph_token = rs.randint(3, vocab, size=Tp).astype(np.int64)
    n_head_ph, n_tail_ph = int(phones_per_word[:head_words].sum()), int(phones_per_word[w1:].sum())
    edited_ph_token = np.concatenate([ph_token[:n_head_ph], rs.randint(3, vocab, size=int(np.sum(new_span_phones))),
                                      ph_token[Tp - n_tail_ph:] if n_tail_ph else ph_token[:0]]).astype(np.int64)
    mel = np.clip(rs.standard_normal((T, n_mels)) * 1.5 - 3.0, -6.0, 1.5).astype(F32)
    f0 = rs.uniform(6.5, 8.5, T).astype(F32)
    uv = (rs.uniform(size=T) < 0.3).astype(F32)
    spk = (rs.standard_normal(256) / 16).astype(F32)
    assert len(edited_ph_token) == len(edited_ph2word) and tail_words >= 0
    return dict(ph2word=ph2word, edited_ph2word=edited_ph2word, mel2ph=mel2ph, mel2word=mel2word, dur=dur, ph_token=ph_token,
                edited_ph_token=edited_ph_token, words_region=[(w0, w1)], edited_words_region=[(c0, c1)], mel=mel, f0=f0, uv=uv, spk_embed=spk)
