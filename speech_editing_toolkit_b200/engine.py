"""Thin Python owners of the C-ABI handles.  PyTorch is plumbing here: it allocates device memory and
provides the CUDA stream; all arithmetic of the hot path happens inside libfse_b200.so."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import FseError, MODES, check


def _as_f32_numpy(v) -> np.ndarray:
    if isinstance(v, torch.Tensor):
        v = v.detach().to("cpu", torch.float32).numpy()
    return np.ascontiguousarray(v, dtype=np.float32)


def _tensor_table(sd: Dict[str, object], skip_suffixes=()):
    keep = []     # keep numpy arrays alive while the C call runs
    names = [k for k in sd if not any(k.endswith(s) for s in skip_suffixes)]
    arr = (_lib.Tensor * len(names))()
    for i, k in enumerate(names):
        a = _as_f32_numpy(sd[k])
        keep.append(a)
        arr[i].name = k.encode()
        arr[i].data = a.ctypes.data_as(C.POINTER(C.c_float))
        arr[i].numel = a.size
    return arr, len(names), keep


class _Range:
    """NVTX range around a C-ABI call when FSE_NVTX=1 (the reference's only hook of this kind is Timer('hifigan') around the
    vocoder, tasks/tts/vocoder_infer/hifigan.py:28); a no-op otherwise."""
    enabled = os.environ.get("FSE_NVTX", "0") == "1"

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _Range.enabled:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if _Range.enabled:
            torch.cuda.nvtx.range_pop()
        return False


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise FseError("tensor must live on a CUDA device; this path has no CPU implementation")


class _Workspace:
    """Caller-owned (torch-allocated) device workspace, grown on demand, 1024-byte aligned."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device):
        if self.buf is None or self.buf.numel() < nbytes + 1024 or self.buf.device != device:
            self.buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        base = self.buf.data_ptr()
        aligned = (base + 1023) // 1024 * 1024
        return C.c_void_p(aligned), nbytes


class Denoiser:
    """Handle of the DiffNet denoiser + sampling loop (fse_denoiser_* / fse_denoise_step / fse_sample)."""

    def __init__(self, n_mels=80, hidden=192, channels=256, layers=20, dilation_cycle_length=1, mode="tc_bf16"):
        self.cfg = _lib.DenoiserConfig(n_mels, hidden, channels, layers, dilation_cycle_length, MODES[mode])
        self.mode = mode
        self._h = C.c_void_p()
        check(_lib.lib().fse_denoiser_create(C.byref(self.cfg), C.byref(self._h)))
        self._ws = _Workspace()
        self.timesteps = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _lib.lib().fse_denoiser_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def load_state_dict(self, sd: Dict[str, object]):
        arr, n, keep = _tensor_table(sd)
        check(_lib.lib().fse_denoiser_load_weights(self._h, arr, n))
        del keep

    def set_schedule(self, coef1, coef2, logvar_clipped):
        c1, c2, lv = (_as_f32_numpy(a) for a in (coef1, coef2, logvar_clipped))
        assert c1.shape == c2.shape == lv.shape and c1.ndim == 1
        self.timesteps = c1.shape[0] - 1
        f = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        check(_lib.lib().fse_denoiser_set_schedule(self._h, self.timesteps, f(c1), f(c2), f(lv)))

    def _workspace(self, B, T, device):
        nbytes = _lib.lib().fse_denoiser_workspace_bytes(self._h, B, T)
        return self._ws.get(nbytes, device)

    @property
    def last_launches(self) -> int:
        return int(_lib.lib().fse_denoiser_last_launches(self._h))

    KINDS = ["input_proj", "gate_gemm", "res_gemm", "skip_proj", "out_proj_posterior"]

    def profile(self, enable: bool):
        check(_lib.lib().fse_denoiser_profile(self._h, int(enable)))

    def profile_read(self):
        ms = (C.c_double * 8)(); cnt = (C.c_int64 * 8)()
        check(_lib.lib().fse_denoiser_profile_read(self._h, ms, cnt))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(self.KINDS)}

    def denoise_step(self, x_t: torch.Tensor, cond: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """x_t[B,M,T] fp32, cond[B,T,H] fp32 (physical layout), t[B] int64 -> x0[B,M,T]."""
        _need_cuda(x_t, cond, t)
        B, M, T = x_t.shape
        x_t = x_t.contiguous().float(); cond = cond.contiguous().float(); t = t.contiguous().long()
        assert cond.shape == (B, T, self.cfg.hidden), f"cond must be [B,T,H], got {tuple(cond.shape)}"
        x0 = torch.empty_like(x_t)
        ws, nbytes = self._workspace(B, T, x_t.device)
        check(_lib.lib().fse_denoise_step(self._h, _ptr(x_t), _ptr(cond), _ptr(t), _ptr(x0), B, T, ws, nbytes, _stream()))
        return x0

    def posterior_step(self, x0, x_t, t, noise=None, seed=0, step=0) -> torch.Tensor:
        _need_cuda(x0, x_t, t, noise)
        B, M, T = x_t.shape
        x0 = x0.contiguous().float(); x_t = x_t.contiguous().float(); t = t.contiguous().long()
        if noise is not None:
            noise = noise.contiguous().float()
        out = torch.empty_like(x_t)
        check(_lib.lib().fse_posterior_step(self._h, _ptr(x0), _ptr(x_t), _ptr(t), _ptr(noise), seed, step, _ptr(out), B, T, _stream()))
        return out

    def sample(self, cond: torch.Tensor, noise: Optional[torch.Tensor] = None, seed: int = 0,
               ref_mel: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None, trace: bool = False):
        """cond[B,T,H] -> mel_out[B,T,M].  noise: None (Philox) or [(S+1),B,M,T]."""
        _need_cuda(cond, noise, ref_mel, mask)
        B, T, H = cond.shape
        M, S = self.cfg.n_mels, self.timesteps
        cond = cond.contiguous().float()
        if noise is not None:
            noise = noise.contiguous().float()
            assert noise.shape == (S + 1, B, M, T), f"noise must be [S+1,B,M,T], got {tuple(noise.shape)}"
        if ref_mel is not None:
            ref_mel = ref_mel.contiguous().float(); mask = mask.reshape(B, T).contiguous().float()
        mel = torch.empty(B, T, M, dtype=torch.float32, device=cond.device)
        xs = torch.empty(S, B, M, T, dtype=torch.float32, device=cond.device) if trace else None
        ws, nbytes = self._workspace(B, T, cond.device)
        with _Range("fse_sample"):
            check(_lib.lib().fse_sample(self._h, _ptr(cond), _ptr(noise), seed, _ptr(ref_mel), _ptr(mask), _ptr(mel), _ptr(xs),
                                        B, T, ws, nbytes, _stream()))
        return (mel, xs) if trace else mel

    def sample_host(self, cond: np.ndarray, noise=None, seed: int = 0, ref_mel=None, mask=None) -> np.ndarray:
        """HOST numpy in / out through fse_sample_host (H2D + D2H inside the call)."""
        cond = np.ascontiguousarray(cond, dtype=np.float32)
        B, T, H = cond.shape
        out = np.empty((B, T, self.cfg.n_mels), dtype=np.float32)
        p = lambda a: None if a is None else C.c_void_p(np.ascontiguousarray(a, dtype=np.float32).ctypes.data)
        keep = [np.ascontiguousarray(a, dtype=np.float32) if a is not None else None for a in (noise, ref_mel, mask)]
        q = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
        check(_lib.lib().fse_sample_host(self._h, C.c_void_p(cond.ctypes.data), q(keep[0]), seed, q(keep[1]), q(keep[2]),
                                         C.c_void_p(out.ctypes.data), B, T))
        return out


HIFIGAN_V1 = dict(
    upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
    resblock="1", resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
)


class Vocoder:
    """Handle of the HiFi-GAN generator (fse_vocoder_*)."""

    def __init__(self, config: dict = None, n_mels: int = 80, mode="tc_bf16"):
        config = dict(HIFIGAN_V1 if config is None else config)
        rb = str(config.get("resblock", "1"))
        if rb not in ("1", "2"):
            raise FseError(f"config['resblock'] must be '1' or '2' (hifigan.py:109), got {rb!r}")
        cfg = _lib.VocoderConfig()
        cfg.n_mels = n_mels
        cfg.upsample_initial_channel = config["upsample_initial_channel"]
        rates, ks = config["upsample_rates"], config["upsample_kernel_sizes"]
        cfg.num_upsamples = len(rates)
        for i, (u, k) in enumerate(zip(rates, ks)):
            cfg.upsample_rates[i] = u
            cfg.upsample_kernel_sizes[i] = k
        rks, rds = config["resblock_kernel_sizes"], config["resblock_dilation_sizes"]
        cfg.num_kernels = len(rks)
        for j, (k, ds) in enumerate(zip(rks, rds)):
            cfg.resblock_kernel_sizes[j] = k
            assert len(ds) == (3 if rb == "1" else 2), "ResBlock1 has three dilations per block, ResBlock2 two"
            for m, d in enumerate(ds):
                cfg.resblock_dilations[j][m] = d
        cfg.mode = MODES[mode]
        cfg.resblock = int(rb)
        self.cfg, self.config, self.mode = cfg, config, mode
        self.hop = int(np.prod(rates))
        self._h = C.c_void_p()
        check(_lib.lib().fse_vocoder_create(C.byref(cfg), C.byref(self._h)))
        self._ws = _Workspace()

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _lib.lib().fse_vocoder_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def load_state_dict(self, sd: Dict[str, object]):
        arr, n, keep = _tensor_table(sd)
        check(_lib.lib().fse_vocoder_load_weights(self._h, arr, n))
        del keep

    @property
    def last_launches(self) -> int:
        return int(_lib.lib().fse_vocoder_last_launches(self._h))

    KINDS = ["conv_pre", "upsample", "res_conv1", "res_conv2", "conv_post"]

    def profile(self, enable: bool):
        check(_lib.lib().fse_vocoder_profile(self._h, int(enable)))

    def profile_read(self):
        ms = (C.c_double * 8)(); cnt = (C.c_int64 * 8)()
        check(_lib.lib().fse_vocoder_profile_read(self._h, ms, cnt))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(self.KINDS)}

    def forward(self, mel: torch.Tensor) -> torch.Tensor:
        """mel[B,T,M] fp32 cuda -> wav[B,T*hop]."""
        _need_cuda(mel)
        B, T, M = mel.shape
        mel = mel.contiguous().float()
        wav = torch.empty(B, T * self.hop, dtype=torch.float32, device=mel.device)
        nbytes = _lib.lib().fse_vocoder_workspace_bytes(self._h, B, T)
        ws, nbytes = self._ws.get(nbytes, mel.device)
        with _Range("fse_vocoder_forward"):
            check(_lib.lib().fse_vocoder_forward(self._h, _ptr(mel), _ptr(wav), B, T, ws, nbytes, _stream()))
        return wav

    def forward_host(self, mel: np.ndarray) -> np.ndarray:
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        B, T, M = mel.shape
        wav = np.empty((B, T * self.hop), dtype=np.float32)
        check(_lib.lib().fse_vocoder_forward_host(self._h, C.c_void_p(mel.ctypes.data), C.c_void_p(wav.ctypes.data), B, T))
        return wav



class MelEncoderKernel:
    """Handle of the context-mel encoder (fse_mel_encoder_*; mel_encoder.py:3-19)."""

    def __init__(self, n_mels: int = 80, hidden: int = 192, mode="tc_bf16"):
        cfg = _lib.MelEncoderConfig()
        cfg.n_mels, cfg.hidden, cfg.mode = n_mels, hidden, MODES[mode]
        self.cfg, self.mode, self.hidden = cfg, mode, hidden
        self._h = C.c_void_p()
        check(_lib.lib().fse_mel_encoder_create(C.byref(cfg), C.byref(self._h)))
        self._ws = _Workspace()

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _lib.lib().fse_mel_encoder_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def load_state_dict(self, sd: Dict[str, object]):
        arr, n, keep = _tensor_table(sd)
        check(_lib.lib().fse_mel_encoder_load_weights(self._h, arr, n))
        del keep

    @property
    def last_launches(self) -> int:
        return int(_lib.lib().fse_mel_encoder_last_launches(self._h))

    def forward(self, x: torch.Tensor, add: Optional[torch.Tensor] = None, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x[B,T,M] fp32 cuda (= ref_mels * (1 - mask)) -> [B,T,hidden]; with add / scale: add + MelEncoder(x) * scale[B,T]."""
        _need_cuda(x, add, scale)
        B, T, M = x.shape
        x = x.contiguous().float()
        if add is not None:
            add = add.contiguous().float()
            assert add.shape == (B, T, self.hidden)
        if scale is not None:
            scale = scale.reshape(B, T).contiguous().float()
        out = torch.empty(B, T, self.hidden, dtype=torch.float32, device=x.device)
        nbytes = _lib.lib().fse_mel_encoder_workspace_bytes(self._h, B, T)
        ws, nbytes = self._ws.get(nbytes, x.device)
        check(_lib.lib().fse_mel_encoder_forward(self._h, _ptr(x), _ptr(add), _ptr(scale), _ptr(out), B, T, ws, nbytes, _stream()))
        return out


class CondEncoderKernel:
    """Handle of the condition encoder (fse_cond_*; FastSpeech.forward(skip_decoder=True), fs.py:83-189).  One method per
    C-ABI entry point; every tensor stays on the device and integer tensors are int64 as in the reference."""

    def __init__(self, vocab: int, hidden: int = 192, enc_dilations=(1, 1, 1, 1), enc_kernel_size: int = 5, layers_in_block: int = 2,
                 enc_post_net_kernel: int = 3, dur_predictor_layers: int = 3, dur_predictor_kernel: int = 5,
                 pitch_predictor_layers: int = 5, predictor_kernel: int = 5, use_pitch_embed: bool = True, use_uv: bool = True,
                 spk_embed_dim: int = 256, mode="tc_bf16"):
        cfg = _lib.CondEncoderConfig()
        cfg.hidden, cfg.vocab, cfg.enc_layers = hidden, vocab, len(enc_dilations)
        if len(enc_dilations) > 8:
            raise FseError("at most 8 encoder blocks (enc_dilations)")
        for i, d in enumerate(enc_dilations):
            cfg.enc_dilations[i] = int(d)
        cfg.enc_kernel_size, cfg.layers_in_block, cfg.enc_post_net_kernel = enc_kernel_size, layers_in_block, enc_post_net_kernel
        cfg.dur_predictor_layers, cfg.dur_predictor_kernel = dur_predictor_layers, dur_predictor_kernel
        cfg.pitch_predictor_layers, cfg.predictor_kernel = pitch_predictor_layers, predictor_kernel
        cfg.use_pitch_embed, cfg.use_uv, cfg.spk_embed_dim, cfg.mode = int(use_pitch_embed), int(use_uv), spk_embed_dim, MODES[mode]
        self.cfg, self.mode, self.hidden, self.use_pitch_embed = cfg, mode, hidden, bool(use_pitch_embed)
        self._h = C.c_void_p()
        check(_lib.lib().fse_cond_encoder_create(C.byref(cfg), C.byref(self._h)))
        self._ws = _Workspace()
        self.launches = 0          # kernels enqueued since the last reset_launches()

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _lib.lib().fse_cond_encoder_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def load_state_dict(self, sd: Dict[str, object]):
        """`fs.*` entries of a reference checkpoint without the prefix; decoder.* / mel_out.* (unused with skip_decoder) are skipped."""
        sd = {k: v for k, v in sd.items() if not k.startswith(("decoder.", "mel_out."))}
        arr, n, keep = _tensor_table(sd)
        check(_lib.lib().fse_cond_encoder_load_weights(self._h, arr, n))
        del keep

    def reset_launches(self):
        self.launches = 0

    def _done(self):
        self.launches += int(_lib.lib().fse_cond_encoder_last_launches(self._h))

    def _workspace(self, B, Tt, T, device):
        nbytes = _lib.lib().fse_cond_encoder_workspace_bytes(self._h, B, Tt, T)
        return self._ws.get(nbytes, device)

    @staticmethod
    def _i64(t):
        return t.contiguous().long()

    @staticmethod
    def _f32(t):
        return None if t is None else t.contiguous().float()

    def text_encoder(self, txt: torch.Tensor) -> torch.Tensor:
        """txt[B,Tt] int64 -> encoder_out[B,Tt,H]   (TextConvEncoder.forward, conv.py:130-139)"""
        _need_cuda(txt)
        B, Tt = txt.shape
        txt = self._i64(txt)
        out = torch.empty(B, Tt, self.hidden, dtype=torch.float32, device=txt.device)
        ws, nbytes = self._workspace(B, Tt, 0, txt.device)
        check(_lib.lib().fse_cond_text_encoder(self._h, _ptr(txt), _ptr(out), B, Tt, ws, nbytes, _stream()))
        self._done()
        return out

    def style_embed(self, spk_embed: torch.Tensor) -> torch.Tensor:
        """spk_embed[B,D] -> [B,H]   (fs.py:117-118)"""
        _need_cuda(spk_embed)
        spk_embed = self._f32(spk_embed)
        B = spk_embed.shape[0]
        out = torch.empty(B, self.hidden, dtype=torch.float32, device=spk_embed.device)
        check(_lib.lib().fse_cond_style_embed(self._h, _ptr(spk_embed), _ptr(out), B, _stream()))
        self._done()
        return out

    def dur_input(self, encoder_out, style, txt) -> torch.Tensor:
        """(encoder_out + style) * (txt > 0)   (fs.py:90)"""
        _need_cuda(encoder_out, style, txt)
        B, Tt, _ = encoder_out.shape
        encoder_out, style, txt = self._f32(encoder_out), self._f32(style), self._i64(txt)
        out = torch.empty_like(encoder_out)
        check(_lib.lib().fse_cond_dur_input(self._h, _ptr(encoder_out), _ptr(style), _ptr(txt), _ptr(out), B, Tt, _stream()))
        self._done()
        return out

    def masked_dur(self, mel2ph, mask, txt) -> torch.Tensor:
        """mel2token_to_dur(mel2ph * (1 - mask).long(), Tt) * (txt != 0)   (fs.py:136-138), int64 [B,Tt]"""
        _need_cuda(mel2ph, mask, txt)
        B, T = mel2ph.shape
        Tt = txt.shape[1]
        mel2ph, txt = self._i64(mel2ph), self._i64(txt)
        mask = None if mask is None else self._f32(mask.reshape(B, T))
        out = torch.empty(B, Tt, dtype=torch.int64, device=txt.device)
        check(_lib.lib().fse_cond_masked_dur(self._h, _ptr(mel2ph), _ptr(mask), _ptr(txt), _ptr(out), B, T, Tt, _stream()))
        self._done()
        return out

    def duration(self, dur_inp, masked_dur, txt) -> torch.Tensor:
        """DurationPredictor(dur_inp + dur_embed(masked_dur), txt == 0) -> dur[B,Tt]   (fs.py:139-148)"""
        _need_cuda(dur_inp, masked_dur, txt)
        B, Tt, _ = dur_inp.shape
        dur_inp, masked_dur, txt = self._f32(dur_inp), self._i64(masked_dur), self._i64(txt)
        out = torch.empty(B, Tt, dtype=torch.float32, device=txt.device)
        ws, nbytes = self._workspace(B, Tt, 0, txt.device)
        check(_lib.lib().fse_cond_duration(self._h, _ptr(dur_inp), _ptr(masked_dur), _ptr(txt), _ptr(out), B, Tt, ws, nbytes, _stream()))
        self._done()
        return out

    def length_regulate(self, dur, txt=None) -> torch.Tensor:
        """LengthRegulator.forward (nar_tts_modules.py:42-72): dur[B,Tt] -> mel2ph[B, max total] int64.  One host sync for the
        data-dependent output length, as in the reference (`dur.sum(-1).max()` sizes an arange)."""
        _need_cuda(dur, txt)
        B, Tt = dur.shape
        dur = self._f32(dur)
        txt = None if txt is None else self._i64(txt)
        cumsum = torch.empty(B, Tt, dtype=torch.int64, device=dur.device)
        totals = torch.empty(B, dtype=torch.int64, device=dur.device)
        check(_lib.lib().fse_cond_length_cumsum(self._h, _ptr(dur), _ptr(txt), _ptr(cumsum), _ptr(totals), B, Tt, _stream()))
        self._done()
        tmax = int(totals.max().item())
        out = torch.zeros(B, max(tmax, 0), dtype=torch.int64, device=dur.device)
        if tmax > 0:
            check(_lib.lib().fse_cond_length_fill(self._h, _ptr(cumsum), _ptr(out), B, Tt, tmax, _stream()))
            self._done()
        return out

    def frames(self, encoder_out, style, mel2ph, mask, f0, uv, use_pred_pitch: bool):
        """fs.py:93-102 -> dict(decoder_inp[B,T,H], and with use_pitch_embed pitch_pred[B,T,2], f0_denorm, f0_denorm_pred, pitch)"""
        _need_cuda(encoder_out, style, mel2ph, mask, f0, uv)
        B, Tt, _ = encoder_out.shape
        T = mel2ph.shape[1]
        dev = encoder_out.device
        encoder_out, style, mel2ph = self._f32(encoder_out), self._f32(style), self._i64(mel2ph)
        mask = None if mask is None else self._f32(mask.reshape(B, T))
        ret = {"decoder_inp": torch.empty(B, T, self.hidden, dtype=torch.float32, device=dev)}
        pp = fd = fdp = pitch = None
        if self.use_pitch_embed:
            if f0 is None or uv is None:
                raise FseError("use_pitch_embed: f0 and uv are required")
            f0, uv = self._f32(f0), self._f32(uv)
            ret["pitch_pred"] = pp = torch.empty(B, T, 2, dtype=torch.float32, device=dev)
            ret["f0_denorm"] = fd = torch.empty(B, T, dtype=torch.float32, device=dev)
            ret["f0_denorm_pred"] = fdp = torch.empty(B, T, dtype=torch.float32, device=dev)
            ret["pitch"] = pitch = torch.empty(B, T, dtype=torch.int64, device=dev)
        else:
            f0 = uv = None
        ws, nbytes = self._workspace(B, Tt, T, dev)
        check(_lib.lib().fse_cond_frames(self._h, _ptr(encoder_out), _ptr(style), _ptr(mel2ph), _ptr(mask), _ptr(f0), _ptr(uv),
                                         int(bool(use_pred_pitch)), _ptr(ret["decoder_inp"]), _ptr(pp), _ptr(fd), _ptr(fdp), _ptr(pitch),
                                         B, Tt, T, ws, nbytes, _stream()))
        self._done()
        return ret


class CampNetKernel:
    """Handle of the CampNet mask-predict forward (fse_campnet_*; modules/speech_editing/campnet/campnet.py:14-69)."""

    def __init__(self, vocab: int, hidden: int = 192, n_mels: int = 80, enc_layers: int = 3, dec_layers: int = 6, heads: int = 2,
                 ffn_kernel: int = 9, fine_blocks: int = 5, fine_kernel: int = 5, mode="tc_bf16"):
        cfg = _lib.CampNetConfig()
        cfg.hidden, cfg.vocab, cfg.n_mels, cfg.enc_layers, cfg.dec_layers = hidden, vocab, n_mels, enc_layers, dec_layers
        cfg.heads, cfg.ffn_kernel, cfg.fine_blocks, cfg.fine_kernel, cfg.mode = heads, ffn_kernel, fine_blocks, fine_kernel, MODES[mode]
        self.cfg, self.mode, self.hidden, self.n_mels = cfg, mode, hidden, n_mels
        self._h = C.c_void_p()
        check(_lib.lib().fse_campnet_create(C.byref(cfg), C.byref(self._h)))
        self._ws = _Workspace()

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _lib.lib().fse_campnet_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def load_state_dict(self, sd: Dict[str, object]):
        """A reference CampNet state_dict; entries the forward never reads (encoder.pre_net.*, mel_out.*, *._float_tensor) are skipped."""
        sd = {k: v for k, v in sd.items() if not (k.startswith(("encoder.pre_net.", "mel_out.")) or k.endswith("_float_tensor"))}
        arr, n, keep = _tensor_table(sd)
        check(_lib.lib().fse_campnet_load_weights(self._h, arr, n))
        del keep

    @property
    def last_launches(self) -> int:
        return int(_lib.lib().fse_campnet_last_launches(self._h))

    def forward(self, txt: torch.Tensor, mels: torch.Tensor, time_mel_masks: torch.Tensor, need_attn: bool = True, need_encoder_out: bool = False):
        """txt[B,Tt] int64, mels[B,T,M], time_mel_masks[B,T(,1)] 0/1 -> dict(mel_out_coarse, mel_out_fine [B,T,M], attn [B,T,Tt])"""
        _need_cuda(txt, mels, time_mel_masks)
        B, Tt = txt.shape
        T = mels.shape[1]
        dev = mels.device
        txt, mels = txt.contiguous().long(), mels.contiguous().float()
        mask = time_mel_masks.reshape(B, T).contiguous().float()
        coarse = torch.empty(B, T, self.n_mels, dtype=torch.float32, device=dev)
        fine = torch.empty_like(coarse)
        attn = torch.empty(B, T, Tt, dtype=torch.float32, device=dev) if need_attn else None
        enc = torch.empty(B, Tt, self.hidden, dtype=torch.float32, device=dev) if need_encoder_out else None
        nbytes = _lib.lib().fse_campnet_workspace_bytes(self._h, B, Tt, T)
        ws, nbytes = self._ws.get(nbytes, dev)
        with _Range("fse_campnet_forward"):
            check(_lib.lib().fse_campnet_forward(self._h, _ptr(txt), _ptr(mels), _ptr(mask), _ptr(coarse), _ptr(fine), _ptr(attn), _ptr(enc),
                                                 B, Tt, T, ws, nbytes, _stream()))
        ret = {"mel_out_coarse": coarse, "mel_out_fine": fine, "attn": attn}
        if need_encoder_out:
            ret["encoder_out"] = enc
        return ret


# ---------------------------------------------------------------------------------------------- region surgery (stateless)
def _i64(t):
    return None if t is None else t.contiguous().long()


def edit_prepare(mel2ph, mel2word, ph2word, dur, regions, n_edited_phones: int, T_len=None, Tp_len=None, Tpe_len=None):
    """inference/tts/spec_denoiser.py:88-97 for a padded batch -> (masked_dur [B,Tpe], masked_mel2ph [B,T], time_mel_masks_orig [B,T]).
    regions [B,4] int64 = (w0, w1, c0, c1)."""
    _need_cuda(mel2ph, mel2word, ph2word, dur, regions, T_len, Tp_len, Tpe_len)
    B, T = mel2ph.shape
    Tp, Tpe = ph2word.shape[1], int(n_edited_phones)
    mel2ph, mel2word, ph2word, dur, regions = _i64(mel2ph), _i64(mel2word), _i64(ph2word), _i64(dur), _i64(regions)
    T_len, Tp_len, Tpe_len = _i64(T_len), _i64(Tp_len), _i64(Tpe_len)
    dev = mel2ph.device
    masked_dur = torch.empty(B, Tpe, dtype=torch.int64, device=dev)
    masked_mel2ph = torch.empty(B, T, dtype=torch.int64, device=dev)
    mask_orig = torch.empty(B, T, dtype=torch.float32, device=dev)
    check(_lib.lib().fse_edit_prepare(_ptr(mel2ph), _ptr(mel2word), _ptr(T_len), _ptr(ph2word), _ptr(dur), _ptr(Tp_len), _ptr(Tpe_len), _ptr(regions),
                                      _ptr(masked_dur), _ptr(masked_mel2ph), _ptr(mask_orig), B, T, Tp, Tpe, _stream()))
    return masked_dur, masked_mel2ph, mask_orig


def edit_assemble(mel2ph, mel2word, edited_ph2word, edited_mel2ph, regions, mel, f0, uv, T_len=None, Tpe_len=None, Te_len=None):
    """:99-131 for a padded batch -> dict(mel2ph [B,Tn] int64, ref_mels [B,Tn,M], f0, uv, time_mel_masks [B,Tn], plan [B,8] (host)).
    One host sync reads the per-item output lengths (the reference's own slicing implies the same)."""
    _need_cuda(mel2ph, mel2word, edited_ph2word, edited_mel2ph, regions, mel, f0, uv, T_len, Tpe_len, Te_len)
    B, T = mel2ph.shape
    Tpe, Te, M = edited_ph2word.shape[1], edited_mel2ph.shape[1], mel.shape[2]
    mel2ph, mel2word, edited_ph2word, edited_mel2ph, regions = _i64(mel2ph), _i64(mel2word), _i64(edited_ph2word), _i64(edited_mel2ph), _i64(regions)
    T_len, Tpe_len, Te_len = _i64(T_len), _i64(Tpe_len), _i64(Te_len)
    mel = mel.contiguous().float()
    f0 = None if f0 is None else f0.contiguous().float()
    uv = None if uv is None else uv.contiguous().float()
    dev = mel2ph.device
    sel_edit = torch.empty(B, Te, dtype=torch.int32, device=dev)
    sel_tail = torch.empty(B, T, dtype=torch.int32, device=dev)
    plan = torch.empty(B, 8, dtype=torch.int64, device=dev)
    check(_lib.lib().fse_edit_plan(_ptr(mel2ph), _ptr(mel2word), _ptr(T_len), _ptr(edited_ph2word), _ptr(Tpe_len), _ptr(regions), _ptr(edited_mel2ph),
                                   _ptr(Te_len), _ptr(sel_edit), _ptr(sel_tail), _ptr(plan), B, T, Tpe, Te, _stream()))
    plan_h = plan.cpu()
    Tn = int(plan_h[:, 0].max().item())
    if Tn <= 0:
        raise FseError("region surgery produced an empty sequence")
    out = dict(mel2ph=torch.empty(B, Tn, dtype=torch.int64, device=dev), ref_mels=torch.empty(B, Tn, M, dtype=torch.float32, device=dev),
               f0=torch.empty(B, Tn, dtype=torch.float32, device=dev), uv=torch.empty(B, Tn, dtype=torch.float32, device=dev),
               time_mel_masks=torch.empty(B, Tn, dtype=torch.float32, device=dev), plan=plan_h)
    check(_lib.lib().fse_edit_assemble(_ptr(mel2ph), _ptr(T_len), _ptr(regions), _ptr(plan), _ptr(edited_mel2ph), _ptr(sel_edit), _ptr(sel_tail),
                                       _ptr(mel), _ptr(f0), _ptr(uv), _ptr(out["mel2ph"]), _ptr(out["ref_mels"]), _ptr(out["f0"]), _ptr(out["uv"]),
                                       _ptr(out["time_mel_masks"]), B, T, Te, Tn, M, _stream()))
    return out


class MelFrontend:
    """Handle of the wav -> log10-mel front-end (fse_mel_frontend_*; utils/audio/__init__.py:34-81 `librosa_wav2spec`)."""

    def __init__(self, sample_rate: int = 22050, fft_size: int = 1024, hop_size: int = 256, win_length: int = 1024, num_mels: int = 80,
                 fmin: float = 80, fmax: float = -1, eps: float = 1e-6):
        cfg = _lib.MelFrontendConfig()
        cfg.sample_rate, cfg.fft_size, cfg.hop_size, cfg.win_length, cfg.num_mels = sample_rate, fft_size, hop_size, win_length, num_mels
        cfg.fmin, cfg.fmax, cfg.eps = float(fmin), float(fmax), float(eps)
        self.cfg, self.hop, self.num_mels = cfg, hop_size, num_mels
        self._h = C.c_void_p()
        check(_lib.lib().fse_mel_frontend_create(C.byref(cfg), C.byref(self._h)))
        self._ws = _Workspace()

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _lib.lib().fse_mel_frontend_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    @property
    def last_launches(self) -> int:
        return int(_lib.lib().fse_mel_frontend_last_launches(self._h))

    def forward(self, wav: torch.Tensor) -> torch.Tensor:
        """wav[B, n] fp32 cuda -> mel[B, 1 + n // hop, num_mels] (librosa's frame count; the tail is zero-padded to a multiple of hop,
        which is what librosa's own centre padding would read there)."""
        _need_cuda(wav)
        B, n = wav.shape
        wav = wav.contiguous().float()
        pad = (-n) % self.hop
        if pad:
            wav = torch.nn.functional.pad(wav, (0, pad))
        npad = n + pad
        frames = 1 + npad // self.hop
        mel = torch.empty(B, frames, self.num_mels, dtype=torch.float32, device=wav.device)
        nbytes = _lib.lib().fse_mel_frontend_workspace_bytes(self._h, B, npad)
        ws, nbytes = self._ws.get(nbytes, wav.device)
        with _Range("fse_mel_frontend_forward"):
            check(_lib.lib().fse_mel_frontend_forward(self._h, _ptr(wav), _ptr(mel), B, npad, ws, nbytes, _stream()))
        return mel[:, :1 + n // self.hop]
