"""Host-side mirror of the reference's model seams for the hot path:

  DiffNetB200            <-> modules/speech_editing/spec_denoiser/diffnet.py::DiffNet          (:84-132)
  GaussianDiffusionB200  <-> modules/speech_editing/spec_denoiser/spec_denoiser.py::GaussianDiffusion (:16-185)
  MelEncoderB200         <-> modules/speech_editing/commons/mel_encoder.py::MelEncoder        (:3-19)
  FastSpeechB200         <-> modules/speech_editing/spec_denoiser/fs.py::FastSpeech           (:49-189, skip_decoder=True)
  CampNetB200            <-> modules/speech_editing/campnet/campnet.py::CampNet               (:14-69)

Same constructor arguments, same state_dict keys/shapes (reference checkpoints load unchanged), same
forward() signatures and return values.  The nn.Modules only OWN the parameters; every forward goes
through the C ABI (engine.Denoiser) — no torch arithmetic on the hot path and no CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import schedule as _schedule
from .engine import CampNetKernel, CondEncoderKernel, Denoiser, MelEncoderKernel
from .hparams import hparams as _global_hparams


# Arithmetic of the FluentSpeech path when hparams['b200_mode'] is not given: tcgen05 kind::tf32, the reference's own GPU arithmetic
# (cuDNN TF32 convolutions).  'tc_bf16' is ~1.6x faster at ~6x the error; 'simt_f32' is the exact-fp32 checking mode.  CampNet keeps
# 'tc_bf16' as its default (its tensor-core attention kernel is bf16).
DEFAULT_MODE = "tc_tf32"

class _ResidualBlockParams(nn.Module):
    """Parameter container with the reference ResidualBlock's names/shapes (diffnet.py:60-66)."""

    def __init__(self, encoder_hidden, residual_channels, dilation):
        super().__init__()
        C = residual_channels
        self.dilation = dilation
        self.dilated_conv = nn.Conv1d(C, 2 * C, 3, padding=dilation, dilation=dilation)
        self.diffusion_projection = nn.Linear(C, C)
        self.conditioner_projection = nn.Conv1d(encoder_hidden, 2 * C, 1)
        self.output_projection = nn.Conv1d(C, 2 * C, 1)
        for conv in (self.dilated_conv, self.conditioner_projection, self.output_projection):
            nn.init.kaiming_normal_(conv.weight)


class DiffNetB200(nn.Module):
    """Drop-in for DiffNet: `DIFF_DECODERS['wavenet'](hparams)` (tasks/speech_editing/spec_denoiser.py:13-15)."""

    def __init__(self, in_dims: int = 80, hparams: Optional[dict] = None):
        super().__init__()
        hp = _global_hparams if hparams is None else hparams
        self.in_dims = in_dims
        self.encoder_hidden = hp["hidden_size"]
        self.n_layers = hp["residual_layers"]
        self.channels = C = hp["residual_channels"]
        self.dilation_cycle_length = hp["dilation_cycle_length"]
        self.mode = hp.get("b200_mode", DEFAULT_MODE)
        self.input_projection = nn.Conv1d(in_dims, C, 1)
        nn.init.kaiming_normal_(self.input_projection.weight)
        self.mlp = nn.Sequential(nn.Linear(C, C * 4), nn.Identity(), nn.Linear(C * 4, C))   # index 1 is Mish (no params)
        self.residual_layers = nn.ModuleList([
            _ResidualBlockParams(self.encoder_hidden, C, 2 ** (i % self.dilation_cycle_length)) for i in range(self.n_layers)])
        self.skip_projection = nn.Conv1d(C, C, 1)
        nn.init.kaiming_normal_(self.skip_projection.weight)
        self.output_projection = nn.Conv1d(C, in_dims, 1)
        nn.init.zeros_(self.output_projection.weight)             # as the reference (diffnet.py:108)
        self._engine: Optional[Denoiser] = None
        self._engine_key = None

    # -- engine management -----------------------------------------------------------------------
    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self) -> Denoiser:
        """The C-ABI handle holding the repacked weights; rebuilt when any parameter changed."""
        key = self._weights_key()
        if self._engine is None or key != self._engine_key:
            eng = Denoiser(n_mels=self.in_dims, hidden=self.encoder_hidden, channels=self.channels, layers=self.n_layers,
                           dilation_cycle_length=self.dilation_cycle_length, mode=self.mode)
            eng.load_state_dict(self.state_dict())
            self._engine, self._engine_key = eng, key
        return self._engine

    def forward(self, spec: torch.Tensor, diffusion_step: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
        """spec[B,1,M,T], diffusion_step[B], cond[B,H,T] -> [B,1,M,T]   (diffnet.py:110-132)"""
        if torch.is_grad_enabled() and (cond.requires_grad or spec.requires_grad or (self.training and any(p.requires_grad for p in self.parameters()))):
            # training (GaussianDiffusion.forward(infer=False), spec_denoiser.py:168-176): native forward + activation-gradient chain,
            # weight gradients by library GEMMs (train.py)
            from .train import diffnet_train_forward
            return diffnet_train_forward(self, spec, diffusion_step.reshape(-1).long(), cond)
        x = spec[:, 0]
        # the reference hands over cond as decoder_inp.transpose(1,2): its physical layout is already [B,T,H]
        cond_bth = cond.transpose(1, 2).contiguous()
        x0 = self.engine().denoise_step(x, cond_bth, diffusion_step.reshape(-1).long())
        return x0[:, None, :, :]


class MelEncoderB200(nn.Module):
    """Drop-in for MelEncoder (mel_encoder.py:3-19): same constructor, state_dict keys (encoder.0 / encoder.2 / fc_out) and
    forward(x[B,T,M]) -> [B,T,H].  The module only owns the parameters; forward runs three tensor-core GEMM launches behind
    fse_mel_encoder_forward.  `fused(x, add, scale)` also performs the call site's `add + out * scale` (spec_denoiser.py:164)
    in the last epilogue."""

    def __init__(self, mel_bins=80, hidden_size=192, mode="tc_bf16"):
        super().__init__()
        self.mel_bins, self.hidden_size, self.mode = mel_bins, hidden_size, mode
        self.encoder = nn.Sequential(nn.Linear(mel_bins, hidden_size), nn.ReLU(), nn.Linear(hidden_size, hidden_size), nn.ReLU())
        self.fc_out = nn.Linear(hidden_size, hidden_size)
        self._engine: Optional[MelEncoderKernel] = None
        self._engine_key = None

    def engine(self) -> MelEncoderKernel:
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._engine is None or key != self._engine_key:
            eng = MelEncoderKernel(self.mel_bins, self.hidden_size, self.mode)
            eng.load_state_dict(self.state_dict())
            self._engine, self._engine_key = eng, key
        return self._engine

    @torch.no_grad()
    def forward(self, x):
        return self.engine().forward(x)

    @torch.no_grad()
    def fused(self, x, add, scale):
        return self.engine().forward(x, add, scale)


MelEncoder = MelEncoderB200     # the name the reference uses


class _ConvBlocksParams(nn.Module):
    """Parameter container with the names/shapes of ConvBlocks (modules/commons/conv.py:68-97, norm_type 'ln'):
    res_blocks.{i}.blocks.{j}.{0: LayerNorm, 1: Conv1d(C, 2C, k), 4: Conv1d(2C, C, 1)}, last_norm, post_net1."""

    def __init__(self, hidden, dilations, kernel_size, layers_in_block=2, post_net_kernel=3):
        super().__init__()

        class _Block(nn.Module):
            def __init__(self, d):
                super().__init__()
                self.blocks = nn.ModuleList([
                    nn.Sequential(nn.LayerNorm(hidden, eps=1e-5),
                                  nn.Conv1d(hidden, 2 * hidden, kernel_size, dilation=d, padding=d * (kernel_size - 1) // 2),
                                  nn.Identity(), nn.Identity(),          # indices 2, 3: the k^-0.5 scale and GELU (no parameters)
                                  nn.Conv1d(2 * hidden, hidden, 1, dilation=d))
                    for _ in range(layers_in_block)])

        self.res_blocks = nn.Sequential(*[_Block(d) for d in dilations])
        self.last_norm = nn.LayerNorm(hidden, eps=1e-5)
        self.post_net1 = nn.Conv1d(hidden, hidden, kernel_size=post_net_kernel, padding=post_net_kernel // 2)
        for m in self.modules():                                        # init_weights_func (conv.py:18-21)
            if isinstance(m, nn.Conv1d):
                nn.init.xavier_uniform_(m.weight)


def _embedding(n, dim, padding_idx=None):
    """modules/commons/layers.py:45-50."""
    m = nn.Embedding(n, dim, padding_idx=padding_idx)
    nn.init.normal_(m.weight, mean=0, std=dim ** -0.5)
    if padding_idx is not None:
        nn.init.constant_(m.weight[padding_idx], 0)
    return m


class _PredictorParams(nn.Module):
    """conv.{i}.{0: Conv1d, 2: LayerNorm} + linear, the layout shared by DurationPredictor and PitchPredictor
    (modules/commons/nar_tts_modules.py:8-22, 75-88)."""

    def __init__(self, idim, n_layers, n_chans, kernel_size, odim, dur: bool):
        super().__init__()
        self.conv = nn.ModuleList([
            nn.Sequential(nn.Conv1d(idim if i == 0 else n_chans, n_chans, kernel_size, padding=kernel_size // 2), nn.Identity(),
                          nn.LayerNorm(n_chans), nn.Identity()) for i in range(n_layers)])
        self.linear = nn.Sequential(nn.Linear(n_chans, odim), nn.Identity()) if dur else nn.Linear(n_chans, odim)


class FastSpeechB200(nn.Module):
    """Drop-in for the reference's condition encoder `FastSpeech` (fs.py:49-189) on the path the speech-editing model uses:
    encoder_type 'conv', skip_decoder=True.  Same constructor (dict_size, hparams, out_dims), same state_dict keys and shapes
    (the unused `decoder.*` / `mel_out.*` parameters are kept so reference checkpoints strict-load), same methods with the same
    arguments — forward, forward_style_embed, forward_dur (incl. the inference script's masked_dur / use_pred_mel2ph call,
    inference/tts/spec_denoiser.py:84-98) — and `self.encoder(txt_tokens)` is callable.  The module only owns the parameters:
    every tensor operation is a kernel behind the fse_cond_* C ABI (no torch arithmetic, no CPU fallback)."""

    def __init__(self, dict_size, hparams: Optional[dict] = None, out_dims=None):
        super().__init__()
        hp = dict(_global_hparams if hparams is None else hparams)
        self.hparams = hp
        H = self.hidden_size = hp["hidden_size"]
        if hp.get("encoder_type", "conv") != "conv":
            raise NotImplementedError("FastSpeechB200 implements encoder_type 'conv' (egs/spec_denoiser*.yaml); other encoders are "
                                      "other model families (SURVEY.md section 2 row 21)")
        if hp.get("use_spk_id", False):
            raise NotImplementedError("use_spk_id is false in the speech-editing configs; only the spk_embed branch is native")
        self.dict_size = dict_size
        self.mode = hp.get("b200_mode", DEFAULT_MODE)
        self.encoder = _ConvBlocksParams(H, hp.get("enc_dilations", [1, 1, 1, 1]), hp.get("enc_kernel_size", 5),
                                         hp.get("layers_in_block", 2), hp.get("enc_post_net_kernel", 3))
        self.encoder.embed_tokens = _embedding(dict_size, H, 0)
        self.encoder.forward = self._encode                     # `model.fs.encoder(txt_tokens)` (inference/tts/spec_denoiser.py:84)
        if hp.get("decoder_type", "conv") == "conv":            # unused with skip_decoder=True; kept for checkpoint compatibility
            self.decoder = _ConvBlocksParams(H, hp.get("dec_dilations", [1, 1, 1, 1]), hp.get("dec_kernel_size", 5),
                                             hp.get("layers_in_block", 2), hp.get("dec_post_net_kernel", 3))
        self.out_dims = hp.get("audio_num_mel_bins", 80) if out_dims is None else out_dims
        self.mel_out = nn.Linear(H, self.out_dims, bias=True)
        self.use_spk_embed = bool(hp.get("use_spk_embed", True))
        if self.use_spk_embed:
            self.spk_embed_proj = nn.Linear(256, H, bias=True)
        ph = hp.get("predictor_hidden", -1)
        if ph > 0 and ph != H:
            raise NotImplementedError("predictor_hidden must equal hidden_size (-1 in the shipped configs)")
        self.dur_embed = _embedding(2000, H, 0)
        self.dur_predictor = _PredictorParams(H, hp.get("dur_predictor_layers", 3), H, hp.get("dur_predictor_kernel", 5), 1, dur=True)
        self.use_pitch_embed = bool(hp.get("use_pitch_embed", True))
        if self.use_pitch_embed:
            self.pitch_embed = _embedding(300, H, 0)
            self.pitch_predictor = _PredictorParams(H, 5, H, hp.get("predictor_kernel", 5), 2, dur=False)
        self._engine: Optional[CondEncoderKernel] = None
        self._engine_key = None

    def engine(self) -> CondEncoderKernel:
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._engine is None or key != self._engine_key:
            hp = self.hparams
            use_uv = hp.get("pitch_type", "frame") == "frame" and bool(hp.get("use_uv", True))
            eng = CondEncoderKernel(self.dict_size, self.hidden_size, tuple(hp.get("enc_dilations", [1, 1, 1, 1])),
                                    hp.get("enc_kernel_size", 5), hp.get("layers_in_block", 2), hp.get("enc_post_net_kernel", 3),
                                    hp.get("dur_predictor_layers", 3), hp.get("dur_predictor_kernel", 5), 5, hp.get("predictor_kernel", 5),
                                    self.use_pitch_embed, use_uv, 256 if self.use_spk_embed else 0, self.mode)
            eng.load_state_dict(self.state_dict())
            self._engine, self._engine_key = eng, key
        return self._engine

    # -- the reference's methods -----------------------------------------------------------------
    @torch.no_grad()
    def _encode(self, txt_tokens):
        return self.engine().text_encoder(txt_tokens)

    @torch.no_grad()
    def forward_style_embed(self, spk_embed=None, spk_id=None):
        """fs.py:114-121 -> [B,1,H] (0 without a speaker embedding)."""
        if not self.use_spk_embed:
            return 0
        return self.engine().style_embed(spk_embed)[:, None, :]

    @torch.no_grad()
    def forward_dur(self, dur_input, time_mel_masks, mel2ph, txt_tokens, ret, masked_dur=None, use_pred_mel2ph=False):
        """fs.py:123-151."""
        eng = self.engine()
        if masked_dur is None:
            masked_dur = eng.masked_dur(mel2ph, time_mel_masks, txt_tokens)
        ret["dur"] = dur = eng.duration(dur_input, masked_dur, txt_tokens)
        if use_pred_mel2ph:
            mel2ph = eng.length_regulate(dur, txt_tokens)
        fm = self.hparams.get("frames_multiple", 1)                       # clip_mel2token_to_multiple (align_ops.py:15-18)
        ret["mel2ph"] = mel2ph = mel2ph[:, :mel2ph.shape[1] // fm * fm]
        return mel2ph

    @torch.no_grad()
    def forward(self, txt_tokens, time_mel_masks, mel2ph, spk_embed, f0, uv, spk_id=None, skip_decoder=True, infer=False,
                use_pred_mel2ph=False, use_pred_pitch=False, **kwargs):
        """fs.py:83-105 with skip_decoder=True: returns the reference's dict (decoder_inp, dur, mel2ph, pitch_pred, f0_denorm,
        f0_denorm_pred)."""
        if not skip_decoder:
            raise NotImplementedError("the speech-editing model calls FastSpeech with skip_decoder=True (spec_denoiser.py:159-161); "
                                      "the FastSpeech mel decoder is not on this path")
        eng = self.engine()
        ret = {}
        encoder_out = eng.text_encoder(txt_tokens)
        style = eng.style_embed(spk_embed) if self.use_spk_embed else None
        dur_inp = eng.dur_input(encoder_out, style, txt_tokens)
        mel2ph = self.forward_dur(dur_inp, time_mel_masks, mel2ph, txt_tokens, ret, use_pred_mel2ph=use_pred_mel2ph)
        out = eng.frames(encoder_out, style, mel2ph, time_mel_masks, f0, uv, use_pred_pitch)
        for k in ("pitch_pred", "f0_denorm", "f0_denorm_pred", "decoder_inp"):
            if k in out:
                ret[k] = out[k]
        return ret


FastSpeech = FastSpeechB200     # the name the reference uses


class _MHAParams(nn.Module):
    """MultiheadAttention with bias=False (speech_editing/commons/transformer.py:160-174): in_proj_weight [3C, C], out_proj.weight."""

    def __init__(self, c):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * c, c))
        self.out_proj = nn.Linear(c, c, bias=False)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.xavier_uniform_(self.out_proj.weight)


class _FFNParams(nn.Module):
    """TransformerFFNLayer (transformer.py:74-90): ffn_1 = Conv1d ('SAME') or Sequential(ConstantPad1d, Conv1d) ('LEFT'), ffn_2 = Linear."""

    def __init__(self, c, k, left):
        super().__init__()
        conv = nn.Conv1d(c, 4 * c, k, padding=0 if left else k // 2)
        self.ffn_1 = nn.Sequential(nn.Identity(), conv) if left else conv
        self.ffn_2 = nn.Linear(4 * c, c)


class _SALayerParams(nn.Module):
    """EncSALayer / DecSALayer parameters under `.op` (transformer.py:489-546, 611-636)."""

    def __init__(self, c, k, decoder):
        super().__init__()
        op = nn.Module()
        op.layer_norm1 = nn.LayerNorm(c)
        op.self_attn = _MHAParams(c)
        op.layer_norm2 = nn.LayerNorm(c)
        if decoder:
            op.encoder_attn = _MHAParams(c)
            op.layer_norm3 = nn.LayerNorm(c)
        op.ffn = _FFNParams(c, k, left=decoder)
        self.op = op


class _PosEmbedParams(nn.Module):
    """SinusoidalPositionalEmbedding keeps one persistent buffer, `_float_tensor` (transformer.py:30)."""

    def __init__(self):
        super().__init__()
        self.register_buffer("_float_tensor", torch.zeros(1))


class CampNetB200(nn.Module):
    """Drop-in for the reference's CampNet on its forward (campnet.py:14-69): same constructor (ph_dict_size, word_dict_size,
    hparams, out_dims), the same 237 state_dict keys and shapes (unused ones included, so checkpoints strict-load) and
    forward(txt_tokens, spk_embed=None, spk_id=None, mels=None, stutter_mel_masks=None, time_mel_masks=None, infer=False,
    global_step=None) -> dict(mel_out_coarse, mel_out_fine, attn).  The module only owns the parameters; the forward is one
    C-ABI call (fse_campnet_forward)."""

    def __init__(self, ph_dict_size, word_dict_size=None, hparams: Optional[dict] = None, out_dims=None):
        super().__init__()
        hp = dict(_global_hparams if hparams is None else hparams)
        self.hparams = hp
        C_ = self.hidden_size = hp["hidden_size"]
        k = self.ffn_kernel = hp.get("dec_ffn_kernel_size", 9)
        self.out_dims = hp.get("audio_num_mel_bins", 80) if out_dims is None else out_dims
        self.ph_dict_size = ph_dict_size
        self.mode = hp.get("b200_mode", "tc_bf16")
        enc = nn.Module()
        enc.layers = nn.ModuleList([_SALayerParams(C_, k, decoder=False) for _ in range(3)])
        enc.layer_norm = nn.LayerNorm(C_)
        enc.embed_tokens = _embedding(ph_dict_size, C_, 0)
        enc.pre_net = _ConvBlocksParams(C_, [1] * 3, 1, 2, 3)            # constructed by the reference, never called (transformer.py:751)
        enc.embed_positions = _PosEmbedParams()
        self.encoder = enc
        self.mel_out = nn.Linear(C_, self.out_dims, bias=True)             # inherited from FastSpeech, unused
        self.mel_encoder = nn.Module()
        self.mel_encoder.encoder = nn.Sequential(nn.Linear(self.out_dims, C_), nn.ReLU(), nn.Linear(C_, C_), nn.ReLU())
        self.mel_encoder.fc_out = nn.Linear(C_, C_)
        dec = nn.Module()
        dec.pos_embed_alpha = nn.Parameter(torch.ones(1))
        dec.embed_positions = _PosEmbedParams()
        dec.layers = nn.ModuleList([_SALayerParams(C_, k, decoder=True) for _ in range(6)])
        dec.layer_norm = nn.LayerNorm(C_)
        self.decoder_coarse = dec
        self.decoder_fine = _ConvBlocksParams(C_, [1] * 5, 5, 2, 3)
        self.mel_out_coarse = nn.Linear(C_, self.out_dims, bias=False)
        self.mel_out_fine = nn.Linear(C_, self.out_dims, bias=False)
        self.mask_emb = nn.Parameter(torch.zeros(1, 1, self.out_dims))
        self._engine: Optional[CampNetKernel] = None
        self._engine_key = None

    def engine(self) -> CampNetKernel:
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._engine is None or key != self._engine_key:
            eng = CampNetKernel(self.ph_dict_size, self.hidden_size, self.out_dims, 3, 6, 2, self.ffn_kernel, 5, 5, self.mode)
            eng.load_state_dict(self.state_dict())
            self._engine, self._engine_key = eng, key
        return self._engine

    @torch.no_grad()
    def forward(self, txt_tokens, spk_embed=None, spk_id=None, mels=None, stutter_mel_masks=None, time_mel_masks=None, infer=False,
                global_step=None, *args, **kwargs):
        """campnet.py:40-45 (spk_embed / spk_id / stutter_mel_masks are accepted and ignored, as in the reference)."""
        out = self.engine().forward(txt_tokens, mels, time_mel_masks)
        return {"mel_out_coarse": out["mel_out_coarse"], "mel_out_fine": out["mel_out_fine"], "attn": out["attn"]}


CampNet = CampNetB200     # the name the reference uses


class GaussianDiffusionB200(nn.Module):
    """Drop-in for GaussianDiffusion: the sampling path (infer=True) and the training branch (infer=False, one denoiser evaluation
    under autograd through the native training kernels).

    `fs` (the FastSpeech condition encoder, fs.py:49-112) defaults to the native FastSpeechB200 and `mel_encoder` to
    MelEncoderB200, so the whole forward(infer=True) — text tokens to mel — runs behind the C ABI; a reference FastSpeech
    module may still be injected (`fs=`, `from_reference(native_fs=False)`).  Everything from `cond` on — x_S draw, the S
    p_sample iterations and the mask compositing — is one C-ABI call (fse_sample)."""

    def __init__(self, phone_encoder, out_dims, denoise_fn, timesteps=1000, time_scale=1, loss_type="l1", betas=None,
                 spec_min=None, spec_max=None, fs: Optional[nn.Module] = None, mel_encoder: Optional[nn.Module] = None,
                 hparams: Optional[dict] = None):
        super().__init__()
        hp = _global_hparams if hparams is None else hparams
        self.denoise_fn = denoise_fn
        if fs is None and phone_encoder is not None:            # as the reference: self.fs = FastSpeech(len(phone_encoder), hparams)
            fs = FastSpeechB200(len(phone_encoder), hp, out_dims)
        self.fs = fs
        self.mel_encoder = mel_encoder if mel_encoder is not None else MelEncoderB200(out_dims, hp.get("hidden_size", 192),
                                                                                      hp.get("b200_mode", DEFAULT_MODE))
        self.mel_bins = out_dims
        self.time_scale = time_scale
        self.num_timesteps = int(timesteps)
        self.loss_type = loss_type
        if betas is not None and isinstance(betas, torch.Tensor):
            betas = betas.detach().cpu().numpy()
        bufs = _schedule.diffusion_buffers(self.num_timesteps, hp.get("schedule_type", "vpsde"), betas)
        self.register_buffer("timesteps", torch.tensor(float(self.num_timesteps)))
        self.register_buffer("timescale", torch.tensor(float(self.time_scale)))
        for name in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                     "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                     "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"):
            self.register_buffer(name, torch.from_numpy(bufs[name]))
        keep = hp.get("keep_bins", out_dims)
        self.register_buffer("spec_min", torch.FloatTensor(spec_min or [])[None, None, :keep])
        self.register_buffer("spec_max", torch.FloatTensor(spec_max or [])[None, None, :keep])
        self._sched_key = None

    @classmethod
    def from_reference(cls, ref_model, mode: str = DEFAULT_MODE, native_fs: bool = True):
        """Convert an instantiated reference GaussianDiffusion: copies the denoise_fn / mel_encoder / fs weights into the native
        drop-ins (native_fs=False keeps the reference's own FastSpeech module as the condition encoder)."""
        rd = ref_model.denoise_fn
        hp = dict(hidden_size=rd.params.encoder_hidden, residual_layers=rd.params.residual_layers,
                  residual_channels=rd.params.residual_channels, dilation_cycle_length=rd.params.dilation_cycle_length,
                  b200_mode=mode, keep_bins=ref_model.mel_bins)
        den = DiffNetB200(ref_model.mel_bins, hp)
        den.load_state_dict(rd.state_dict())
        enc = MelEncoderB200(ref_model.mel_bins, rd.params.encoder_hidden, mode)
        enc.load_state_dict(ref_model.mel_encoder.state_dict())
        fs = ref_model.fs
        if native_fs:
            fhp = dict(ref_model.fs.hparams)
            fhp["b200_mode"] = mode
            fs = FastSpeechB200(ref_model.fs.encoder.embed_tokens.num_embeddings, fhp, ref_model.fs.out_dims)
            fs.load_state_dict(ref_model.fs.state_dict(), strict=True)
        new = cls(None, ref_model.mel_bins, den, timesteps=ref_model.num_timesteps, time_scale=ref_model.time_scale,
                  loss_type=ref_model.loss_type, fs=fs, mel_encoder=enc, hparams=hp)
        for name, buf in ref_model.named_buffers(recurse=False):     # the reference's own float64-derived schedule, bit for bit
            if name in new._buffers:
                new._buffers[name] = buf.detach().clone()
        return new.to(next(ref_model.parameters()).device)

    # -- engine plumbing -------------------------------------------------------------------------
    def _engine(self) -> Denoiser:
        eng = self.denoise_fn.engine()
        key = (id(eng), self.posterior_mean_coef1.data_ptr(), self.posterior_mean_coef1._version)
        if key != self._sched_key:
            eng.set_schedule(self.posterior_mean_coef1, self.posterior_mean_coef2, self.posterior_log_variance_clipped)
            self._sched_key = key
        return eng

    # -- reference API ---------------------------------------------------------------------------
    def q_posterior(self, x_start, x_t, t):
        """spec_denoiser.py:86-93 (tiny gathers; plain torch)."""
        shape = (t.shape[0],) + (1,) * (x_t.dim() - 1)
        mean = self.posterior_mean_coef1[t].reshape(shape) * x_start + self.posterior_mean_coef2[t].reshape(shape) * x_t
        return mean, self.posterior_variance[t].reshape(shape), self.posterior_log_variance_clipped[t].reshape(shape)

    @torch.no_grad()
    def q_posterior_sample(self, x_start, x_t, t, noise=None, seed=0, step=0):
        """spec_denoiser.py:95-101; x_*[B,1,M,T]."""
        out = self._engine().posterior_step(x_start[:, 0], x_t[:, 0], t, None if noise is None else noise[:, 0], seed, step)
        return out[:, None]

    @torch.no_grad()
    def p_sample(self, x_t, t, cond, spk_emb=None, clip_denoised=True, repeat_noise=False, noise=None, seed=0):
        """spec_denoiser.py:103-108."""
        x0 = self.denoise_fn(x_t, t, cond)
        return self.q_posterior_sample(x0, x_t, t, noise=noise, seed=seed, step=int(t[0]) + 1)

    @torch.no_grad()
    def sample(self, cond_bth: torch.Tensor, noise=None, seed: int = 0, ref_mels=None, time_mel_masks=None):
        """cond[B,T,H] -> mel_out[B,T,M] (optionally composited with ref_mels by the 0/1 mask)."""
        mask = None if time_mel_masks is None else time_mel_masks.reshape(cond_bth.shape[0], cond_bth.shape[1])
        return self._engine().sample(cond_bth, noise, seed, ref_mels if mask is not None else None, mask)

    def forward(self, txt_tokens, time_mel_masks, mel2ph, spk_embed, ref_mels, f0, uv, energy=None, infer=False,
                use_pred_mel2ph=False, use_pred_pitch=False, noise=None, seed=None, composite=False):
        """spec_denoiser.py:154-185, both branches.  Returns the reference's dict (mel_out[B,T,M], ...).
        Extras over the reference signature: `noise` ([(S+1),B,M,T] injected draws, tests), `seed` (Philox key) and `composite`
        (mel_out * mask + ref_mels * (1 - mask), the call sites' next line — tasks/speech_editing/spec_denoiser.py:53,84,
        inference/tts/spec_denoiser.py:136 — done in the last step's epilogue)."""
        if self.fs is None:
            raise RuntimeError("no condition encoder: pass fs=<reference FastSpeech> or use GaussianDiffusionB200.from_reference()")
        ret = self.fs(txt_tokens, time_mel_masks, mel2ph, spk_embed, f0, uv, energy, skip_decoder=True, infer=infer,
                      use_pred_mel2ph=use_pred_mel2ph, use_pred_pitch=use_pred_pitch)
        decoder_inp = ret["decoder_inp"]
        tgt_nonpadding = (mel2ph > 0).float()[:, :, None]
        if isinstance(self.mel_encoder, MelEncoderB200):     # add + MelEncoder(x) * nonpadding in the last GEMM's epilogue
            decoder_inp = self.mel_encoder.fused(ref_mels * (1 - time_mel_masks), decoder_inp, tgt_nonpadding)
        else:
            decoder_inp = decoder_inp + self.mel_encoder(ref_mels * (1 - time_mel_masks)) * tgt_nonpadding
        ret["decoder_inp"] = decoder_inp
        if not infer:
            # the training branch (spec_denoiser.py:168-176): t ~ U{0..S}, x_t = q_sample(ref, t) * nonpadding, one denoiser evaluation.
            # DiffNetB200 runs its native forward + activation-gradient chain under autograd (train.py); the gradient reaches the
            # condition encoder through `decoder_inp` when that is a torch module (a reference FastSpeech / MelEncoder injected via
            # from_reference(native_fs=False)); the native FastSpeechB200 / MelEncoderB200 are forward-only, so with them the
            # denoiser branch alone is trained.  `noise`: an injected [B,1,M,T] draw, `seed`: unused here (torch's generator draws t).
            b = txt_tokens.shape[0]
            nonpadding = (mel2ph != 0).float()[:, None, None, :]
            t = torch.randint(0, self.num_timesteps + 1, (b,), device=decoder_inp.device).long()
            x_t = self.diffuse_fn(ref_mels, t, noise=noise) * nonpadding
            x_0_pred = self.denoise_fn(x_t, t, decoder_inp.transpose(1, 2)) * nonpadding
            ret["mel_out"] = x_0_pred[:, 0].transpose(1, 2)
            return ret
        if seed is None:
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())      # follows torch.manual_seed like the reference's randn
        if composite:
            B, T = decoder_inp.shape[:2]
            x = self._engine().sample(decoder_inp.contiguous(), noise, seed, ref_mels, time_mel_masks.reshape(B, T))
        else:
            x = self._engine().sample(decoder_inp.contiguous(), noise, seed)  # = x[:, 0].transpose(1, 2) of the reference loop
        ret["mel_out"] = x
        return ret

    def q_sample(self, x_start, t, noise=None):
        """spec_denoiser.py:126-132"""
        if noise is None:
            noise = torch.randn_like(x_start)
        shape = (t.shape[0],) + (1,) * (x_start.dim() - 1)
        return self.sqrt_alphas_cumprod[t].reshape(shape) * x_start + self.sqrt_one_minus_alphas_cumprod[t].reshape(shape) * noise

    def diffuse_fn(self, x_start, t, noise=None):
        """spec_denoiser.py:144-152: [B,T,M] -> x_t [B,1,M,T]; items with t < 0 keep the ground-truth mel (and t is clamped in place)."""
        x_start = self.norm_spec(x_start).transpose(1, 2)[:, None, :, :]
        zero_idx = t < 0
        t[zero_idx] = 0
        out = self.q_sample(x_start, t, noise)
        out[zero_idx] = x_start[zero_idx]
        return out

    def norm_spec(self, x):
        return x

    def denorm_spec(self, x):
        return x

    def out2mel(self, x):
        return x
