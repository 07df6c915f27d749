"""Vocoder plugin seam of the reference (tasks/tts/vocoder_infer/base_vocoder.py:6-29,
tasks/tts/vocoder_infer/hifigan.py:11-31, inference/tts/base_tts_infer.py:36-47) over the C-ABI HiFi-GAN."""
from __future__ import annotations

import glob
import re

import numpy as np
import torch
import yaml

from .engine import HIFIGAN_V1, Vocoder
from .hparams import hparams

REGISTERED_VOCODERS = {}


def register_vocoder(name):
    def _f(cls):
        REGISTERED_VOCODERS[name] = cls
        return cls
    return _f


def get_vocoder_cls(vocoder_name):
    return REGISTERED_VOCODERS.get(vocoder_name)


class BaseVocoder:
    def spec2wav(self, mel):
        """mel [T, 80] -> wav [T * hop]"""
        raise NotImplementedError

    @staticmethod
    def wav2spec(wav):
        """tasks/tts/vocoder_infer/base_vocoder.py:30-47 for an array of samples (reading a file needs librosa): -> (wav, mel [T, 80]),
        the transform on the GPU (audio.wav2spec -> fse_mel_frontend_forward), parameters from the same hparams keys."""
        from .audio import wav2spec as _wav2spec
        d = _wav2spec(wav, fft_size=hparams["fft_size"], hop_size=hparams["hop_size"], win_length=hparams["win_size"],
                      num_mels=hparams["audio_num_mel_bins"], fmin=hparams["fmin"], fmax=hparams["fmax"],
                      sample_rate=hparams["audio_sample_rate"], loud_norm=hparams["loud_norm"])
        return d["wav"], d["mel"]


def _generator_config(cfg: dict) -> dict:
    keys = ("upsample_rates", "upsample_kernel_sizes", "upsample_initial_channel", "resblock", "resblock_kernel_sizes",
            "resblock_dilation_sizes")
    return {k: cfg[k] for k in keys}


@register_vocoder("HifiGAN_B200")
@register_vocoder("HifiGAN")
class HifiGANB200(BaseVocoder):
    """Same construction contract as the reference's HifiGAN wrapper: reads `<vocoder_ckpt>/config.yaml` and the
    newest `model_ckpt_steps_*.ckpt` (state_dict['model_gen'], weight_g / weight_v pairs)."""

    def __init__(self, base_dir: str = None, config: dict = None, state_dict: dict = None, mode: str = None):
        mode = mode or hparams.get("b200_mode", "tc_tf32")      # the reference's GPU arithmetic (see modules.DEFAULT_MODE)
        if state_dict is None:
            base_dir = base_dir or hparams["vocoder_ckpt"]
            with open(f"{base_dir}/config.yaml") as f:
                config = yaml.safe_load(f)
            ckpts = sorted(glob.glob(f"{base_dir}/model_ckpt_steps_*.ckpt"),
                           key=lambda p: -int(re.findall(r"steps_(\d+)\.ckpt", p)[0]))
            if not ckpts:
                raise FileNotFoundError(f"no model_ckpt_steps_*.ckpt under {base_dir}")
            state_dict = torch.load(ckpts[0], map_location="cpu", weights_only=False)["state_dict"]["model_gen"]
        self.config = dict(HIFIGAN_V1 if config is None else config)
        self.device = torch.device("cuda")
        self.model = Vocoder(_generator_config(self.config), n_mels=self.config.get("audio_num_mel_bins", 80), mode=mode)
        self.model.load_state_dict(state_dict)

    def spec2wav(self, mel, **kwargs):
        """tasks/tts/vocoder_infer/hifigan.py:23-31: numpy/torch [T,80] -> numpy float32 [T*hop] (host in, host out)."""
        mel = mel.detach().cpu().numpy() if isinstance(mel, torch.Tensor) else np.asarray(mel)
        return self.model.forward_host(mel[None].astype(np.float32))[0]

    def __call__(self, c: torch.Tensor) -> torch.Tensor:
        """run_vocoder contract (inference/tts/base_tts_infer.py:44-47): c[B,T,80] cuda -> [B, T*hop] cuda."""
        return self.model.forward(c)
