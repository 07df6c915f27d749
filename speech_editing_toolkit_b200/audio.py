"""Mirror of the reference's mel front-end entry point for arrays: `librosa_wav2spec` (utils/audio/__init__.py:34-81), the call
the inference script makes on the original recording (inference/tts/spec_denoiser.py:258).  Same keyword arguments and the same
`{'wav', 'mel'}` entries of the returned dict; the transform runs on the GPU behind fse_mel_frontend_forward.  Loading a wav
FILE (librosa.core.load) and loudness normalisation / silence trimming (pyloudnorm, webrtcvad) stay outside: pass samples."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from .engine import MelFrontend

_CACHE: Dict[Tuple, MelFrontend] = {}


def librosa_pad_lr(n: int, fsize: int, fshift: int, pad_sides: int = 1):
    """utils/audio/__init__.py:8-17 on a length."""
    assert pad_sides in (1, 2)
    pad = (n // fshift + 1) * fshift - n
    return (0, pad) if pad_sides == 1 else (pad // 2, pad // 2 + pad % 2)


def wav2spec(wav, fft_size=1024, hop_size=256, win_length=1024, window="hann", num_mels=80, fmin=80, fmax=-1, eps=1e-6, sample_rate=22050,
             loud_norm=False, trim_long_sil=False, device="cuda"):
    """librosa_wav2spec for an array of samples -> {'wav': padded / trimmed samples (:73-75), 'mel': [T, num_mels] float32 log10}."""
    if isinstance(wav, str):
        raise NotImplementedError("pass samples: reading audio files needs librosa, which is outside this path")
    if loud_norm or trim_long_sil:
        raise NotImplementedError("loud_norm / trim_long_sil use pyloudnorm / webrtcvad and are outside this path (both false in the shipped configs)")
    if window != "hann":
        raise NotImplementedError("only the Hann window of the shipped configs is built")
    key = (sample_rate, fft_size, hop_size, win_length, num_mels, float(fmin), float(fmax), float(eps))
    if key not in _CACHE:
        _CACHE[key] = MelFrontend(sample_rate, fft_size, hop_size, win_length, num_mels, fmin, fmax, eps)
    x = np.ascontiguousarray(wav, dtype=np.float32)
    mel = _CACHE[key].forward(torch.from_numpy(x)[None].to(device))[0].cpu().numpy()
    l_pad, r_pad = librosa_pad_lr(len(x), fft_size, hop_size, 1)
    out_wav = np.pad(x, (l_pad, r_pad), mode="constant", constant_values=0.0)[:mel.shape[0] * hop_size]
    return {"wav": out_wav, "mel": mel}
