"""CPU oracle of the mel front-end of the inference entry point: wav -> log10-mel.

TEST INFRASTRUCTURE — NOT PRODUCT CODE (same rules as oracle/fluentspeech_oracle.py).

Restates utils/audio/__init__.py:34-81 (`librosa_wav2spec`; called at inference/tts/spec_denoiser.py:258 with fmin=55,
fmax=7600, sample_rate=22050 and the defaults fft_size=1024, hop_size=256, win_length=1024, window='hann', num_mels=80, eps=1e-6):
  x_stft = librosa.stft(wav, n_fft, hop_length, win_length, window, pad_mode="constant")   # center=True: n_fft/2 zeros both sides
  mel    = log10(max(eps, librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) @ |x_stft|))

PARITY UNPINNED: the arithmetic lives in librosa (requirements.txt pins none; the `librosa.stft(..., pad_mode=...)` /
`librosa.filters.mel(sr=..., n_fft=...)` keyword forms are librosa >= 0.8), which is absent from this container and from
/root/reference, so neither the reference function nor golden vectors of it can be produced here.  What is restated is
librosa's published algorithm:
  stft: periodic Hann window (scipy get_window(..., fftbins=True)), frame t covers samples [t*hop - n_fft/2, t*hop + n_fft/2)
        of the zero-padded signal, n_frames = 1 + len(wav) // hop, rfft per frame (complex64 in librosa for float32 input)
  filters.mel: Slaney mel scale (htk=False: linear below 1 kHz at 200/3 Hz per mel, log above with step ln(6.4)/27), n_mels + 2
        band edges, triangular weights on the rfft bin centres, Slaney area normalisation 2 / (f[i+2] - f[i]), float32
Secondary pins used by tests/test_mel_frontend_oracle.py: torch.stft (an independent implementation of the same transform), the
filterbank's invariants (band edges, unit-area triangles, zero outside [fmin, fmax]), the fixture tests/golden/mel_frontend.npz
made without this file (scipy.signal.stft + a Slaney filterbank checked against the constants in librosa's documentation), and
two third-party implementations that are documented and tested upstream to reproduce librosa — torchaudio's
melscale_fbanks / MelSpectrogram with norm="slaney", mel_scale="slaney" (filterbank equal to 7e-8, whole wav -> log10-mel pipeline to
2e-6) and transformers.audio_utils.mel_filter_bank (equal to 1e-9).  librosa's own outputs remain unavailable here.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def hann_periodic(n: int) -> np.ndarray:
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float64)


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-12) / min_log_hz) / logstep, f / f_sp)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_basis(sr=22050, n_fft=1024, n_mels=80, fmin=55.0, fmax=7600.0) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults htk=False, norm='slaney' -> [n_mels, n_fft/2+1] float32."""
    fftfreqs = np.linspace(0, sr / 2.0, n_fft // 2 + 1)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, len(fftfreqs)))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(F32)


def stft_mag(wav, n_fft=1024, hop=256) -> np.ndarray:
    """|librosa.stft(wav, n_fft, hop, n_fft, 'hann', center=True, pad_mode='constant')| -> [n_frames, n_fft/2+1] float32."""
    wav = np.asarray(wav, dtype=F32)
    n_frames = 1 + len(wav) // hop
    x = np.pad(wav.astype(np.float64), (n_fft // 2, n_fft // 2))
    win = hann_periodic(n_fft)
    frames = np.stack([x[t * hop:t * hop + n_fft] * win for t in range(n_frames)])
    return np.abs(np.fft.rfft(frames, axis=1)).astype(F32)


def wav2mel(wav, sr=22050, n_fft=1024, hop=256, n_mels=80, fmin=55.0, fmax=7600.0, eps=1e-6) -> np.ndarray:
    """librosa_wav2spec(...)['mel'] (utils/audio/__init__.py:59-75): [n_frames, n_mels] float32, log10."""
    mag = stft_mag(wav, n_fft, hop)
    mel = mag.astype(np.float64) @ mel_basis(sr, n_fft, n_mels, fmin, fmax).astype(np.float64).T
    return np.log10(np.maximum(eps, mel)).astype(F32)
