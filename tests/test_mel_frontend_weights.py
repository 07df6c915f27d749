"""The mel front-end's conv-GEMM formulation on the CPU: the weight matrices are built by the product's own C++ header
(csrc/mel_frontend_weights.h, compiled here with g++), the two GEMMs of mel_frontend.cu are emulated with numpy in exactly the
layout the kernel consumes (rows of hop samples, R taps at offsets -R/2..R/2-1, zero rows outside the signal, re/im in adjacent
columns), and the result is compared with oracle/mel_frontend_oracle.py."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import mel_frontend_oracle as MO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_FFT, HOP, NM, SR = 1024, 256, 80, 22050
NDFT = (2 * (N_FFT // 2 + 1) + 63) // 64 * 64


@pytest.fixture(scope="module")
def weights(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("mf") / "libmf_host.so"
    subprocess.run([gxx, "-std=c++17", "-O2", "-Wall", "-Werror", "-shared", "-fPIC", os.path.join(ROOT, "tests", "tools", "mel_frontend_host.cpp"),
                    "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    R = N_FFT // HOP
    dft = np.zeros((NDFT, HOP, R), dtype=np.float32)
    lib.mf_dft_weights(N_FFT, HOP, NDFT, C.c_void_p(dft.ctypes.data))
    mel = np.zeros((NM, NDFT // 2), dtype=np.float32)
    lib.mf_mel_weights.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p]
    lib.mf_mel_weights(SR, N_FFT, NM, 55.0, 7600.0, NDFT // 2, C.c_void_p(mel.ctypes.data))
    return dft, mel


def emulate(wav, dft, mel, eps=1e-6):
    """What fse_mel_frontend_forward computes, in float64 numpy over the fp32 weights."""
    R = N_FFT // HOP
    Tc = len(wav) // HOP
    rows = wav.reshape(Tc, HOP).astype(np.float64)
    Tf = Tc + 1
    acc = np.zeros((Tf, NDFT))
    for r in range(R):
        off = r - R // 2
        shifted = np.zeros((Tf, HOP))
        lo, hi = max(0, -off), min(Tf, Tc - off)
        if hi > lo:
            shifted[lo:hi] = rows[lo + off:hi + off]
        acc += shifted @ dft[:, :, r].astype(np.float64).T
    mag = np.sqrt(acc[:, 0::2] ** 2 + acc[:, 1::2] ** 2)
    return mag, np.log10(np.maximum(eps, mag @ mel.astype(np.float64).T))


def test_mel_weights_equal_the_oracles_filterbank(weights):
    _, mel = weights
    ref = MO.mel_basis(SR, N_FFT, NM, 55.0, 7600.0)
    assert np.abs(mel[:, :513] - ref).max() < 1e-7 and np.abs(mel[:, 513:]).max() == 0.0


def test_conv_gemm_formulation_equals_stft_and_log_mel(weights):
    dft, mel = weights
    rs = np.random.RandomState(1)
    for n in (HOP * 12, HOP * 3, HOP * 1):                      # also signals shorter than one window
        wav = (rs.standard_normal(n) * 0.2).astype(np.float32)
        mag, logmel = emulate(wav, dft, mel)
        assert mag.shape == (1 + n // HOP, NDFT // 2)
        want_mag = MO.stft_mag(wav, N_FFT, HOP)
        assert np.abs(mag[:, :513] - want_mag).max() < 1e-4 and np.abs(mag[:, 513:]).max() == 0.0
        assert np.abs(logmel - MO.wav2mel(wav)).max() < 1e-4
    t = np.arange(HOP * 40) / SR
    tone = (0.4 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32)
    # a pure tone has 100 dB of dynamic range inside a frame: the leakage bins sit at the fp32 noise floor of the transform
    # (librosa's complex64 FFT has the same floor with a different realisation), so the comparison is in the linear domain,
    # relative to the frame's strongest band; the bands that carry the tone still agree tightly in log10
    got, want = emulate(tone, dft, mel)[1], MO.wav2mel(tone).astype(np.float64)
    lin_g, lin_w = 10.0 ** got, 10.0 ** want
    assert (np.abs(lin_g - lin_w) <= 1e-5 * lin_w.max(axis=1, keepdims=True) + 1e-7).all()
    strong = lin_w > 1e-2 * lin_w.max(axis=1, keepdims=True)
    assert np.abs(got - want)[strong].max() < 1e-4
