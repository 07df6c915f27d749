"""The torch CPU port used as bench.py's CPU baseline reproduces the reference fixtures."""
import numpy as np
import torch

from conftest import golden
from oracle import fluentspeech_oracle as O
from oracle import torch_port as P
from speech_editing_toolkit_b200 import schedule, synth


def test_port_diffnet_and_sampling_match_reference_fixtures():
    torch.set_num_threads(4)
    g = golden("diffnet_step.npz")
    p = P.to_torch(synth.denoiser_state_dict(1234))
    cond = torch.from_numpy(synth.synthetic_cond(int(g["seed"]), int(g["B"]), int(g["T"]))).transpose(1, 2)
    x0 = P.diffnet_forward(p, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), cond).numpy()
    assert np.abs(x0 - g["x0"]).max() < 2e-5
    g = golden("sample_c1.npz")
    seed, B, T, S = int(g["seed"]), int(g["B"]), int(g["T"]), int(g["S"])
    sched = {k: torch.from_numpy(v) for k, v in schedule.diffusion_buffers(S).items()}
    cond = torch.from_numpy(synth.synthetic_cond(seed, B, T)).transpose(1, 2)
    noise = torch.from_numpy(synth.synthetic_noise(seed, S, B, T))
    mel = P.sample_loop(p, sched, cond, S, noise=noise).numpy()
    assert np.abs(mel - g["mel_out"]).max() < 1e-4


def test_port_hifigan_matches_reference_fixture():
    g = golden("hifigan_v1.npz")
    p = P.to_torch(synth.hifigan_state_dict(int(g["seed"])))
    wav = P.hifigan_forward(p, O.HIFIGAN_V1, torch.from_numpy(g["mel"]).transpose(1, 2)).numpy()
    assert np.abs(wav[:, 0] - g["wav"]).max() < 2e-5
