"""Pins the numpy oracle (oracle/fluentspeech_oracle.py) against fixtures produced by the unmodified
reference (oracle/make_golden.py), and — when /root/reference is present — against the live reference."""
import numpy as np
import pytest

from conftest import golden, rel_l1
from oracle import fluentspeech_oracle as O
from oracle import refshim
from speech_editing_toolkit_b200 import schedule, synth


def test_schedule_matches_reference_buffers():
    k = golden("kat.npz")
    for S in (4, 8, 10, 100):
        sc = O.make_schedule(S)
        prod = schedule.diffusion_buffers(S)
        for name in ("betas", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped"):
            assert np.array_equal(sc[name], k[f"sched{S}_{name}"]), (S, name)
            assert np.array_equal(prod[name], k[f"sched{S}_{name}"]), (S, name)
    sc = O.make_schedule(100)
    # SURVEY appendix A known answers
    assert sc["posterior_mean_coef1"][0] == 1.0 and sc["posterior_mean_coef2"][0] == 0.0
    assert abs(sc["posterior_log_variance_clipped"][0] - (-46.0517)) < 1e-3
    assert abs(sc["betas"][0] - 0.00294) < 1e-5 and abs(sc["betas"][99] - 0.32306) < 1e-5


def test_integer_ops_bit_exact():
    k = golden("kat.npz")
    assert np.array_equal(O.f0_to_coarse(k["f0_in"]), k["f0_coarse"])
    assert np.array_equal(k["f0_coarse"], [1, 1, 14, 23, 69, 141, 185, 255, 255])
    assert np.array_equal(O.mel2token_to_dur(k["mel2token"], 5), k["mel2token_dur"])
    assert np.array_equal(O.expand_states(k["expand_h"], k["expand_idx"]), k["expand_out"])


def test_sinusoid_and_mish_known_answers():
    e = O.sinusoidal_pos_emb(np.array([1, 99]), 256)
    assert abs(e[0, 0] - 0.841471) < 1e-6 and abs(e[0, 128] - 0.540302) < 1e-6 and abs(e[0, 255] - 1.0) < 1e-6
    assert abs(e[0, 127] - 1.0e-4) < 1e-8 and abs(e[1, 0] - (-0.999207)) < 1e-5
    m = O.mish(np.array([-2, -1, 0, 1, 2], dtype=np.float32))
    assert np.allclose(m, [-0.2525015, -0.3034015, 0, 0.8650984, 1.9439590], atol=1e-6)


def test_diffnet_step_matches_reference_fixture():
    g = golden("diffnet_step.npz")
    sd = synth.denoiser_state_dict(int(g["seed"]))
    cond = synth.synthetic_cond(int(g["seed"]), int(g["B"]), int(g["T"])).transpose(0, 2, 1)
    x0 = O.diffnet_forward(sd, g["x"], g["t"], cond)
    assert np.abs(x0 - g["x0"]).max() < 2e-5
    # bf16-operand contract of the tensor-core kernels stays within the stated tolerance of the fp32 reference
    xb = O.diffnet_forward(sd, g["x"], g["t"], cond, gemm_dtype="bf16")
    assert rel_l1(xb, g["x0"]) < 1e-2


def test_sampling_loop_c1_matches_reference_fixture():
    g = golden("sample_c1.npz")
    seed, B, T, S = int(g["seed"]), int(g["B"]), int(g["T"]), int(g["S"])
    sd = synth.denoiser_state_dict(1234)
    cond = synth.synthetic_cond(seed, B, T).transpose(0, 2, 1)
    noise = synth.synthetic_noise(seed, S, B, T)
    mel, xs = O.sample_loop(sd, O.make_schedule(S), cond, noise, S, trace=True)
    assert np.abs(xs[0] - g["x_after_first"]).max() < 5e-5
    assert np.abs(xs[4] - g["x_after_fifth"]).max() < 1e-4
    assert np.abs(mel - g["mel_out"]).max() < 1e-4


def test_posterior_t0_returns_x0_exactly():
    sc = O.make_schedule(8)
    rs = np.random.RandomState(0)
    x0, xt, z = (rs.standard_normal((2, 80, 5)).astype(np.float32) for _ in range(3))
    out = O.posterior_sample(sc, x0, xt, np.array([0, 0]), z)
    assert np.array_equal(out, x0)


def test_hifigan_matches_reference_fixture():
    g = golden("hifigan_v1.npz")
    sd = synth.hifigan_state_dict(int(g["seed"]))
    wav = O.hifigan_forward(sd, O.HIFIGAN_V1, g["mel"].transpose(0, 2, 1))
    assert wav.shape == (1, 1, int(g["T"]) * 256)
    assert np.abs(wav[:, 0] - g["wav"]).max() < 2e-5


def test_weight_norm_fold_dim0():
    rs = np.random.RandomState(3)
    v = rs.standard_normal((4, 3, 5)).astype(np.float32); gmag = rs.uniform(0.5, 2, (4, 1, 1)).astype(np.float32)
    w = O.weight_norm_fold(v, gmag)
    assert np.allclose(np.sqrt((w.astype(np.float64) ** 2).sum((1, 2))), gmag[:, 0, 0], rtol=1e-6)


@pytest.mark.reference
@pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")
def test_oracle_against_live_reference_random_shapes():
    import torch
    hp = refshim.install("egs/spec_denoiser.yaml", overrides="timesteps=8")
    from modules.speech_editing.spec_denoiser.diffnet import DiffNet
    sd = synth.denoiser_state_dict(7)
    net = DiffNet(80).eval()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    rs = np.random.RandomState(5)
    for B, T in ((1, 1), (3, 37)):
        x = rs.standard_normal((B, 80, T)).astype(np.float32)
        cond = rs.standard_normal((B, 192, T)).astype(np.float32)
        t = rs.randint(0, 8, size=(B,)).astype(np.int64)
        with torch.no_grad():
            ref = net(torch.from_numpy(x)[:, None], torch.from_numpy(t), torch.from_numpy(cond))[:, 0].numpy()
        assert np.abs(O.diffnet_forward(sd, x, t, cond) - ref).max() < 2e-5


def test_mel_encoder_matches_reference_fixture():
    """MelEncoder.forward and `decoder_inp += out * tgt_nonpadding` (mel_encoder.py:15-19, spec_denoiser.py:162-164)."""
    from speech_editing_toolkit_b200 import synth
    g = golden("mel_encoder.npz")
    sd = synth.mel_encoder_state_dict(int(g["seed"]))
    ref, mask = synth.synthetic_ref_and_mask(int(g["seed"]), int(g["B"]), int(g["T"]))
    out = O.mel_encoder_forward(sd, ref * (1 - mask))
    assert np.abs(out - g["out"]).max() < 2e-5
    cond = g["decoder_inp"] + out * g["nonpad"]
    assert np.abs(cond - g["cond"]).max() < 2e-5
    assert np.array_equal(cond[1, 60:], g["decoder_inp"][1, 60:])      # padding frames keep decoder_inp bit for bit
