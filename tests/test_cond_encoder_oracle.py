"""Pins oracle/cond_encoder_oracle.py (numpy restatement of FastSpeech.forward(skip_decoder=True), fs.py:83-189) against
tests/golden/cond_encoder.npz — outputs of the unmodified reference on a ragged batch (oracle/make_golden.py cond_encoder)."""
import numpy as np
import pytest

from conftest import golden, rel_l1
from oracle import cond_encoder_oracle as CO
from oracle import fluentspeech_oracle as O
from speech_editing_toolkit_b200 import synth


def _case():
    g = golden("cond_encoder.npz")
    seed, B, T, vocab = int(g["seed"]), int(g["B"]), int(g["T"]), int(g["vocab"])
    batch = synth.pad_edit_batch(synth.synthetic_edit_batch(seed, B, T, vocab=vocab), item=1, n_tokens=3)
    return g, synth.fastspeech_state_dict(seed, vocab), batch


def test_fixture_is_not_vacuous():
    g, _, batch = _case()
    assert (batch["txt_tokens"][1, -3:] == 0).all() and (batch["mel2ph"][1, -24:] == 0).all()       # ragged second item
    assert np.unique(g["pitch_predpitch"]).size > 20 and np.unique(g["mel2ph_pred"]).size > 5
    assert np.abs(g["decoder_inp"][1, -24:]).max() == 0.0                                            # padded frames are zero
    assert np.abs(g["decoder_inp"] - g["decoder_inp_predpitch"]).max() > 0.1


def test_text_encoder_and_style_match_reference():
    g, sd, batch = _case()
    enc = CO.text_encoder(sd, batch["txt_tokens"])
    assert np.abs(enc - g["encoder_out"]).max() < 2e-5
    assert np.abs(enc[1, -3:]).max() == 0.0
    assert np.abs(CO.style_embed(sd, batch["spk_embed"]) - g["style_embed"]).max() < 1e-6


@pytest.mark.parametrize("use_pred_pitch", [False, True])
def test_fastspeech_forward_matches_reference(use_pred_pitch):
    g, sd, batch = _case()
    sfx = "_predpitch" if use_pred_pitch else ""
    ret = CO.fastspeech_forward(sd, batch["txt_tokens"], batch["time_mel_masks"], batch["mel2ph"], batch["spk_embed"], batch["f0"],
                                batch["uv"], use_pred_pitch=use_pred_pitch)
    assert np.array_equal(ret["mel2ph"], g["mel2ph" + sfx])
    assert np.array_equal(ret["pitch"], g["pitch" + sfx])                                            # integer bins: bit-exact
    assert np.abs(ret["dur"] - g["dur" + sfx]).max() < 2e-5
    assert np.abs(ret["pitch_pred"] - g["pitch_pred" + sfx]).max() < 2e-5
    assert np.abs(ret["f0_denorm"] - g["f0_denorm" + sfx]).max() < 2e-3                              # Hz, values up to 900
    assert np.abs(ret["f0_denorm_pred"] - g["f0_denorm_pred" + sfx]).max() < 2e-3
    assert np.abs(ret["decoder_inp"] - g["decoder_inp" + sfx]).max() < 2e-5


def test_forward_dur_with_masked_dur_and_length_regulator():
    """The inference script's call (inference/tts/spec_denoiser.py:84-98): explicit masked_dur, predicted mel2ph."""
    g, sd, batch = _case()
    txt = batch["txt_tokens"]
    dur_inp = (g["encoder_out"] + g["style_embed"]) * (txt > 0)[:, :, None].astype(np.float32)
    dur_inp = (dur_inp + sd["dur_embed.weight"][g["masked_dur_in"]]).astype(np.float32)
    dur = CO.duration_predictor(sd, dur_inp, txt == 0)
    assert np.abs(dur - g["dur_masked_dur"]).max() < 2e-5
    assert np.array_equal(CO.length_regulator(g["dur_masked_dur"], txt == 0), g["mel2ph_pred"])      # integer: bit-exact


def test_integer_helpers_known_answers():
    assert np.array_equal(CO.length_regulator(np.array([[2.4, 0.6, 3.5]], dtype=np.float32)), [[1, 1, 2, 3, 3, 3, 3]])
    assert np.array_equal(CO.length_regulator(np.array([[0.5, 1.5, 2.5]], dtype=np.float32)), [[2, 2, 3, 3]])           # half to even
    d = CO.denorm_f0(np.array([5, 6.64, 7.78, 10.5], dtype=np.float32), np.array([0, 0, 1, 0], dtype=np.float32))
    assert np.allclose(d, [50.0, 99.73306, 0.0, 900.0], atol=1e-3)
    m = CO.masked_dur_gt(np.array([[1, 1, 2, 2, 2, 4, 0, 0]]), np.array([[0, 0, 0, 1, 1, 0, 0, 0]], dtype=np.float32),
                         np.array([[5, 6, 7, 8, 0]]))
    assert np.array_equal(m, [[2, 1, 0, 1, 0]])


def test_bf16_contract_within_stated_tolerance():
    g, sd, batch = _case()
    ret = CO.fastspeech_forward(sd, batch["txt_tokens"], batch["time_mel_masks"], batch["mel2ph"], batch["spk_embed"], batch["f0"],
                                batch["uv"], use_pred_pitch=False, gemm_dtype="bf16")
    assert rel_l1(ret["decoder_inp"], g["decoder_inp"]) < 1e-2
    assert rel_l1(ret["dur"], g["dur"]) < 1e-2


def test_e2e_fixture_reproduced_by_the_oracles():
    """GaussianDiffusion.forward(infer=True) of the reference = condition encoder + MelEncoder + sampling loop."""
    g = golden("fluentspeech_e2e.npz")
    seed, B, T, S, L, vocab = (int(g[k]) for k in ("seed", "B", "T", "S", "layers", "vocab"))
    batch = synth.pad_edit_batch(synth.synthetic_edit_batch(seed, B, T, vocab=vocab), item=1, n_tokens=3)
    ret = CO.fastspeech_forward(synth.fastspeech_state_dict(seed, vocab), batch["txt_tokens"], batch["time_mel_masks"], batch["mel2ph"],
                                batch["spk_embed"], batch["f0"], batch["uv"], use_pred_pitch=True)
    m = batch["time_mel_masks"][:, :, None]
    nonpad = (batch["mel2ph"] > 0).astype(np.float32)[:, :, None]
    cond = ret["decoder_inp"] + O.mel_encoder_forward(synth.mel_encoder_state_dict(seed), batch["ref_mels"] * (1 - m)) * nonpad
    assert np.abs(cond - g["decoder_inp"]).max() < 5e-5
    noise = synth.synthetic_noise(seed + 5, S, B, T)
    mel = O.sample_loop(synth.denoiser_state_dict(seed, layers=L), O.make_schedule(S), cond.transpose(0, 2, 1), noise, S)
    assert np.abs(mel - g["mel_out"]).max() < 2e-4
