"""Build-container only (skipped where /root/reference is absent): the numpy oracles against the LIVE reference modules on
shapes / seeds / flags other than the committed fixtures, so that the oracles are not fitted to one vector."""
import numpy as np
import pytest
import torch

from oracle import campnet_oracle as KO
from oracle import cond_encoder_oracle as CO
from oracle import refshim
from speech_editing_toolkit_b200 import synth

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")


def _t(sd):
    return {k: torch.from_numpy(v.copy()) for k, v in sd.items()}


@pytest.mark.parametrize("seed,B,T,fpp,pads", [(5, 3, 96, 4, [(0, 2), (2, 7)]), (6, 1, 50, 5, [])])
def test_cond_encoder_oracle_vs_live_fastspeech(seed, B, T, fpp, pads):
    hp = refshim.install("egs/spec_denoiser.yaml")
    from modules.speech_editing.spec_denoiser.fs import FastSpeech
    vocab = 40
    sd = synth.fastspeech_state_dict(seed, vocab)
    fs = FastSpeech(vocab, hp).eval()
    fs.load_state_dict(_t(sd), strict=False)
    batch = synth.synthetic_edit_batch(seed + 1, B, T, vocab=vocab, frames_per_phone=fpp)
    for item, n in pads:
        batch = synth.pad_edit_batch(batch, item, n, fpp)
    tb = {k: torch.from_numpy(v) for k, v in batch.items()}
    for flag in (False, True):
        with torch.no_grad():
            ret = fs(tb["txt_tokens"], tb["time_mel_masks"][:, :, None], tb["mel2ph"], tb["spk_embed"], tb["f0"], tb["uv"], skip_decoder=True,
                     infer=True, use_pred_pitch=flag)
        ours = CO.fastspeech_forward(sd, batch["txt_tokens"], batch["time_mel_masks"], batch["mel2ph"], batch["spk_embed"], batch["f0"],
                                     batch["uv"], use_pred_pitch=flag)
        for k in ("decoder_inp", "dur", "pitch_pred"):
            assert np.abs(ours[k] - ret[k].numpy()).max() < 3e-5, (k, flag)
        assert np.abs(ours["f0_denorm"] - ret["f0_denorm"].numpy()).max() < 2e-3
    # predicted alignment: LengthRegulator on the predicted durations (integer, bit-exact)
    with torch.no_grad():
        r2 = {}
        enc = fs.encoder(tb["txt_tokens"])
        style = fs.forward_style_embed(tb["spk_embed"], None)
        dur_inp = (enc + style) * (tb["txt_tokens"] > 0).float()[:, :, None]
        m2p = fs.forward_dur(dur_inp, tb["time_mel_masks"][:, :, None], tb["mel2ph"], tb["txt_tokens"], r2, use_pred_mel2ph=True)
    assert np.array_equal(CO.length_regulator(r2["dur"].numpy(), batch["txt_tokens"] == 0), m2p.numpy())


def test_cond_encoder_oracle_vs_live_fastspeech_libritts_config():
    """egs/spec_denoiser_libritts.yaml: use_pitch_embed false (no pitch branch, f0 / uv are None)."""
    hp = refshim.install("egs/spec_denoiser_libritts.yaml")
    assert hp["use_pitch_embed"] is False
    from modules.speech_editing.spec_denoiser.fs import FastSpeech
    vocab = 30
    ohp = {"use_pitch_embed": False}
    sd = synth.fastspeech_state_dict(9, vocab, ohp)
    fs = FastSpeech(vocab, hp).eval()
    missing, unexpected = fs.load_state_dict(_t(sd), strict=False)
    assert not unexpected and all(k.startswith(("decoder.", "mel_out.")) for k in missing)
    batch = synth.pad_edit_batch(synth.synthetic_edit_batch(10, 2, 64, vocab=vocab), 1, 2)
    tb = {k: torch.from_numpy(v) for k, v in batch.items()}
    with torch.no_grad():
        ret = fs(tb["txt_tokens"], tb["time_mel_masks"][:, :, None], tb["mel2ph"], tb["spk_embed"], None, None, skip_decoder=True, infer=True)
    ours = CO.fastspeech_forward(sd, batch["txt_tokens"], batch["time_mel_masks"], batch["mel2ph"], batch["spk_embed"], None, None, hp=ohp)
    assert np.abs(ours["decoder_inp"] - ret["decoder_inp"].numpy()).max() < 3e-5
    assert "pitch_pred" not in ret and "pitch_pred" not in ours


@pytest.mark.parametrize("seed,B,T,fpp,pads", [(21, 3, 140, 4, [(0, 6), (2, 1)]), (22, 1, 64, 8, [])])
def test_campnet_oracle_vs_live_campnet(seed, B, T, fpp, pads):
    hp = refshim.install("egs/campnet.yaml")
    from modules.speech_editing.campnet.campnet import CampNet
    vocab = 50
    sd = synth.campnet_state_dict(seed, vocab)
    net = CampNet(vocab, 100, hp).eval()
    net.load_state_dict(_t(sd), strict=False)
    b = synth.synthetic_campnet_batch(seed + 1, B, T, vocab=vocab, frames_per_phone=fpp, pad_items=pads)
    with torch.no_grad():
        ret = net(torch.from_numpy(b["txt_tokens"]), mels=torch.from_numpy(b["mels"]), time_mel_masks=torch.from_numpy(b["time_mel_masks"]), infer=True)
    ours = KO.campnet_forward(sd, b["txt_tokens"], b["mels"], b["time_mel_masks"])
    assert np.abs(ours["mel_out_coarse"] - ret["mel_out_coarse"].numpy()).max() < 2e-4
    assert np.abs(ours["mel_out_fine"] - ret["mel_out_fine"].numpy()).max() < 2e-4
    assert np.abs(ours["attn"] - ret["attn"].numpy()).max() < 2e-5


def test_integer_helpers_vs_live_reference_on_random_inputs():
    """mel2token_to_dur, f0_to_coarse o denorm_f0, LengthRegulator, expand_states: bit-exact against the reference's own functions
    over a few hundred random inputs (fixed seeds), including the boundaries (unvoiced frames, clamps, .5 durations, padding)."""
    refshim.install("egs/spec_denoiser.yaml")
    from modules.commons.nar_tts_modules import LengthRegulator
    from modules.tts.commons.align_ops import expand_states
    from utils.audio.align import mel2token_to_dur
    from utils.audio.pitch.utils import denorm_f0, f0_to_coarse
    from oracle import fluentspeech_oracle as O
    rs = np.random.RandomState(123)
    lr = LengthRegulator()
    for _ in range(40):
        B, Tt = int(rs.randint(1, 4)), int(rs.randint(1, 30))
        dur = np.round(rs.uniform(0, 6, (B, Tt)) * 2) / 2                           # many exact .5 values: round half to even
        dur = dur.astype(np.float32)
        pad = rs.uniform(size=(B, Tt)) < 0.15
        if dur[~pad].sum() == 0:
            continue
        want = lr(torch.from_numpy(dur), torch.from_numpy(pad)).numpy()
        got = CO.length_regulator(dur, pad)
        assert np.array_equal(got, want)
        T = got.shape[1]
        if T == 0:
            continue
        assert np.array_equal(O.mel2token_to_dur(got, Tt), mel2token_to_dur(torch.from_numpy(got), Tt).numpy())
        h = rs.standard_normal((B, Tt, 5)).astype(np.float32)
        assert np.array_equal(O.expand_states(h, got), expand_states(torch.from_numpy(h), torch.from_numpy(got)).numpy())
    for _ in range(40):
        n = int(rs.randint(1, 400))
        f0 = rs.uniform(4.5, 10.5, n).astype(np.float32)                             # log2 Hz: below 50 Hz .. above 900 Hz
        uv = (rs.uniform(size=n) < 0.3).astype(np.float32)
        pad = rs.uniform(size=n) < 0.1
        want_d = denorm_f0(torch.from_numpy(f0.copy()), torch.from_numpy(uv), pitch_padding=torch.from_numpy(pad))
        got_d = CO.denorm_f0(f0, uv, pad)
        assert np.abs(got_d - want_d.numpy()).max() < 2e-3
        assert np.array_equal(O.f0_to_coarse(want_d.numpy()), f0_to_coarse(want_d).numpy())   # same input -> same bins, bit-exact


def test_hifigan_resblock2_oracle_vs_live_generator():
    """ResBlock2 generators (hifigan.py:67-88; HiFi-GAN V3-style config): the oracle against the live HifiGanGenerator."""
    refshim.install("egs/spec_denoiser.yaml")
    from modules.vocoder.hifigan.hifigan import HifiGanGenerator
    from oracle import fluentspeech_oracle as O
    cfg = dict(upsample_rates=[8, 8, 4], upsample_kernel_sizes=[16, 16, 8], upsample_initial_channel=256, resblock="2",
               resblock_kernel_sizes=[3, 5, 7], resblock_dilation_sizes=[[1, 2], [2, 6], [3, 12]])
    sd = synth.hifigan_state_dict(17, cfg)
    gen = HifiGanGenerator(dict(cfg)).eval()
    gen.load_state_dict(_t(sd), strict=True)
    mel = np.clip(np.random.RandomState(18).standard_normal((2, 19, 80)) * 1.5 - 3.0, -6, 1.5).astype(np.float32)
    with torch.no_grad():
        want = gen(torch.from_numpy(mel).transpose(1, 2)).numpy()
    ours = O.hifigan_forward(sd, cfg, mel.transpose(0, 2, 1))
    assert ours.shape == want.shape == (2, 1, 19 * 256)
    assert np.abs(ours - want).max() < 2e-5
