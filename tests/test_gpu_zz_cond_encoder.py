"""Parity of the native condition encoder (fse_cond_*; FastSpeechB200) on the GPU against
  * tests/golden/cond_encoder.npz / fluentspeech_e2e.npz — outputs of the unmodified reference FastSpeech / GaussianDiffusion
    (oracle/make_golden.py cond_encoder), and
  * oracle/cond_encoder_oracle.py on a larger ragged multi-tile batch.
Stated tolerances: FSE_MODE_SIMT_F32 max-abs <= 2e-4 on every float output; FSE_MODE_TC_BF16 relative L1 <= 2e-2 vs the fp32
reference and <= 8e-3 vs the bf16-operand oracle; integer outputs (masked durations, length regulator, mel2ph) bit-exact;
pitch bins bit-exact in fp32 mode given the reference's own pitch prediction is not involved (use_pred_pitch=False).
(The file sorts last on purpose: it is the newest GPU surface of the round.)"""
import numpy as np
import pytest

from conftest import golden, rel_l1

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL_F32_ABS = 2e-4
HP = dict(audio_num_mel_bins=80, hidden_size=192, residual_layers=4, residual_channels=256, dilation_cycle_length=1,
          timesteps=4, timescale=1, diff_loss_type="l1", spec_min=[], spec_max=[], keep_bins=80, schedule_type="vpsde",
          diff_decoder_type="wavenet_b200", b200_mode="tc_bf16")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _case():
    from speech_editing_toolkit_b200 import synth
    g = golden("cond_encoder.npz")
    seed, B, T, vocab = int(g["seed"]), int(g["B"]), int(g["T"]), int(g["vocab"])
    batch = synth.pad_edit_batch(synth.synthetic_edit_batch(seed, B, T, vocab=vocab), item=1, n_tokens=3)
    return g, synth.fastspeech_state_dict(seed, vocab), batch, vocab


def _module(sd, vocab, mode):
    from speech_editing_toolkit_b200.modules import FastSpeechB200
    fs = FastSpeechB200(vocab, dict(HP, b200_mode=mode)).cuda()
    fs.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    return fs


@pytest.mark.parametrize("mode", ["simt_f32", "tc_bf16"])
def test_text_encoder_style_and_duration_vs_reference_fixture(lib_built, mode):
    _need_gpu()
    g, sd, batch, vocab = _case()
    fs = _module(sd, vocab, mode)
    txt = cu(batch["txt_tokens"])
    enc = fs.encoder(txt).cpu().numpy()
    style = fs.forward_style_embed(cu(batch["spk_embed"])).cpu().numpy()
    assert enc.shape == g["encoder_out"].shape and style.shape == g["style_embed"].shape
    assert np.abs(style - g["style_embed"]).max() < 1e-5
    assert np.abs(enc[1, -3:]).max() == 0.0                                          # padded tokens stay exactly zero
    # the inference script's forward_dur call: explicit masked_dur, predicted mel2ph (inference/tts/spec_denoiser.py:84-98)
    dur_inp = (cu(g["encoder_out"]) + cu(g["style_embed"])) * (txt > 0).float()[:, :, None]
    ret = {}
    mel2ph_pred = fs.forward_dur(dur_inp, cu(batch["time_mel_masks"]), cu(batch["mel2ph"]), txt, ret, masked_dur=cu(g["masked_dur_in"]),
                                 use_pred_mel2ph=True)
    dur = ret["dur"].cpu().numpy()
    if mode == "simt_f32":
        assert np.abs(enc - g["encoder_out"]).max() < TOL_F32_ABS
        assert np.abs(dur - g["dur_masked_dur"]).max() < TOL_F32_ABS
        assert np.array_equal(mel2ph_pred.cpu().numpy(), g["mel2ph_pred"])            # integer: bit-exact
    else:
        assert rel_l1(enc, g["encoder_out"]) < 2e-2
        assert rel_l1(dur, g["dur_masked_dur"]) < 2e-2
    # the length regulator alone on the reference's own durations: bit-exact in every mode
    lr = fs.engine().length_regulate(cu(g["dur_masked_dur"]), txt).cpu().numpy()
    assert np.array_equal(lr, g["mel2ph_pred"])


@pytest.mark.parametrize("mode", ["simt_f32", "tc_bf16"])
@pytest.mark.parametrize("use_pred_pitch", [False, True])
def test_fastspeech_forward_vs_reference_fixture(lib_built, mode, use_pred_pitch):
    _need_gpu()
    g, sd, batch, vocab = _case()
    fs = _module(sd, vocab, mode)
    sfx = "_predpitch" if use_pred_pitch else ""
    ret = fs(cu(batch["txt_tokens"]), cu(batch["time_mel_masks"])[:, :, None], cu(batch["mel2ph"]), cu(batch["spk_embed"]), cu(batch["f0"]),
             cu(batch["uv"]), skip_decoder=True, infer=True, use_pred_pitch=use_pred_pitch)
    out = {k: v.cpu().numpy() for k, v in ret.items()}
    assert set(out) == {"decoder_inp", "dur", "mel2ph", "pitch_pred", "f0_denorm", "f0_denorm_pred"}
    assert np.array_equal(out["mel2ph"], g["mel2ph" + sfx])
    assert np.abs(out["decoder_inp"][1, -24:]).max() == 0.0                           # padded frames exactly zero
    if mode == "simt_f32":
        for k in ("decoder_inp", "dur", "pitch_pred"):
            assert np.abs(out[k] - g[k + sfx]).max() < TOL_F32_ABS, k
        for k in ("f0_denorm", "f0_denorm_pred"):
            assert np.abs(out[k] - g[k + sfx]).max() < 0.05, k                        # Hz (values up to 900)
    else:
        assert rel_l1(out["decoder_inp"], g["decoder_inp" + sfx]) < (4e-2 if use_pred_pitch else 2e-2)   # a flipped pitch bin swaps an embedding row
        assert rel_l1(out["dur"], g["dur" + sfx]) < 2e-2
        assert rel_l1(out["pitch_pred"], g["pitch_pred" + sfx]) < 2e-2


@pytest.mark.parametrize("mode", ["simt_f32", "tc_bf16"])
def test_integer_ops_bit_exact(lib_built, mode):
    """masked durations (scatter_add histogram), pitch bins of given f0 (f0_to_coarse o denorm_f0) and the gather by mel2ph."""
    _need_gpu()
    from oracle import cond_encoder_oracle as CO
    g, sd, batch, vocab = _case()
    fs = _module(sd, vocab, mode)
    eng = fs.engine()
    txt, mel2ph, mask = cu(batch["txt_tokens"]), cu(batch["mel2ph"]), cu(batch["time_mel_masks"])
    md = eng.masked_dur(mel2ph, mask, txt).cpu().numpy()
    assert md.dtype == np.int64 and np.array_equal(md, CO.masked_dur_gt(batch["mel2ph"], batch["time_mel_masks"], batch["txt_tokens"]))
    assert np.array_equal(eng.masked_dur(mel2ph, None, txt).cpu().numpy(),
                          CO.masked_dur_gt(batch["mel2ph"], np.zeros_like(batch["time_mel_masks"]), batch["txt_tokens"]))
    enc = cu(g["encoder_out"])
    out = eng.frames(enc, cu(g["style_embed"][:, 0]), mel2ph, mask, cu(batch["f0"]), cu(batch["uv"]), False)
    assert np.array_equal(out["pitch"].cpu().numpy(), g["pitch"])                      # bins of the given f0/uv: bit-exact
    # known answers of SURVEY appendix A through the same kernels (f0 given in log2 Hz, uv = 0)
    hz = np.array([[50, 80, 100, 220, 440, 600, 900, 1200]], dtype=np.float32)
    one = eng.frames(enc[:1], None, torch.ones(1, 8, dtype=torch.int64, device="cuda"), None, cu(np.log2(hz)), cu(np.zeros_like(hz)), False)
    assert np.array_equal(one["pitch"].cpu().numpy(), [[1, 14, 23, 69, 141, 185, 255, 255]])
    lr = eng.length_regulate(cu(np.array([[2.4, 0.6, 3.5], [0.5, 1.5, 2.5]], dtype=np.float32))).cpu().numpy()
    assert np.array_equal(lr, [[1, 1, 2, 3, 3, 3, 3], [2, 2, 3, 3, 0, 0, 0]])           # round half to even; zero past the end


def test_multi_tile_ragged_batch_vs_oracle(lib_built):
    """3 items x 330 frames / 82 tokens (three 128-row tiles per item in the frame branch), ragged tails, against the numpy
    oracle in both arithmetic contracts."""
    _need_gpu()
    from oracle import cond_encoder_oracle as CO
    from speech_editing_toolkit_b200 import synth
    vocab, B, T = 60, 3, 330
    sd = synth.fastspeech_state_dict(77, vocab)
    batch = synth.synthetic_edit_batch(78, B, T, vocab=vocab, frames_per_phone=4)
    batch = synth.pad_edit_batch(synth.pad_edit_batch(batch, 1, 20, 4), 2, 5, 4)
    args = (batch["txt_tokens"], batch["time_mel_masks"], batch["mel2ph"], batch["spk_embed"], batch["f0"], batch["uv"])
    ref32 = CO.fastspeech_forward(sd, *args, use_pred_pitch=False)
    refbf = CO.fastspeech_forward(sd, *args, use_pred_pitch=False, gemm_dtype="bf16")
    for mode in ("simt_f32", "tc_bf16"):
        fs = _module(sd, vocab, mode)
        ret = fs(cu(batch["txt_tokens"]), cu(batch["time_mel_masks"])[:, :, None], cu(batch["mel2ph"]), cu(batch["spk_embed"]),
                 cu(batch["f0"]), cu(batch["uv"]), skip_decoder=True, infer=True, use_pred_pitch=False)
        out = {k: v.cpu().numpy() for k, v in ret.items()}
        if mode == "simt_f32":
            for k in ("decoder_inp", "dur", "pitch_pred"):
                assert np.abs(out[k] - ref32[k]).max() < TOL_F32_ABS, k
        else:
            for k in ("decoder_inp", "dur", "pitch_pred"):
                assert rel_l1(out[k], ref32[k]) < 2e-2, k
                assert rel_l1(out[k], refbf[k]) < 8e-3, k
        assert fs.engine().launches > 0


def test_whole_model_text_to_mel_vs_reference_fixture(lib_built):
    """GaussianDiffusionB200.forward(infer=True) with every sub-module native (FastSpeechB200 + MelEncoderB200 + DiffNetB200)
    against the unmodified reference GaussianDiffusion.forward on the same weights, inputs and injected noise."""
    _need_gpu()
    from speech_editing_toolkit_b200 import plugin, synth
    from speech_editing_toolkit_b200.modules import FastSpeechB200, MelEncoderB200
    g = golden("fluentspeech_e2e.npz")
    seed, B, T, S, L, vocab = (int(g[k]) for k in ("seed", "B", "T", "S", "layers", "vocab"))
    batch = synth.pad_edit_batch(synth.synthetic_edit_batch(seed, B, T, vocab=vocab), item=1, n_tokens=3)
    noise = synth.synthetic_noise(seed + 5, S, B, T)
    for mode, tol in (("simt_f32", None), ("tc_bf16", 3e-2), ("tc_tf32", 4e-3)):
        model = plugin.build_diffusion(dict(HP, timesteps=S, residual_layers=L, b200_mode=mode), phone_encoder=list(range(vocab)))
        assert isinstance(model.fs, FastSpeechB200) and isinstance(model.mel_encoder, MelEncoderB200)
        model.fs.load_state_dict({k: torch.from_numpy(v) for k, v in synth.fastspeech_state_dict(seed, vocab).items()}, strict=False)
        model.mel_encoder.load_state_dict({k: torch.from_numpy(v) for k, v in synth.mel_encoder_state_dict(seed).items()})
        model.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in synth.denoiser_state_dict(seed, layers=L).items()})
        model = model.cuda().eval()
        ret = model(cu(batch["txt_tokens"]), cu(batch["time_mel_masks"])[:, :, None], cu(batch["mel2ph"]), cu(batch["spk_embed"]),
                    cu(batch["ref_mels"]), cu(batch["f0"]), cu(batch["uv"]), infer=True, use_pred_pitch=True, noise=cu(noise))
        mel, cond = ret["mel_out"].cpu().numpy(), ret["decoder_inp"].cpu().numpy()
        assert mel.shape == g["mel_out"].shape
        if tol is None:
            assert np.abs(cond - g["decoder_inp"]).max() < 5e-4
            assert np.abs(mel - g["mel_out"]).max() < 2e-3
        else:
            print(f"[margin] text -> mel {mode}: decoder_inp rel-L1 {rel_l1(cond, g['decoder_inp']):.3e}, mel rel-L1 {rel_l1(mel, g['mel_out']):.3e}")
            assert rel_l1(cond, g["decoder_inp"]) < tol
            assert rel_l1(mel, g["mel_out"]) < tol


def test_libritts_config_without_pitch_embed(lib_built):
    """egs/spec_denoiser_libritts.yaml:169 (BASELINE configs[1]) sets use_pitch_embed: false: f0 / uv are None, forward_pitch is
    skipped (fs.py:97-99) and decoder_inp = (expand_states(encoder_out, mel2ph) + style) * tgt_nonpadding."""
    _need_gpu()
    from oracle import cond_encoder_oracle as CO
    from speech_editing_toolkit_b200 import synth
    from speech_editing_toolkit_b200.modules import FastSpeechB200
    vocab, B, T = 50, 2, 200
    hp = {"use_pitch_embed": False}
    sd = synth.fastspeech_state_dict(31, vocab, hp)
    assert not any(k.startswith("pitch_") for k in sd)
    batch = synth.pad_edit_batch(synth.synthetic_edit_batch(32, B, T, vocab=vocab), item=1, n_tokens=4)
    ref = CO.fastspeech_forward(sd, batch["txt_tokens"], batch["time_mel_masks"], batch["mel2ph"], batch["spk_embed"], None, None, hp=hp)
    for mode in ("simt_f32", "tc_bf16"):
        fs = FastSpeechB200(vocab, dict(HP, b200_mode=mode, use_pitch_embed=False)).cuda()
        fs.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        ret = fs(cu(batch["txt_tokens"]), cu(batch["time_mel_masks"])[:, :, None], cu(batch["mel2ph"]), cu(batch["spk_embed"]), None, None,
                 skip_decoder=True, infer=True)
        assert set(ret) == {"decoder_inp", "dur", "mel2ph"}
        out = {k: v.cpu().numpy() for k, v in ret.items()}
        assert np.abs(out["decoder_inp"][1, -32:]).max() == 0.0
        if mode == "simt_f32":
            assert np.abs(out["decoder_inp"] - ref["decoder_inp"]).max() < TOL_F32_ABS
            assert np.abs(out["dur"] - ref["dur"]).max() < TOL_F32_ABS
        else:
            assert rel_l1(out["decoder_inp"], ref["decoder_inp"]) < 2e-2
            assert rel_l1(out["dur"], ref["dur"]) < 2e-2


def test_region_surgery_kernels_vs_reference_code_fixture(lib_built):
    """fse_edit_prepare / fse_edit_plan / fse_edit_assemble (inference/tts/spec_denoiser.py:88-131) against the tensors recorded from
    the reference's own forward_model code; the per-item logic is already checked on the host (tests/test_edit_region_core.py)."""
    _need_gpu()
    from speech_editing_toolkit_b200 import engine as E
    from test_edit_region_oracle import case
    g = golden("edit_region.npz")
    for i in range(3):
        item, want = case(g, i)
        b = lambda k, dt=None: cu(item[k] if dt is None else item[k].astype(dt))[None]
        regions = cu(np.array([[*item["words_region"][0], *item["edited_words_region"][0]]], dtype=np.int64))
        md, mm, mo = E.edit_prepare(b("mel2ph"), b("mel2word"), b("ph2word"), b("dur"), regions, len(item["edited_ph2word"]))
        assert np.array_equal(md[0].cpu().numpy(), want["masked_dur"]) and np.array_equal(mm[0].cpu().numpy(), want["masked_mel2ph"])
        assert np.array_equal(mo[0].cpu().numpy(), want["time_mel_masks_orig"].astype(np.float32))
        out = E.edit_assemble(b("mel2ph"), b("mel2word"), b("edited_ph2word"), cu(want["edited_mel2ph_pred"])[None], regions, b("mel"), b("f0"), b("uv"))
        assert np.array_equal(out["mel2ph"][0].cpu().numpy(), want["mel2ph"])
        assert np.array_equal(out["time_mel_masks"][0].cpu().numpy(), want["time_mel_masks"][:, 0])
        for k in ("ref_mels", "f0", "uv"):
            assert np.array_equal(out[k][0].cpu().numpy(), want[k]), k


def test_edit_forward_end_to_end_shapes(lib_built):
    """SpecDenoiserInferB200.forward_model on the reference's `sample` batch: edited text -> durations -> surgery -> model -> vocoder."""
    _need_gpu()
    from speech_editing_toolkit_b200 import plugin, synth
    from speech_editing_toolkit_b200.vocoder import HifiGANB200
    item = synth.synthetic_edit_item(3, n_words=9, edit_span=(3, 4), new_span_phones=(2, 3, 1))
    hp = dict(HP, timesteps=4, residual_layers=4, vocoder_ckpt="", work_dir="")
    inf = plugin.SpecDenoiserInferB200(hp, vocoder=HifiGANB200(state_dict=synth.hifigan_state_dict(1)), phone_encoder=range(80))
    t_ = lambda sd: {k: torch.from_numpy(v) for k, v in sd.items()}
    inf.model.fs.load_state_dict(t_(synth.fastspeech_state_dict(1, 80)), strict=False)
    inf.model.mel_encoder.load_state_dict(t_(synth.mel_encoder_state_dict(1)))
    inf.model.denoise_fn.load_state_dict(t_(synth.denoiser_state_dict(1, layers=4)))
    sample = {k: torch.from_numpy(item[k])[None] for k in ("mel", "mel2ph", "mel2word", "dur", "ph2word", "edited_ph2word", "f0", "uv", "spk_embed")}
    sample.update(edited_txt_tokens=torch.from_numpy(item["edited_ph_token"])[None], words_region=item["words_region"],
                  edited_words_region=item["edited_words_region"], seed=5)
    wav, mel, aux = inf.forward_model(sample)
    Tn = int(aux["plan"][0, 0])
    assert mel.shape == (1, Tn, 80) and wav.shape == (1, Tn * 256) and torch.isfinite(wav).all()
    m = aux["time_mel_masks"][0].bool()
    assert torch.equal(mel[0][~m], aux["ref_mels"][0][~m])                      # unedited frames: the original mel, bit for bit


def test_mel_frontend_vs_oracle(lib_built):
    """fse_mel_frontend_forward (wav -> log10-mel, utils/audio/__init__.py:34-81) against oracle/mel_frontend_oracle.py: noise-like
    signals within 1e-4 in log10; a pure tone within 1e-5 of the frame's strongest band in the linear domain (fp32 floor)."""
    _need_gpu()
    from oracle import mel_frontend_oracle as MO
    from speech_editing_toolkit_b200 import audio
    rs = np.random.RandomState(3)
    for n in (256 * 37, 256 * 5 + 77, 300):
        wav = (rs.standard_normal(n) * 0.2).astype(np.float32)
        res = audio.wav2spec(wav, fmin=55, fmax=7600, sample_rate=22050)
        want = MO.wav2mel(wav)
        assert res["mel"].shape == want.shape == (1 + n // 256, 80) and len(res["wav"]) == want.shape[0] * 256
        assert np.abs(res["mel"] - want).max() < 1e-4
    t = np.arange(22050) / 22050
    tone = (0.4 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32)
    got, want = audio.wav2spec(tone, fmin=55, fmax=7600)["mel"].astype(np.float64), MO.wav2mel(tone).astype(np.float64)
    assert (np.abs(10 ** got - 10 ** want) <= 1e-5 * (10 ** want).max(axis=1, keepdims=True) + 1e-7).all()
    assert audio.wav2spec(np.zeros(2560, dtype=np.float32))["mel"].max() == -6.0
    # the fixture made without the oracle (scipy.signal.stft + a Slaney filterbank checked against librosa's published constants)
    g = golden("mel_frontend.npz")
    got = audio.wav2spec(g["wav"], fmin=55, fmax=7600, sample_rate=22050)["mel"]
    assert got.shape == g["mel"].shape
    lin_g, lin_w = 10 ** got.astype(np.float64), 10 ** g["mel"].astype(np.float64)
    assert (np.abs(lin_g - lin_w) <= 1e-5 * lin_w.max(axis=1, keepdims=True) + 1e-7).all()
    print(f"[margin] mel front-end vs independent fixture: max |log10 diff| {np.abs(got - g['mel']).max():.3e}")
    # ... and torchaudio's librosa-compatible MelSpectrogram (slaney / slaney, centre zero padding) evaluated on the same samples
    try:
        import torchaudio
    except ImportError:
        return
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, win_length=1024, hop_length=256, f_min=55.0, f_max=7600.0,
                                              n_mels=80, power=1.0, norm="slaney", mel_scale="slaney", center=True, pad_mode="constant")
    lin_t = ms(torch.from_numpy(g["wav"])).T.numpy().astype(np.float64)
    assert (np.abs(lin_g - np.maximum(lin_t, 1e-6)) <= 1e-5 * lin_t.max(axis=1, keepdims=True) + 1e-7).all()
    print(f"[margin] mel front-end vs torchaudio MelSpectrogram: max |log10 diff| {np.abs(got - np.log10(np.maximum(lin_t, 1e-6))).max():.3e}")
