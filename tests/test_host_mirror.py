"""Host-side mirror of the reference seams (no GPU): config surface, state_dict layout, registries,
checkpoint layout."""
import os

import numpy as np
import pytest
import torch
import yaml

from oracle import refshim
from speech_editing_toolkit_b200 import ckpt, hparams as hp_mod, synth

needs_ref = pytest.mark.skipif(not refshim.available(), reason="reference tree not present (GPU box)")

HP = dict(audio_num_mel_bins=80, hidden_size=192, residual_layers=20, residual_channels=256, dilation_cycle_length=1,
          timesteps=8, timescale=1, diff_loss_type="l1", spec_min=[], spec_max=[], keep_bins=80, schedule_type="vpsde",
          diff_decoder_type="wavenet")


def test_set_hparams_base_chain_overrides_and_types(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    (tmp_path / "egs").mkdir()
    (tmp_path / "egs" / "base.yaml").write_text("a: 1\nlr: 0.1\nlst: [1, 2]\nflag: false\nsub: {x: 1, y: 2}\nname: base\n")
    (tmp_path / "egs" / "mid.yaml").write_text("base_config: ./base.yaml\na: 2\nsub: {y: 3}\n")
    (tmp_path / "egs" / "top.yaml").write_text("base_config:\n  - egs/mid.yaml\nname: top\n")
    cfg = hp_mod.set_hparams("egs/top.yaml", hparams_str="lr=0.5,lst=[3 4 5],flag=True,sub.x=7,name=z", print_hparams=False)
    assert cfg["a"] == 2 and cfg["name"] == "z" and cfg["lr"] == 0.5 and cfg["lst"] == [3, 4, 5] and cfg["flag"] is True
    assert cfg["sub"] == {"x": 7, "y": 3}
    assert hp_mod.hparams["a"] == 2 and hp_mod.hparams["work_dir"] == "" and hp_mod.hparams["infer"] is False
    # exp_name: config is saved once and wins over the yaml afterwards unless reset (reference: hparams.py:93-107,130-133)
    hp_mod.set_hparams("egs/top.yaml", exp_name="e1", hparams_str="a=5", print_hparams=False)
    assert yaml.safe_load(open("checkpoints/e1/config.yaml"))["a"] == 5
    again = hp_mod.set_hparams("egs/top.yaml", exp_name="e1", print_hparams=False, global_hparams=False)
    assert again["a"] == 5 and again["work_dir"] == "checkpoints/e1"
    # unknown keys are a KeyError as in the reference (hparams.py:94-105), except the drop-in's own b200_* switches
    with pytest.raises(KeyError):
        hp_mod.set_hparams("egs/top.yaml", hparams_str="nope=1", print_hparams=False, global_hparams=False)
    new = hp_mod.set_hparams("egs/top.yaml", hparams_str="b200_frames=128,b200_mode=tc_bf16", print_hparams=False, global_hparams=False)
    assert new["b200_frames"] == 128 and new["b200_mode"] == "tc_bf16"


@needs_ref
@pytest.mark.parametrize("cfg,ov", [("egs/spec_denoiser.yaml", "timesteps=100"), ("egs/spec_denoiser_libritts.yaml", "timesteps=100,lr=0.001"),
                                    ("egs/campnet.yaml", "max_sentences=64")])
def test_reference_yaml_loads_unchanged_and_matches_reference_loader(cfg, ov, monkeypatch):
    monkeypatch.chdir(refshim.REF_ROOT)
    ours = hp_mod.set_hparams(cfg, hparams_str=ov, print_hparams=False, global_hparams=False)
    ref_hp = refshim.install(cfg, overrides=ov)
    assert ours == dict(ref_hp)
    if "timesteps" in ov:
        assert ours["timesteps"] == 100 and ours["residual_layers"] == 20


def test_diffnet_state_dict_layout_and_checkpoint_roundtrip(tmp_path):
    from speech_editing_toolkit_b200.modules import DiffNetB200, GaussianDiffusionB200
    net = DiffNetB200(80, HP)
    sd = synth.denoiser_state_dict(3)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: v.shape for k, v in sd.items()}
    assert float(net.output_projection.weight.abs().sum()) == 0.0            # zero init like the reference
    model = GaussianDiffusionB200(None, 80, net, timesteps=8, hparams=HP)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    # trainer.py:457-470 layout, load_ckpt(model, dir, 'model')
    torch.save({"state_dict": {"model": model.state_dict()}, "global_step": 7}, tmp_path / "model_ckpt_steps_7.ckpt")
    torch.save({"state_dict": {"model": {}}, "global_step": 3}, tmp_path / "model_ckpt_steps_3.ckpt")
    fresh = GaussianDiffusionB200(None, 80, DiffNetB200(80, HP), timesteps=8, hparams=HP)
    path = ckpt.load_ckpt(fresh, str(tmp_path), "model")
    assert path.endswith("model_ckpt_steps_7.ckpt")
    for k, v in sd.items():
        assert np.array_equal(fresh.denoise_fn.state_dict()[k].numpy(), v)
    assert np.array_equal(fresh.posterior_mean_coef1.numpy(), model.posterior_mean_coef1.numpy())
    with pytest.raises(RuntimeError):                                # no condition encoder was given: both branches refuse
        fresh(None, None, None, None, None, None, None, infer=False)


@needs_ref
def test_state_dict_keys_match_live_reference_modules():
    refshim.install("egs/spec_denoiser.yaml")
    from modules.speech_editing.spec_denoiser.diffnet import DiffNet
    from modules.speech_editing.spec_denoiser.spec_denoiser import GaussianDiffusion
    from speech_editing_toolkit_b200.modules import DiffNetB200, GaussianDiffusionB200
    ref_net = DiffNet(80)
    ours = DiffNetB200(80, HP)
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == {k: tuple(v.shape) for k, v in ref_net.state_dict().items()}
    ref = GaussianDiffusion(list(range(80)), 80, ref_net, timesteps=8, time_scale=1, loss_type="l1", spec_min=[], spec_max=[])
    wrapped = GaussianDiffusionB200.from_reference(ref)
    rsd, osd = ref.state_dict(), wrapped.state_dict()
    assert set(rsd) == set(osd)
    for k in rsd:
        if not k.startswith(("fs.", "mel_encoder.", "denoise_fn.")):
            assert torch.equal(rsd[k], osd[k]), k
    from speech_editing_toolkit_b200.modules import FastSpeechB200, MelEncoderB200
    assert isinstance(wrapped.fs, FastSpeechB200)                    # the condition encoder is the native drop-in ...
    for k, v in ref.fs.state_dict().items():                         # ... strict-loaded from the reference module (147 keys)
        assert torch.equal(v, wrapped.fs.state_dict()[k]), k
    assert GaussianDiffusionB200.from_reference(ref, native_fs=False).fs is ref.fs     # or the reference's own module, shared
    assert isinstance(wrapped.mel_encoder, MelEncoderB200)           # the context-mel encoder is the native drop-in
    for k, v in ref.mel_encoder.state_dict().items():
        assert torch.equal(v, wrapped.mel_encoder.state_dict()[k]), k


@needs_ref
def test_diffuse_fn_matches_live_reference():
    """q_sample / diffuse_fn of the training branch (spec_denoiser.py:126-152) with injected noise, incl. the t = -1 items that keep
    the ground-truth mel and the in-place clamp of t."""
    refshim.install("egs/spec_denoiser.yaml")
    from modules.speech_editing.spec_denoiser.diffnet import DiffNet
    from modules.speech_editing.spec_denoiser.spec_denoiser import GaussianDiffusion
    from speech_editing_toolkit_b200.modules import GaussianDiffusionB200
    ref = GaussianDiffusion(list(range(80)), 80, DiffNet(80), timesteps=8, time_scale=1, loss_type="l1", spec_min=[], spec_max=[])
    ours = GaussianDiffusionB200.from_reference(ref)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 20, 80, generator=g)
    noise = torch.randn(3, 1, 80, 20, generator=g)
    t_ref, t_ours = torch.tensor([8, -1, 3]), torch.tensor([8, -1, 3])
    a, b = ref.diffuse_fn(x, t_ref, noise=noise), ours.diffuse_fn(x, t_ours, noise=noise)
    assert torch.equal(a, b) and torch.equal(t_ref, t_ours) and int(t_ours[1]) == 0
    assert torch.equal(b[1, 0], x[1].t())


def test_registries_and_plugin_surface():
    from speech_editing_toolkit_b200 import plugin, vocoder
    assert vocoder.get_vocoder_cls("HifiGAN") is vocoder.HifiGANB200 and vocoder.get_vocoder_cls("HifiGAN_B200") is vocoder.HifiGANB200
    assert vocoder.get_vocoder_cls("nope") is None
    assert set(plugin.DIFF_DECODERS) == {"wavenet", "wavenet_b200"}
    m = plugin.build_diffusion(HP)
    assert type(m.denoise_fn).__name__ == "DiffNetB200" and m.num_timesteps == 8
    for meth in ("build_model", "build_vocoder", "run_vocoder", "forward_model", "infer_once"):
        assert hasattr(plugin.SpecDenoiserInferB200, meth)
    assert hasattr(plugin.SpeechDenoiserTaskB200, "start") and callable(plugin.run_task)
    for meth in ("run_model", "_training_step", "test_step", "test"):     # tasks/speech_editing/spec_denoiser.py, tasks/tts/speech_base.py:175
        assert hasattr(plugin.SpeechDenoiserTaskB200, meth)
    from speech_editing_toolkit_b200.modules import FastSpeechB200
    m2 = plugin.build_diffusion(HP, phone_encoder=range(41))               # as SpecDenoiserInferB200(phone_encoder=...) / b200_vocab do
    assert isinstance(m2.fs, FastSpeechB200) and m2.fs.dict_size == 41 and m.fs is None
    for meth in ("build_tts_model", "run_model", "start"):                # tasks/speech_editing/campnet.py::CampNetTask
        assert hasattr(plugin.CampNetTaskB200, meth)
    task = plugin.CampNetTaskB200(ph_dict_size=33)
    task.hparams = dict(hidden_size=192, dec_ffn_kernel_size=9, audio_num_mel_bins=80)
    assert type(task.build_tts_model()).__name__ == "CampNetB200" and task.model.encoder.embed_tokens.num_embeddings == 33
    with pytest.raises(NotImplementedError):
        task.run_model({}, infer=False)


def test_vocoder_checkpoint_layout(tmp_path):
    """<vocoder_ckpt>/config.yaml + model_ckpt_steps_N.ckpt with state_dict['model_gen']."""
    from speech_editing_toolkit_b200.engine import HIFIGAN_V1
    from speech_editing_toolkit_b200 import FseError
    from speech_editing_toolkit_b200.vocoder import HifiGANB200
    sd = synth.hifigan_state_dict(1)
    assert sd["ups.0.weight_v"].shape == (512, 256, 16) and sd["ups.0.weight_g"].shape == (512, 1, 1)     # ConvTranspose1d: dim 0 = C_in
    assert sd["conv_post.weight_v"].shape == (1, 32, 7) and len(sd) == 234
    yaml.safe_dump(dict(HIFIGAN_V1, audio_num_mel_bins=80), open(tmp_path / "config.yaml", "w"))
    torch.save({"state_dict": {"model_gen": {k: torch.from_numpy(v) for k, v in sd.items()}}}, tmp_path / "model_ckpt_steps_100.ckpt")
    if not torch.cuda.is_available():
        with pytest.raises(FseError):                 # files are found and parsed; only the CUDA handle cannot be made here
            HifiGANB200(str(tmp_path))


def test_fastspeech_drop_in_surface_and_state_dict():
    """FastSpeechB200 mirrors fs.py:49-189: constructor (dict_size, hparams), the reference's 147 state_dict keys (incl. the unused
    decoder / mel_out), forward / forward_style_embed / forward_dur and a callable `.encoder`."""
    from speech_editing_toolkit_b200.modules import FastSpeechB200, GaussianDiffusionB200, DiffNetB200
    fs = FastSpeechB200(80, HP)
    sd = synth.fastspeech_state_dict(3, 80)
    own = {k: tuple(v.shape) for k, v in fs.state_dict().items()}
    assert all(own[k] == v.shape for k, v in sd.items())
    assert all(k.startswith(("decoder.", "mel_out.")) for k in set(own) - set(sd))
    missing, unexpected = fs.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not unexpected and all(k.startswith(("decoder.", "mel_out.")) for k in missing)
    assert float(fs.dur_embed.weight[0].abs().sum()) == 0.0 and float(fs.encoder.embed_tokens.weight[0].abs().sum()) == 0.0
    for meth in ("forward", "forward_style_embed", "forward_dur", "engine"):
        assert callable(getattr(fs, meth))
    assert callable(fs.encoder)
    with pytest.raises(NotImplementedError):
        FastSpeechB200(80, dict(HP, encoder_type="fft"))
    with pytest.raises(NotImplementedError):
        fs(None, None, None, None, None, None, skip_decoder=False)
    model = GaussianDiffusionB200(list(range(80)), 80, DiffNetB200(80, HP), timesteps=8, hparams=HP)    # as build_tts_model does
    assert isinstance(model.fs, FastSpeechB200) and model.fs.dict_size == 80
    libritts = FastSpeechB200(80, dict(HP, use_pitch_embed=False))                                       # egs/spec_denoiser_libritts.yaml:169
    assert not hasattr(libritts, "pitch_embed") and not any(k.startswith("pitch_") for k in libritts.state_dict())


@needs_ref
def test_fastspeech_state_dict_matches_live_reference():
    hp = refshim.install("egs/spec_denoiser.yaml")
    from modules.speech_editing.spec_denoiser.fs import FastSpeech
    from speech_editing_toolkit_b200.modules import FastSpeechB200
    ref = FastSpeech(80, hp)
    ours = FastSpeechB200(80, dict(hp))
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    ours.load_state_dict(ref.state_dict(), strict=True)


def test_campnet_drop_in_surface_and_state_dict():
    """CampNetB200 mirrors campnet.py:14-69: constructor (ph_dict_size, word_dict_size, hparams), every tensor of the synthetic
    checkpoint at the reference's key / shape, forward signature of the reference."""
    import inspect
    from speech_editing_toolkit_b200.modules import CampNetB200
    net = CampNetB200(80, 100, dict(hidden_size=192, dec_ffn_kernel_size=9, audio_num_mel_bins=80))
    own = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = synth.campnet_state_dict(3, 80)
    assert all(own[k] == v.shape for k, v in sd.items())
    assert all(k.startswith(("encoder.pre_net.", "mel_out.")) or k.endswith("_float_tensor") for k in set(own) - set(sd))
    assert len(own) == 237
    params = list(inspect.signature(net.forward).parameters)
    assert params[:8] == ["txt_tokens", "spk_embed", "spk_id", "mels", "stutter_mel_masks", "time_mel_masks", "infer", "global_step"]


@needs_ref
def test_campnet_state_dict_matches_live_reference():
    hp = refshim.install("egs/campnet.yaml")
    from modules.speech_editing.campnet.campnet import CampNet
    from speech_editing_toolkit_b200.modules import CampNetB200
    ref = CampNet(80, 100, hp)
    ours = CampNetB200(80, 100, dict(hp))
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    ours.load_state_dict(ref.state_dict(), strict=True)


def test_campnet_task_run_model_composites_like_the_reference():
    """tasks/speech_editing/campnet.py:86: output['mel_out'] = mel_out_fine * mask + mels * (1 - mask), with a stand-in model."""
    from speech_editing_toolkit_b200 import plugin
    task = plugin.CampNetTaskB200(ph_dict_size=20)
    B, T = 2, 6
    fine = torch.full((B, T, 80), 2.0)
    seen = {}

    def fake(txt, **kw):
        seen.update(kw)
        return {"mel_out_coarse": fine * 0, "mel_out_fine": fine, "attn": torch.zeros(B, T, 3)}

    task.model = fake
    mels = torch.randn(B, T, 80)
    mask = torch.zeros(B, T); mask[:, 2:4] = 1
    out = task.run_model({"txt_tokens": torch.ones(B, 3, dtype=torch.long), "mels": mels, "time_mel_masks": mask})
    assert seen["time_mel_masks"].shape == (B, T, 1) and seen["infer"] is True and seen["stutter_mel_masks"] is None
    assert torch.equal(out["mel_out"][:, 2:4], fine[:, 2:4]) and torch.equal(out["mel_out"][:, :2], mels[:, :2])


def test_task_validation_step_and_end_host_logic():
    """SpeechDenoiserTaskB200.validation_step / validation_end against the reference's contract (tasks/speech_editing/spec_denoiser.py:64-88,
    utils/commons/base_task.py:154-185) with a stand-in run_model: scalar losses + total + nsamples, the sampling run only for the first
    num_valid_plots batches, nsamples-weighted means rounded to four decimals."""
    from speech_editing_toolkit_b200 import plugin
    task = plugin.SpeechDenoiserTaskB200()
    task.hparams = {"num_valid_plots": 1}
    calls = []

    def fake_run_model(sample, infer=False, *a, **kw):
        calls.append(infer)
        out = {"mel_out": torch.full((2, 4, 80), float(sample["k"]))}
        if infer:
            return out
        return {"l1_coarse": torch.tensor(0.25 * sample["k"]), "ssim_coarse": torch.tensor(0.125)}, out

    task.run_model = fake_run_model
    s1 = {"k": 1, "txt_tokens": torch.ones(2, 3, dtype=torch.long), "nsamples": 2}
    s2 = {"k": 3, "txt_tokens": torch.ones(6, 3, dtype=torch.long)}                      # nsamples falls back to the batch size
    o1, o2 = task.validation_step(s1, 0), task.validation_step(s2, 1)
    assert calls == [False, True, False]                                                 # batch 0 also sampled, batch 1 did not
    assert o1["losses"] == {"l1_coarse": 0.25, "ssim_coarse": 0.125} and o1["total_loss"] == 0.375 and o1["nsamples"] == 2
    assert "mel_out" in o1 and "wav_out" not in o1 and "mel_out" not in o2 and o2["nsamples"] == 6
    end = task.validation_end([o1, {}, o2])
    want_l1 = round((0.25 * 2 + 0.75 * 6) / 8, 4)
    assert end["tb_log"] == {"val/total_loss": round(want_l1 + 0.125, 4), "val/l1_coarse": want_l1, "val/ssim_coarse": 0.125}
    assert end["val_loss"] == end["tb_log"]["val/total_loss"]
    # the (total_loss, losses) tuple form of base_task.py:168-172 counts as one sample
    end2 = task.validation_end([(torch.tensor(1.0), {"a": torch.tensor(1.0)}), (3.0, {"a": 3.0})])
    assert end2["val_loss"] == 2.0 and end2["tb_log"]["val/a"] == 2.0
    import pytest
    with pytest.raises(AssertionError):
        task.validation_end([{"total_loss": 1.0}])


def test_checkpoint_save_resume_roundtrip_in_the_reference_layout(tmp_path):
    """save_ckpt / resume (the trainer's dump_checkpoint / restore_weights / restore_opt_state, utils/commons/trainer.py:372-470): file
    name, keys, atomic write, pruning to num_ckpt_keep, optimizer state; with the reference present its own load_ckpt reads the file."""
    from speech_editing_toolkit_b200 import ckpt
    from speech_editing_toolkit_b200.modules import DiffNetB200
    hp = dict(HP, residual_layers=2)
    net = DiffNetB200(80, hp)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    for p in net.parameters():
        p.grad = torch.ones_like(p) * 0.01
    opt.step()
    for step in (10, 20, 30, 40):
        path = ckpt.save_ckpt(net, str(tmp_path), step, optimizer=opt, epoch=1, best_val=0.5, num_ckpt_keep=2)
    assert path.endswith("model_ckpt_steps_40.ckpt") and not os.path.exists(path + ".part")
    assert [os.path.basename(p) for p in ckpt.get_all_ckpts(str(tmp_path))] == ["model_ckpt_steps_40.ckpt", "model_ckpt_steps_30.ckpt"]
    blob = torch.load(path, map_location="cpu", weights_only=False)
    assert set(blob) == {"epoch", "global_step", "checkpoint_callback_best", "optimizer_states", "state_dict"} and set(blob["state_dict"]) == {"model"}
    fresh = DiffNetB200(80, hp)
    opt2 = torch.optim.AdamW(fresh.parameters(), lr=1e-3)
    assert ckpt.resume(fresh, str(tmp_path), optimizer=opt2) == (40, 1)
    for (k, a), (_, b) in zip(net.state_dict().items(), fresh.state_dict().items()):
        assert torch.equal(a, b), k
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert all(torch.equal(s1[i]["exp_avg"], s2[i]["exp_avg"]) for i in s1)
    assert ckpt.resume(DiffNetB200(80, hp), str(tmp_path / "empty")) == (0, 0)
    if refshim.available():
        refshim.install("egs/spec_denoiser.yaml", overrides="residual_layers=2")
        from utils.commons.hparams import hparams as ref_hparams
        ref_hparams["residual_layers"] = 2
        from modules.speech_editing.spec_denoiser.diffnet import DiffNet
        from utils.commons.ckpt_utils import load_ckpt as ref_load_ckpt
        ref_net = DiffNet(80)
        ref_load_ckpt(ref_net, str(tmp_path), "model", strict=True)
        for k, v in net.state_dict().items():
            assert torch.equal(v, ref_net.state_dict()[k]), k
