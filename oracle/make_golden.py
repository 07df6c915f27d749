"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on
seeded synthetic weights / inputs.  Runs only in the build container (the reference cannot travel);
the fixtures it writes are committed.  Usage:  python oracle/make_golden.py

The reference has no tests or golden vectors of its own, so these fixtures — outputs of
the reference's own code — are what pins both the numpy oracle and the CUDA path.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from speech_editing_toolkit_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED = 1234


def to_torch(sd):
    return {k: torch.from_numpy(v.copy()) for k, v in sd.items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(SEED)
    torch.set_num_threads(8)
    hp = refshim.install("egs/spec_denoiser.yaml", overrides="timesteps=10")
    from modules.speech_editing.spec_denoiser.diffnet import DiffNet
    from modules.speech_editing.spec_denoiser import spec_denoiser as sdmod
    from modules.vocoder.hifigan.hifigan import HifiGanGenerator

    # ---- DiffNet.forward, one step, ragged T and two different t ------------------------------
    sd = synth.denoiser_state_dict(SEED)
    net = DiffNet(hp["audio_num_mel_bins"]).eval()
    net.load_state_dict(to_torch(sd), strict=True)
    B, T = 2, 80
    rs = np.random.RandomState(SEED + 1)
    x = rs.standard_normal((B, 80, T)).astype(np.float32)
    cond = synth.synthetic_cond(SEED, B, T)
    t = np.array([7, 0], dtype=np.int64)
    with torch.no_grad():
        # the reference passes cond as the transposed VIEW of a [B,T,H] tensor (spec_denoiser.py:167)
        x0 = net(torch.from_numpy(x)[:, None], torch.from_numpy(t), torch.from_numpy(cond).transpose(1, 2))[:, 0].numpy()
    np.savez_compressed(os.path.join(OUT, "diffnet_step.npz"), seed=SEED, B=B, T=T, x=x, t=t, x0=x0)
    print("diffnet_step", x0.shape, float(np.abs(x0).mean()))

    # ---- config C1: B=1, T=256, S=10 sampling through the reference's own p_sample -----------------
    S, B, T = 10, 1, 256
    model = sdmod.GaussianDiffusion(phone_encoder=list(range(80)), out_dims=80, denoise_fn=net, timesteps=S,
                                    time_scale=hp["timescale"], loss_type=hp["diff_loss_type"],
                                    spec_min=hp["spec_min"], spec_max=hp["spec_max"]).eval()
    cond = synth.synthetic_cond(SEED + 1, B, T)
    noise = synth.synthetic_noise(SEED + 1, S, B, T)
    draws = iter(torch.from_numpy(noise[1:]))
    orig = sdmod.noise_like
    sdmod.noise_like = lambda shape, device, repeat=False: next(draws)[:, None]     # inject the pre-drawn normals
    try:
        xt = torch.from_numpy(noise[0])[:, None]
        cond_t = torch.from_numpy(cond).transpose(1, 2)
        trace = []
        for i in reversed(range(S)):                                                # spec_denoiser.py:181-182
            xt = model.p_sample(xt, torch.full((B,), i, dtype=torch.long), cond_t)
            trace.append(xt[:, 0].numpy().copy())
        mel = xt[:, 0].transpose(1, 2).numpy().copy()                               # :183
    finally:
        sdmod.noise_like = orig
    np.savez_compressed(os.path.join(OUT, "sample_c1.npz"), seed=SEED + 1, B=B, T=T, S=S, mel_out=mel,
                        x_after_first=trace[0], x_after_fifth=trace[4])
    print("sample_c1", mel.shape, float(np.abs(mel).mean()))
    sched = {k: getattr(model, k).numpy() for k in ("betas", "alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
                                                    "posterior_log_variance_clipped", "posterior_variance")}
    kat = {f"sched10_{k}": v for k, v in sched.items()}
    for S2 in (4, 8, 100):
        m2 = sdmod.GaussianDiffusion(phone_encoder=list(range(80)), out_dims=80, denoise_fn=net, timesteps=S2, time_scale=1,
                                     loss_type="l1", spec_min=[], spec_max=[])
        for k in ("betas", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped"):
            kat[f"sched{S2}_{k}"] = getattr(m2, k).numpy()

    # ---- integer / index known answers, from the reference functions ----------
    from utils.audio.pitch.utils import f0_to_coarse
    from utils.audio.align import mel2token_to_dur
    from modules.tts.commons.align_ops import expand_states
    f0 = np.array([0, 50, 80, 100, 220, 440, 600, 900, 1200], dtype=np.float32)
    kat["f0_in"] = f0
    kat["f0_coarse"] = f0_to_coarse(torch.from_numpy(f0)).numpy()
    m2t = np.array([[1, 1, 2, 2, 2, 4, 0, 0]], dtype=np.int64)
    kat["mel2token"] = m2t
    kat["mel2token_dur"] = mel2token_to_dur(torch.from_numpy(m2t), 5).numpy()
    hs = np.arange(6, dtype=np.float32).reshape(1, 3, 2)
    m2p = np.array([[1, 1, 3, 0]], dtype=np.int64)
    kat["expand_h"], kat["expand_idx"] = hs, m2p
    kat["expand_out"] = expand_states(torch.from_numpy(hs), torch.from_numpy(m2p)).numpy()
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **kat)
    print("kat", sorted(kat)[:4], "...")

    # ---- HiFi-GAN V1 generator ------------------------------------------------------------------
    from oracle.fluentspeech_oracle import HIFIGAN_V1
    gen = HifiGanGenerator(dict(HIFIGAN_V1)).eval()
    hsd = synth.hifigan_state_dict(SEED)
    gen.load_state_dict(to_torch(hsd), strict=True)
    B, T = 1, 24
    mel_in = np.clip(np.random.RandomState(SEED + 2).standard_normal((B, T, 80)) * 1.5 - 3.0, -6, 1.5).astype(np.float32)
    with torch.no_grad():
        wav = gen(torch.from_numpy(mel_in).transpose(1, 2)).numpy()          # [B,1,T*256]
    np.savez_compressed(os.path.join(OUT, "hifigan_v1.npz"), seed=SEED, B=B, T=T, mel=mel_in, wav=wav[:, 0])
    print("hifigan", wav.shape, float(np.abs(wav).mean()), float(np.abs(wav).max()))


def mel_encoder_fixture():
    """MelEncoder.forward (mel_encoder.py:15-19) and its call site's arithmetic (spec_denoiser.py:162-164) from the
    unmodified reference module: `python oracle/make_golden.py mel_encoder` writes only this fixture."""
    os.makedirs(OUT, exist_ok=True)
    refshim.install("egs/spec_denoiser.yaml", overrides="timesteps=10")
    from modules.speech_editing.commons.mel_encoder import MelEncoder
    sd = synth.mel_encoder_state_dict(SEED)
    enc = MelEncoder(80, 192).eval()
    enc.load_state_dict(to_torch(sd), strict=True)
    B, T = 2, 72
    ref, mask = synth.synthetic_ref_and_mask(SEED, B, T)
    rs = np.random.RandomState(SEED + 3)
    decoder_inp = rs.standard_normal((B, T, 192)).astype(np.float32)
    nonpad = np.ones((B, T, 1), dtype=np.float32)
    nonpad[1, 60:] = 0.0                                       # trailing padding frames of the second item (mel2ph == 0)
    with torch.no_grad():
        x = torch.from_numpy(ref) * (1 - torch.from_numpy(mask))
        out = enc(x)
        cond = torch.from_numpy(decoder_inp.copy())
        cond += out * torch.from_numpy(nonpad)                  # spec_denoiser.py:164
    np.savez_compressed(os.path.join(OUT, "mel_encoder.npz"), seed=SEED, B=B, T=T, decoder_inp=decoder_inp, nonpad=nonpad,
                        out=out.numpy(), cond=cond.numpy())
    print("mel_encoder", out.shape, float(np.abs(out.numpy()).mean()))


def cond_encoder_fixture():
    """FastSpeech.forward(skip_decoder=True) (fs.py:83-105), its parts as the inference script calls them
    (inference/tts/spec_denoiser.py:84-98: encoder, forward_style_embed, forward_dur with masked_dur + LengthRegulator) and
    the whole GaussianDiffusion.forward(infer=True) (spec_denoiser.py:154-185) from the unmodified reference on a ragged
    batch: `python oracle/make_golden.py cond_encoder` writes tests/golden/cond_encoder.npz and fluentspeech_e2e.npz."""
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(SEED)
    hp = refshim.install("egs/spec_denoiser.yaml", overrides="timesteps=4")
    from modules.speech_editing.spec_denoiser.fs import FastSpeech
    from modules.speech_editing.spec_denoiser.diffnet import DiffNet
    from modules.speech_editing.spec_denoiser import spec_denoiser as sdmod
    from utils.audio.pitch.utils import f0_to_coarse
    vocab = 80
    fs = FastSpeech(vocab, hp).eval()
    fsd = synth.fastspeech_state_dict(SEED, vocab)
    missing, unexpected = fs.load_state_dict(to_torch(fsd), strict=False)
    assert not unexpected and all(k.startswith(("decoder.", "mel_out.")) for k in missing), (missing, unexpected)
    B, T = 2, 80
    batch = synth.pad_edit_batch(synth.synthetic_edit_batch(SEED, B, T, vocab=vocab), item=1, n_tokens=3)
    tb = {k: torch.from_numpy(v) for k, v in batch.items()}
    mask3 = tb["time_mel_masks"][:, :, None]
    out = dict(seed=SEED, B=B, T=T, vocab=vocab)
    with torch.no_grad():
        out["encoder_out"] = fs.encoder(tb["txt_tokens"]).numpy()
        out["style_embed"] = fs.forward_style_embed(tb["spk_embed"], None).numpy()
        for flag in (False, True):
            ret = fs(tb["txt_tokens"], mask3, tb["mel2ph"], tb["spk_embed"], tb["f0"], tb["uv"], skip_decoder=True, infer=True,
                     use_pred_pitch=flag)
            sfx = "_predpitch" if flag else ""
            for k in ("decoder_inp", "dur", "pitch_pred", "f0_denorm", "f0_denorm_pred", "mel2ph"):
                out[k + sfx] = ret[k].numpy()
            out["pitch" + sfx] = f0_to_coarse(ret["f0_denorm"]).numpy()
        # the inference script's use of forward_dur: explicit masked_dur, predicted mel2ph (LengthRegulator)
        src_nonpadding = (tb["txt_tokens"] > 0).float()[:, :, None]
        dur_inp = (torch.from_numpy(out["encoder_out"]) + torch.from_numpy(out["style_embed"])) * src_nonpadding
        masked_dur = torch.from_numpy(np.random.RandomState(SEED + 7).randint(0, 12, size=batch["txt_tokens"].shape)) * (tb["txt_tokens"] > 0)
        r2 = {}
        mel2ph_pred = fs.forward_dur(dur_inp, tb["time_mel_masks"], tb["mel2ph"].clone(), tb["txt_tokens"], r2, masked_dur=masked_dur,
                                     use_pred_mel2ph=True)
        out["masked_dur_in"] = masked_dur.numpy()
        out["dur_masked_dur"] = r2["dur"].numpy()
        out["mel2ph_pred"] = mel2ph_pred.numpy()
    np.savez_compressed(os.path.join(OUT, "cond_encoder.npz"), **out)
    print("cond_encoder", out["decoder_inp"].shape, float(np.abs(out["decoder_inp"]).mean()), "dur", out["dur"][0, :6],
          "mel2ph_pred", out["mel2ph_pred"].shape, "pitch bins", np.unique(out["pitch_predpitch"]).size)

    # ---- whole model, S = 4, noise injected (x_S via torch.randn, per-step draws via noise_like) ----
    S, L = 4, 4
    from utils.commons.hparams import hparams as ref_hparams      # DiffNet reads the module-global dict (diffnet.py:9,89-92)
    ref_hparams["residual_layers"] = hp["residual_layers"] = L
    net = DiffNet(hp["audio_num_mel_bins"]).eval()
    net.load_state_dict(to_torch(synth.denoiser_state_dict(SEED, layers=L)), strict=True)
    model = sdmod.GaussianDiffusion(phone_encoder=list(range(vocab)), out_dims=80, denoise_fn=net, timesteps=S,
                                    time_scale=hp["timescale"], loss_type=hp["diff_loss_type"], spec_min=hp["spec_min"],
                                    spec_max=hp["spec_max"]).eval()
    model.fs.load_state_dict(to_torch(fsd), strict=False)
    model.mel_encoder.load_state_dict(to_torch(synth.mel_encoder_state_dict(SEED)), strict=True)
    noise = synth.synthetic_noise(SEED + 5, S, B, T)
    draws = iter(torch.from_numpy(noise[1:]))
    orig_nl, orig_randn = sdmod.noise_like, torch.randn
    sdmod.noise_like = lambda shape, device, repeat=False: next(draws)[:, None]
    torch.randn = lambda *a, **k: torch.from_numpy(noise[0])[:, None]
    try:
        with torch.no_grad():
            ret = model(tb["txt_tokens"], mask3, tb["mel2ph"], tb["spk_embed"], tb["ref_mels"], tb["f0"], tb["uv"], infer=True,
                        use_pred_pitch=True)
    finally:
        sdmod.noise_like, torch.randn = orig_nl, orig_randn
    np.savez_compressed(os.path.join(OUT, "fluentspeech_e2e.npz"), seed=SEED, B=B, T=T, S=S, layers=L, vocab=vocab,
                        mel_out=ret["mel_out"].numpy(), decoder_inp=ret["decoder_inp"].numpy(), dur=ret["dur"].numpy())
    print("fluentspeech_e2e", ret["mel_out"].shape, float(np.abs(ret["mel_out"].numpy()).mean()))


def campnet_fixture():
    """CampNet.forward (modules/speech_editing/campnet/campnet.py:40-69) from the unmodified reference on a ragged batch:
    `python oracle/make_golden.py campnet` writes tests/golden/campnet.npz."""
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(SEED)
    hp = refshim.install("egs/campnet.yaml")
    from modules.speech_editing.campnet.campnet import CampNet
    vocab = 80
    net = CampNet(vocab, 100, hp).eval()
    sd = synth.campnet_state_dict(SEED, vocab)
    missing, unexpected = net.load_state_dict(to_torch(sd), strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("encoder.pre_net.", "mel_out.")) or k.endswith("_float_tensor") for k in missing), missing
    B, T = 2, 160
    batch = synth.synthetic_campnet_batch(SEED, B, T, vocab=vocab, pad_items=[(1, 4)])
    with torch.no_grad():
        ret = net(torch.from_numpy(batch["txt_tokens"]), mels=torch.from_numpy(batch["mels"]),
                  time_mel_masks=torch.from_numpy(batch["time_mel_masks"]), infer=True)
        enc, _ = net.run_text_encoder(torch.from_numpy(batch["txt_tokens"]), {})
    np.savez_compressed(os.path.join(OUT, "campnet.npz"), seed=SEED, B=B, T=T, vocab=vocab, mel_out_coarse=ret["mel_out_coarse"].numpy(),
                        mel_out_fine=ret["mel_out_fine"].numpy(), attn=ret["attn"].numpy().astype(np.float16), encoder_out=enc.numpy())
    print("campnet", ret["mel_out_fine"].shape, float(np.abs(ret["mel_out_fine"].numpy()).mean()), "attn", ret["attn"].shape)


def edit_region_fixture():
    """The region surgery of the inference script, from the reference's OWN code: inference/tts/spec_denoiser.py cannot be
    imported (inference_acl, resemblyzer, g2p_en, ... are absent), so `SpecDenoiserInfer.forward_model` (:63-149) is cut out of
    the source file with ast and run unmodified against a stand-in `self` whose model.fs is the real reference FastSpeech; the
    tensors it hands to forward_dur and to the model are recorded.  `python oracle/make_golden.py edit_region` writes
    tests/golden/edit_region.npz (three utterances: edit in the middle, at the end (no tail), longer replacement)."""
    import ast
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(SEED)
    hp = refshim.install("egs/spec_denoiser.yaml")
    from modules.speech_editing.spec_denoiser.fs import FastSpeech
    src_path = os.path.join(refshim.REF_ROOT, "inference", "tts", "spec_denoiser.py")
    tree = ast.parse(open(src_path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SpecDenoiserInfer")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "forward_model")
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), src_path, "exec"), ns)          # the reference's code, unmodified
    forward_model = ns["forward_model"]
    vocab = 80
    fs = FastSpeech(vocab, hp).eval()
    fs.load_state_dict(to_torch(synth.fastspeech_state_dict(SEED, vocab)), strict=False)

    class Model:
        def __init__(self):
            self.fs, self.calls, self.dur_calls = fs, [], []
            real = fs.forward_dur

            def spy(dur_inp, masks, mel2ph, txt, ret, masked_dur=None, use_pred_mel2ph=False):
                out = real(dur_inp, masks, mel2ph, txt, ret, masked_dur=masked_dur, use_pred_mel2ph=use_pred_mel2ph)
                self.dur_calls.append(dict(masked_dur=masked_dur.clone(), masked_mel2ph=mel2ph.clone(), time_mel_masks_orig=masks.clone(),
                                           edited_mel2ph=out.clone()))
                return out
            fs.forward_dur = spy

        def __call__(self, txt, **kw):
            self.calls.append(dict(kw, txt=txt))
            return {"mel_out": torch.zeros(1, kw["mel2ph"].shape[1], 80)}

    class Self:
        device = "cpu"

        def __init__(self, sample):
            self.model, self._sample = Model(), sample

        def input_to_batch(self, inp):
            return self._sample

        def run_vocoder(self, c):
            return torch.zeros(1, 8)

    out = {"seed": SEED, "vocab": vocab}
    cases = [dict(seed=SEED, n_words=9, edit_span=(3, 4), new_span_phones=(2, 3, 1)),
             dict(seed=SEED + 1, n_words=7, edit_span=(6, 7), new_span_phones=(4,)),                 # edit at the end: no tail
             dict(seed=SEED + 2, n_words=8, edit_span=(1, 2), new_span_phones=(3, 3, 2, 2))]          # edit at the start: no head
    for i, kw in enumerate(cases):
        item = synth.synthetic_edit_item(vocab=vocab, **kw)
        lt = lambda k: torch.from_numpy(item[k])[None]
        sample = {"edited_txt_tokens": lt("edited_ph_token"), "mel": lt("mel"), "mel2ph": lt("mel2ph").clone(), "mel2word": lt("mel2word"),
                  "dur": lt("dur"), "ph2word": lt("ph2word"), "edited_ph2word": lt("edited_ph2word"), "f0": lt("f0"), "uv": lt("uv"),
                  "words_region": item["words_region"], "edited_words_region": item["edited_words_region"], "text": ["synthetic"],
                  "spk_embed": lt("spk_embed")}
        me = Self(sample)
        with torch.no_grad():
            forward_model(me, None)
        call, dcall = me.model.calls[0], me.model.dur_calls[0]
        assert call["use_pred_pitch"] is True and call["infer"] is True
        for k, v in (("masked_dur", dcall["masked_dur"]), ("masked_mel2ph", dcall["masked_mel2ph"]), ("time_mel_masks_orig", dcall["time_mel_masks_orig"]),
                     ("edited_mel2ph_pred", dcall["edited_mel2ph"]), ("mel2ph", call["mel2ph"]), ("ref_mels", call["ref_mels"]), ("f0", call["f0"]),
                     ("uv", call["uv"]), ("time_mel_masks", call["time_mel_masks"])):
            out[f"c{i}_{k}"] = v[0].numpy()
        out[f"c{i}_kw"] = np.array([kw["seed"], kw["n_words"], *kw["edit_span"], len(kw["new_span_phones"]), *kw["new_span_phones"]], dtype=np.int64)
        print("edit_region case", i, "T", item["mel2ph"].shape[0], "->", call["mel2ph"].shape[1], "masked frames", int(call["time_mel_masks"].sum()))
    np.savez_compressed(os.path.join(OUT, "edit_region.npz"), **out)


def bench_config_fixture():
    """Parity pins AT THE BENCHMARKED SHAPES (BASELINE configs[1]/[2]/[3]: T = 1024 frames, S = 100 iterations), from the
    unmodified reference: `python oracle/make_golden.py bench_config` writes
      sample_s100_t1024.npz   the reference's own p_sample loop (spec_denoiser.py:177-185), B=1 x T=1024 x S=100, the 101 normal
                              draws injected from synth.synthetic_noise(seed) (regenerated by the test; not stored)
      hifigan_t1024.npz       HifiGanGenerator.forward (hifigan.py:126-142), B=1 x T=1024 (262 144 samples)
      campnet_t1024.npz       CampNet.forward (campnet.py:40-69), one item of T=1024 frames / 128 tokens
    Items are independent in every kernel (no cross-item op), so B=1 at the full T and S pins the B=32 runs
    together with the item-independence tests."""
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(SEED)
    torch.set_num_threads(8)
    S, B, T = 100, 1, 1024
    hp = refshim.install("egs/spec_denoiser.yaml", overrides=f"timesteps={S}")
    from modules.speech_editing.spec_denoiser.diffnet import DiffNet
    from modules.speech_editing.spec_denoiser import spec_denoiser as sdmod
    from modules.vocoder.hifigan.hifigan import HifiGanGenerator
    net = DiffNet(hp["audio_num_mel_bins"]).eval()
    net.load_state_dict(to_torch(synth.denoiser_state_dict(SEED)), strict=True)
    model = sdmod.GaussianDiffusion(phone_encoder=list(range(80)), out_dims=80, denoise_fn=net, timesteps=S, time_scale=hp["timescale"],
                                    loss_type=hp["diff_loss_type"], spec_min=hp["spec_min"], spec_max=hp["spec_max"]).eval()
    cond = synth.synthetic_cond(SEED + 11, B, T)
    noise = synth.synthetic_noise(SEED + 11, S, B, T)
    draws = iter(torch.from_numpy(noise[1:]))
    orig = sdmod.noise_like
    sdmod.noise_like = lambda shape, device, repeat=False: next(draws)[:, None]
    try:
        with torch.no_grad():
            xt = torch.from_numpy(noise[0])[:, None]
            cond_t = torch.from_numpy(cond).transpose(1, 2)
            mid = {}
            for i in reversed(range(S)):                                            # spec_denoiser.py:181-182
                xt = model.p_sample(xt, torch.full((B,), i, dtype=torch.long), cond_t)
                if i in (75, 50, 25):
                    mid[i] = xt[:, 0, :, :64].numpy().copy()                         # 64 frames of the running x_t (trace spot checks)
            mel = xt[:, 0].transpose(1, 2).numpy().copy()
    finally:
        sdmod.noise_like = orig
    np.savez_compressed(os.path.join(OUT, "sample_s100_t1024.npz"), seed=SEED + 11, B=B, T=T, S=S, mel_out=mel,
                        x_t75=mid[75], x_t50=mid[50], x_t25=mid[25])
    print("sample_s100_t1024", mel.shape, float(np.abs(mel).mean()))

    from oracle.fluentspeech_oracle import HIFIGAN_V1
    gen = HifiGanGenerator(dict(HIFIGAN_V1)).eval()
    gen.load_state_dict(to_torch(synth.hifigan_state_dict(SEED)), strict=True)
    mel_in = np.clip(np.random.RandomState(SEED + 12).standard_normal((1, T, 80)) * 1.5 - 3.0, -6, 1.5).astype(np.float32)
    with torch.no_grad():
        wav = gen(torch.from_numpy(mel_in).transpose(1, 2)).numpy()
    np.savez_compressed(os.path.join(OUT, "hifigan_t1024.npz"), seed=SEED + 12, B=1, T=T, wav=wav[:, 0])      # mel regenerated from the seed
    print("hifigan_t1024", wav.shape, float(np.abs(wav).mean()))

    hp = refshim.install("egs/campnet.yaml")
    from modules.speech_editing.campnet.campnet import CampNet
    vocab = 80
    camp = CampNet(vocab, 100, hp).eval()
    camp.load_state_dict(to_torch(synth.campnet_state_dict(SEED, vocab)), strict=False)
    batch = synth.synthetic_campnet_batch(SEED + 13, 1, T, vocab=vocab)
    with torch.no_grad():
        ret = camp(torch.from_numpy(batch["txt_tokens"]), mels=torch.from_numpy(batch["mels"]),
                   time_mel_masks=torch.from_numpy(batch["time_mel_masks"]), infer=True)
    np.savez_compressed(os.path.join(OUT, "campnet_t1024.npz"), seed=SEED + 13, B=1, T=T, vocab=vocab,
                        mel_out_coarse=ret["mel_out_coarse"].numpy().astype(np.float32), mel_out_fine=ret["mel_out_fine"].numpy().astype(np.float32))
    print("campnet_t1024", ret["mel_out_fine"].shape, float(np.abs(ret["mel_out_fine"].numpy()).mean()))


def mel_frontend_fixture():
    """tests/golden/mel_frontend.npz WITHOUT the oracle's own transform: `scipy.signal.stft` (an independent implementation: Hann
    window from scipy.signal.get_window(fftbins=True), zero boundary = librosa's center=True / pad_mode="constant") gives |X|, and the
    Slaney filterbank is built here from librosa's documented definition (Slaney scale 200/3 Hz per mel below 1 kHz, log step
    ln(6.4)/27 above; triangles between n_mels + 2 band edges; area normalisation) after checking the scale against the constants
    published in librosa's documentation (mel_frequencies(n_mels=40) example, hz_to_mel(60) = 0.9, mel_to_hz(3) = 200).  librosa
    itself is absent from the container and from /root/reference, so this is the strongest pin available:
    `python oracle/make_golden.py mel_frontend`."""
    import scipy.signal as ss
    os.makedirs(OUT, exist_ok=True)
    sr, n_fft, hop, n_mels, fmin, fmax, eps = 22050, 1024, 256, 80, 55.0, 7600.0, 1e-6
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    h2m = lambda f: np.where(np.asarray(f, float) >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-12) / min_log_hz) / logstep, np.asarray(f, float) / f_sp)
    m2h = lambda m: np.where(np.asarray(m, float) >= min_log_mel, min_log_hz * np.exp(logstep * (np.asarray(m, float) - min_log_mel)), f_sp * np.asarray(m, float))
    doc = np.array([0., 85.317, 170.635, 255.952, 341.269, 426.586, 511.904, 597.221, 682.538, 767.855, 853.173, 938.49, 1024.856, 1119.114,
                    1222.042, 1334.436, 1457.167, 1591.187, 1737.532, 1897.337, 2071.84, 2262.393, 2470.47, 2697.686, 2945.799, 3216.731,
                    3512.582, 3835.643, 4188.417, 4573.636, 4994.285, 5453.621, 5955.205, 6502.92, 7101.009, 7754.107, 8467.272, 9246.028,
                    10096.408, 11025.])            # librosa documentation: librosa.mel_frequencies(n_mels=40)
    assert np.abs(m2h(np.linspace(h2m(0.0), h2m(11025.0), 40)) - doc).max() < 1e-3
    assert abs(float(h2m(60.0)) - 0.9) < 1e-12 and abs(float(m2h(3.0)) - 200.0) < 1e-9
    freqs = np.linspace(0, sr / 2.0, n_fft // 2 + 1)
    edges = m2h(np.linspace(h2m(fmin), h2m(fmax), n_mels + 2))
    fb = np.zeros((n_mels, len(freqs)))
    for i in range(n_mels):
        up = (freqs - edges[i]) / (edges[i + 1] - edges[i])
        down = (edges[i + 2] - freqs) / (edges[i + 2] - edges[i + 1])
        fb[i] = np.maximum(0.0, np.minimum(up, down)) * 2.0 / (edges[i + 2] - edges[i])
    fb = fb.astype(np.float32)
    rs = np.random.RandomState(SEED + 21)
    n = hop * 41 + 77
    t = np.arange(n) / sr
    wav = (0.1 * rs.standard_normal(n) + 0.3 * np.sin(2 * np.pi * 440.0 * t) * (t > 0.2)).astype(np.float32)
    win = ss.get_window("hann", n_fft, fftbins=True)
    _, _, Z = ss.stft(wav.astype(np.float64), fs=sr, window=win, nperseg=n_fft, noverlap=n_fft - hop, nfft=n_fft, boundary="zeros", padded=False,
                      return_onesided=True, scaling="spectrum")
    T = 1 + n // hop
    mag = (np.abs(Z).T * win.sum())[:T].astype(np.float32)          # undo scipy's 1 / sum(window) scaling; librosa emits 1 + n // hop frames
    assert mag.shape == (T, n_fft // 2 + 1)
    mel = np.log10(np.maximum(eps, mag.astype(np.float64) @ fb.astype(np.float64).T)).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "mel_frontend.npz"), seed=SEED + 21, wav=wav, mel=mel, mel_basis_row0=fb[0], mel_basis_row40=fb[40],
                        mel_basis_row79=fb[79], band_edges=edges)
    print("mel_frontend", mel.shape, float(mel.mean()))


def train_fixture():
    """Gradients of the UNMODIFIED reference DiffNet under torch.autograd (diffnet.py:60-132; the denoiser call of
    GaussianDiffusion.forward(infer=False), spec_denoiser.py:168-176): `python oracle/make_golden.py train` writes
    tests/golden/diffnet_train.npz: x0 for seeded (x_t, t, cond), and for a seeded upstream gradient dx0 the gradient of cond and of
    every parameter (small ones whole, large ones as every 61st element + their L2 norm)."""
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(SEED)
    torch.set_num_threads(8)
    L, B, T = 4, 2, 80
    hp = refshim.install("egs/spec_denoiser.yaml", overrides=f"timesteps=100,residual_layers={L}")
    from utils.commons.hparams import hparams as ref_hparams
    ref_hparams["residual_layers"] = L
    from modules.speech_editing.spec_denoiser.diffnet import DiffNet
    net = DiffNet(hp["audio_num_mel_bins"]).train()
    net.load_state_dict(to_torch(synth.denoiser_state_dict(SEED + 31, layers=L)), strict=True)
    rs = np.random.RandomState(SEED + 32)
    x = rs.standard_normal((B, 80, T)).astype(np.float32)
    cond = synth.synthetic_cond(SEED + 33, B, T)                       # [B, T, H]
    t = np.array([7, 93], dtype=np.int64)
    dx0 = (rs.standard_normal((B, 80, T)) * 0.1).astype(np.float32)
    cond_t = torch.from_numpy(cond).requires_grad_(True)
    x0 = net(torch.from_numpy(x)[:, None], torch.from_numpy(t), cond_t.transpose(1, 2))[:, 0]
    x0.backward(torch.from_numpy(dx0))
    out = dict(seed=SEED + 31, B=B, T=T, layers=L, x=x, t=t, dx0=dx0, x0=x0.detach().numpy(), dcond=cond_t.grad.numpy())
    for name, p in net.named_parameters():
        g = p.grad.numpy().reshape(-1)
        out["gnorm__" + name] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        out["g__" + name] = g if g.size <= 4096 else g[::61].copy()
    np.savez_compressed(os.path.join(OUT, "diffnet_train.npz"), **out)
    print("diffnet_train", x0.shape, float(np.abs(out["dcond"]).mean()), len([k for k in out if k.startswith("g__")]), "parameter gradients")


def mel_loss_fixture():
    """Mel losses of the training step and their gradient from the reference's OWN methods: SpeechBaseTask.l1_loss / ssim_loss are cut
    out of tasks/tts/speech_base.py with ast (the task module cannot be imported: matplotlib, librosa, ...) and run over the
    reference's ssim() and weights_nonzero_speech under torch.autograd.  `python oracle/make_golden.py mel_loss` writes
    tests/golden/mel_loss.npz: a ragged, partly masked batch (padding frames and unmasked frames are zero rows of the target, weight 0),
    losses (fp32 as the reference computes them, and an fp64 run of the same code as the arbiter of rounding) and d(0.5 l1 + 0.5 ssim)/d mel_out."""
    import ast
    os.makedirs(OUT, exist_ok=True)
    refshim.install("egs/spec_denoiser.yaml")
    import utils.metrics.ssim as ref_ssim
    from utils.nn.seq_utils import weights_nonzero_speech
    import torch.nn.functional as F
    src = open(os.path.join(refshim.REF_ROOT, "tasks", "tts", "speech_base.py")).read()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "SpeechBaseTask")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("l1_loss", "ssim_loss")]
    ns = {"F": F, "ssim": ref_ssim.ssim, "weights_nonzero_speech": weights_nonzero_speech, "torch": torch}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "speech_base.py", "exec"), ns)
    rs = np.random.RandomState(SEED + 41)
    B, T, M = 3, 70, 80
    target = np.clip(rs.standard_normal((B, T, M)) * 1.5 - 3.0, -6.0, 1.5).astype(np.float32)
    out = (target + rs.standard_normal((B, T, M)) * 0.4).astype(np.float32)
    mask = np.zeros((B, T, 1), np.float32)
    mask[0, 10:45] = 1
    mask[1, 0:20] = 1                    # region touching the first frame (zero padding of the window)
    mask[2, 50:70] = 1                   # ... and the last
    mask[2, 58:60] = 0                   # a hole inside a region
    out[0, 20, 7] = target[0, 20, 7]     # an exact tie: sign(0) = 0 in the l1 gradient
    res = {}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        ref_ssim.window = None           # the module caches its window (and its dtype) in a global
        a = torch.from_numpy(out * mask).to(dt).requires_grad_(True)
        b = torch.from_numpy(target * mask).to(dt)
        l1 = ns["l1_loss"](None, a, b) * 0.5
        ss = ns["ssim_loss"](None, a, b) * 0.5
        (l1 + ss).backward()
        res["l1_" + tag], res["ssim_" + tag], res["grad_" + tag] = l1.detach().numpy(), ss.detach().numpy(), a.grad.numpy()
    ref_ssim.window = None
    np.savez_compressed(os.path.join(OUT, "mel_loss.npz"), mel_out=out * mask, target=target * mask, lambda_l1=0.5, lambda_ssim=0.5, **res)
    print("mel_loss", float(res["l1_f32"]), float(res["ssim_f32"]), float(res["l1_f64"]), float(res["ssim_f64"]),
          float(np.abs(res["grad_f32"] - res["grad_f64"]).max()), float(np.abs(res["grad_f64"]).max()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mel_encoder":
        mel_encoder_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "cond_encoder":
        cond_encoder_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "campnet":
        campnet_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "edit_region":
        edit_region_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "bench_config":
        bench_config_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "mel_frontend":
        mel_frontend_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "train":
        train_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "mel_loss":
        mel_loss_fixture()
    else:
        main()
        mel_encoder_fixture()
        cond_encoder_fixture()
        campnet_fixture()
        edit_region_fixture()
        bench_config_fixture()
        mel_frontend_fixture()
        train_fixture()
