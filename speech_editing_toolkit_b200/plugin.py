"""The reference's task / inference seams for the spec_denoiser path, backed by the C ABI:

  DIFF_DECODERS                       tasks/speech_editing/spec_denoiser.py:13-15 (+ inference/tts/spec_denoiser.py:26-28)
  SpeechDenoiserTaskB200.start()      task_cls contract of tasks/run.py:9-14 (inference / test side of
                                      tasks/speech_editing/spec_denoiser.py::SpeechDenoiserTask)
  SpecDenoiserInferB200               class + method surface of inference/tts/spec_denoiser.py::SpecDenoiserInfer
                                      (build_model / build_vocoder / run_vocoder / forward_model / infer_once)
  install_into_reference()            registers the B200 denoiser / vocoder in the reference's own registries
  CampNetTaskB200                     inference side of tasks/speech_editing/campnet.py::CampNetTask (build_tts_model / run_model)

Text / MFA / pitch front-ends (g2p_en, MFA binary, parselmouth, resemblyzer) are outside the hot path; these
classes take tensors at the `GaussianDiffusion.forward` boundary, like the benchmark does.
"""
from __future__ import annotations

import importlib
import os
from typing import Optional

import numpy as np
import torch

from . import synth
from .ckpt import load_ckpt
from .hparams import hparams, set_hparams
from .modules import CampNetB200, DiffNetB200, GaussianDiffusionB200
from .vocoder import HifiGANB200, get_vocoder_cls, register_vocoder  # noqa: F401

DIFF_DECODERS = {
    "wavenet": lambda hp: DiffNetB200(hp["audio_num_mel_bins"], hp),
    "wavenet_b200": lambda hp: DiffNetB200(hp["audio_num_mel_bins"], hp),
}


def install_into_reference() -> dict:
    """Inside the reference tree: add 'wavenet_b200' to both DIFF_DECODERS dicts and 'HifiGAN_B200' to the
    vocoder registry, so that `-hp diff_decoder_type=wavenet_b200,vocoder=HifiGAN_B200` selects this path."""
    done = {}
    for mod_name in ("tasks.speech_editing.spec_denoiser", "inference.tts.spec_denoiser"):
        try:
            mod = importlib.import_module(mod_name)
            mod.DIFF_DECODERS["wavenet_b200"] = DIFF_DECODERS["wavenet_b200"]
            done[mod_name] = True
        except Exception as e:                       # reference (or one of its heavy deps) not importable here
            done[mod_name] = repr(e)
    try:
        bv = importlib.import_module("tasks.tts.vocoder_infer.base_vocoder")
        bv.REGISTERED_VOCODERS["HifiGAN_B200"] = HifiGANB200
        done["vocoder"] = True
    except Exception as e:
        done["vocoder"] = repr(e)
    return done


def build_diffusion(hp: dict, phone_encoder=None, fs=None, mel_encoder=None) -> GaussianDiffusionB200:
    """build_tts_model of the task (tasks/speech_editing/spec_denoiser.py:29-36) with the same hparams keys."""
    return GaussianDiffusionB200(
        phone_encoder=phone_encoder, out_dims=hp["audio_num_mel_bins"],
        denoise_fn=DIFF_DECODERS[hp.get("diff_decoder_type", "wavenet")](hp),
        timesteps=hp["timesteps"], time_scale=hp["timescale"], loss_type=hp["diff_loss_type"],
        spec_min=hp["spec_min"], spec_max=hp["spec_max"], fs=fs, mel_encoder=mel_encoder, hparams=hp)


class SpecDenoiserInferB200:
    """inference/tts/spec_denoiser.py::SpecDenoiserInfer from the tensor boundary on."""

    def __init__(self, hparams_: dict, device=None, fs=None, vocoder: Optional[HifiGANB200] = None, phone_encoder=None):
        """`phone_encoder`: anything with len() = phone vocabulary (the reference's `self.ph_encoder`, inference/tts/spec_denoiser.py:38-41);
        with it (or hparams['b200_vocab']) the native FastSpeechB200 condition encoder is built like the reference builds `fs`;
        `fs=` injects a ready module instead."""
        self.hparams = hparams_
        self.device = torch.device(device or "cuda")
        self._fs = fs
        if phone_encoder is None and fs is None and hparams_.get("b200_vocab"):
            phone_encoder = range(int(hparams_["b200_vocab"]))
        self.ph_encoder = phone_encoder
        self.model = self.build_model()
        self.vocoder = vocoder if vocoder is not None else self.build_vocoder()

    def build_model(self):
        model = build_diffusion(self.hparams, phone_encoder=None if self._fs is not None else self.ph_encoder, fs=self._fs)
        work_dir = self.hparams.get("work_dir", "")
        if work_dir:
            # as the reference (inference/tts/spec_denoiser.py:58: load_ckpt(model, work_dir, 'model'), strict, forced): a missing
            # checkpoint or any missing / unexpected / mis-shaped weight is an error, never a silent random initialisation.  Only
            # the schedule buffers, whose length is `timesteps` + 1, are exempt (a -hp timesteps=100 run of an 8-step checkpoint).
            from .ckpt import SCHEDULE_BUFFERS
            load_ckpt(model, work_dir, "model", force=True, strict=True, drop_keys=SCHEDULE_BUFFERS)
        return model.to(self.device).eval()

    def build_vocoder(self):
        ckpt = self.hparams.get("vocoder_ckpt", "")
        if ckpt and os.path.exists(f"{ckpt}/config.yaml"):
            return HifiGANB200(ckpt)
        raise FileNotFoundError(f"vocoder checkpoint dir '{ckpt}' not found (config.yaml + model_ckpt_steps_*.ckpt)")

    def run_vocoder(self, c: torch.Tensor) -> torch.Tensor:
        """c[B,T,80] -> [B, T*hop]   (inference/tts/base_tts_infer.py:44-47)"""
        return self.vocoder(c.to(self.device))

    @torch.no_grad()
    def forward_model(self, inp: dict):
        """`inp` holds the tensors the reference builds just before calling the model
        (inference/tts/spec_denoiser.py:133-138): either the full condition-encoder inputs (needs `fs`) or a
        ready `cond[B,T,H]`.  Returns (wav_out, mel_out) like the reference's first two outputs per item."""
        if "edited_txt_tokens" in inp:
            return self.edit_forward(inp)
        dev = self.device
        ref = inp["ref_mels"].to(dev)
        mask = inp["time_mel_masks"].to(dev).reshape(ref.shape[0], ref.shape[1], 1)
        if "cond" in inp:
            mel = self.model.sample(inp["cond"].to(dev), inp.get("noise"), int(inp.get("seed", 0)), ref, mask)
        else:
            out = self.model(inp["txt_tokens"].to(dev), time_mel_masks=mask, mel2ph=inp["mel2ph"].to(dev),
                             spk_embed=inp["spk_embed"].to(dev), ref_mels=ref, f0=inp.get("f0"), uv=inp.get("uv"), energy=None,
                             infer=True, use_pred_pitch=inp.get("use_pred_pitch", True), seed=inp.get("seed"))
            mel = out["mel_out"] * mask + ref * (1 - mask)                       # :136
        wav = self.run_vocoder(mel)
        return wav, mel

    @torch.no_grad()
    def edit_forward(self, sample: dict):
        """The reference's forward_model from its `sample` batch on (inference/tts/spec_denoiser.py:63-149, after input_to_batch):
        encoder / style / forward_dur on the EDITED text with the unedited phones' durations given (:84-98), the region surgery
        (:99-131, fse_edit_*), the model call and compositing (:133-136), the vocoder (:137).  `sample` holds the reference's keys
        (edited_txt_tokens, mel, mel2ph, mel2word, dur, ph2word, edited_ph2word, f0, uv, spk_embed, words_region,
        edited_words_region), every tensor with a leading batch axis; B > 1 is allowed when the items share their padded
        lengths (`*_len` tensors may carry the real ones).  Returns (wav_out, mel_out, aux dict)."""
        from . import engine as E
        dev = self.device
        g = lambda k: sample[k].to(dev)
        txt, mel, mel2ph, mel2word = g("edited_txt_tokens"), g("mel"), g("mel2ph"), g("mel2word")
        B = txt.shape[0]
        lens = {k: (sample[k].to(dev) if torch.is_tensor(sample.get(k)) else sample.get(k)) for k in ("T_len", "Tp_len", "Tpe_len", "Te_len")}
        wr, er = sample["words_region"], sample["edited_words_region"]
        if torch.is_tensor(wr):
            regions = torch.cat([wr.reshape(B, 2), er.reshape(B, 2)], 1).long().to(dev)
        else:                                                   # the reference's list-of-tuples form, one region per item
            regions = torch.tensor([[*wr[b if len(wr) > 1 else 0], *er[b if len(er) > 1 else 0]] for b in range(B)], dtype=torch.int64, device=dev)
        fs = self.model.fs
        encoder_out = fs.encoder(txt)
        style = fs.forward_style_embed(g("spk_embed"), None)
        masked_dur, masked_mel2ph, mask_orig = E.edit_prepare(mel2ph, mel2word, g("ph2word"), g("dur"), regions, txt.shape[1],
                                                               lens["T_len"], lens["Tp_len"], lens["Tpe_len"])
        dur_inp = fs.engine().dur_input(encoder_out, None if isinstance(style, int) else style[:, 0], txt)
        ret = {}
        edited_mel2ph = fs.forward_dur(dur_inp, mask_orig, masked_mel2ph, txt, ret, masked_dur=masked_dur, use_pred_mel2ph=True)
        out = E.edit_assemble(mel2ph, mel2word, g("edited_ph2word"), edited_mel2ph, regions, mel, g("f0"), g("uv"), lens["T_len"],
                              lens["Tpe_len"], lens["Te_len"])
        mask = out["time_mel_masks"][:, :, None]
        res = self.model(txt, time_mel_masks=mask, mel2ph=out["mel2ph"], spk_embed=g("spk_embed"), ref_mels=out["ref_mels"], f0=out["f0"],
                         uv=out["uv"], energy=None, infer=True, use_pred_pitch=True, seed=sample.get("seed"), composite=True)
        mel_out = res["mel_out"]                                # already mel_out * mask + ref_mels * (1 - mask)  (:136)
        wav_out = self.run_vocoder(mel_out)
        return wav_out, mel_out, dict(out, dur=ret["dur"], edited_mel2ph_pred=edited_mel2ph, masked_dur=masked_dur,
                                      time_mel_masks_orig=mask_orig)

    def infer_once(self, inp: dict):
        wav, mel = self.forward_model(inp)[:2]
        return wav.cpu().numpy(), mel.cpu().numpy()


class SpeechDenoiserTaskB200:
    """task_cls for `--config egs/spec_denoiser.yaml -hp task_cls=speech_editing_toolkit_b200.plugin.SpeechDenoiserTaskB200`.

    The inference / test leg of the reference task (sampling + vocoder) and the model side of its training step (`run_model(infer=False)`
    -> mel losses, `_training_step`); `start()` runs the inference leg over synthetic editing batches of the configured shape and
    reports throughput (no dataset, no trainer loop)."""

    def __init__(self):
        self.hparams = hparams
        self.model = None
        self.vocoder = None

    def build_model(self):
        """build_tts_model (tasks/speech_editing/spec_denoiser.py:29-36); the phone vocabulary comes from the dataset's
        phone_set.json in the reference — here from hparams['b200_vocab'] when given (native condition encoder), else none."""
        vocab = self.hparams.get("b200_vocab")
        self.model = build_diffusion(self.hparams, phone_encoder=range(int(vocab)) if vocab else None).cuda().eval()
        return self.model

    def build_vocoder(self):
        ckpt = self.hparams.get("vocoder_ckpt", "")
        if ckpt and os.path.exists(f"{ckpt}/config.yaml"):
            self.vocoder = HifiGANB200(ckpt)
        else:                                                      # no shipped checkpoint: seeded HiFi-GAN V1 weights
            self.vocoder = HifiGANB200(state_dict=synth.hifigan_state_dict(self.hparams.get("seed", 1234)))
        return self.vocoder

    def run_model(self, sample: dict, infer: bool = False, *args, **kwargs):
        """SpeechDenoiserTask.run_model (tasks/speech_editing/spec_denoiser.py:39-62): the model call, the masked mel losses
        (`l1_coarse`, `ssim_coarse` with the `mel_losses` weights of the yaml, native loss kernels) and the composited mel_out.
        Returns (losses, output) for infer=False and the output dict for infer=True, as the reference does.  The duration / pitch
        losses of the reference belong to the condition encoder, which is forward-only here: they are not formed (train the
        encoder with the reference task; this drop-in trains the denoiser branch)."""
        from . import train
        target = sample["mels"]
        tmm = sample["time_mel_masks"]
        tmm = tmm[:, :, None] if tmm.dim() == 2 else tmm
        spk = sample.get("spk_embed") if not self.hparams.get("use_spk_id") else sample.get("spk_ids")
        output = self.model(sample["txt_tokens"], tmm, mel2ph=sample["mel2ph"], spk_embed=spk, ref_mels=target, f0=sample.get("f0"),
                            uv=sample.get("uv"), energy=None, infer=infer, **kwargs)
        lam = []
        for item in str(self.hparams.get("mel_losses", "l1:0.5|ssim:0.5")).split("|"):     # parse_mel_losses, tasks/tts/tts_utils.py:21-34
            if item == "":
                continue
            name, _, w = item.partition(":")
            lam.append((name, float(w) if w else 1.0))
        losses = {f"{k}_coarse": v for k, v in train.mel_losses(output["mel_out"] * tmm, target * tmm, tuple(lam)).items()}
        output["mel_out"] = output["mel_out"] * tmm + target * (1 - tmm)
        return output if infer else (losses, output)

    def _training_step(self, sample: dict, batch_idx: int = 0, optimizer_idx: int = -1):
        """SpeechBaseTask._training_step (tasks/tts/speech_base.py:175-179): total loss = sum of the weighted terms that require grad."""
        losses, _ = self.run_model(sample, infer=False)
        total = sum(v for v in losses.values() if isinstance(v, torch.Tensor) and v.requires_grad)
        losses["batch_size"] = sample["txt_tokens"].shape[0]
        return total, losses

    def train_synthetic(self, steps: int):
        """`-hp b200_train_steps=N` (needs `b200_vocab` for the native condition encoder): N optimizer steps of the training step over one
        seeded synthetic batch — `_training_step` (model call + native mel losses), backward through the native DiffNet chain, the optimizer
        the reference builds (AdamW, lr / betas / weight_decay of the yaml; tasks/tts/speech_base.py:110-118), gradients all-reduced over
        the process group when one is initialised; with an experiment directory (`--exp_name`) it resumes from / saves a checkpoint in the
        trainer's layout (ckpt.resume / ckpt.save_ckpt).  A smoke run of the training leg, not the reference's Trainer (no dataset, no
        validation)."""
        import time
        from . import train
        hp = self.hparams
        assert self.model.fs is not None, "training from sample dicts needs the condition encoder: pass -hp b200_vocab=<phone vocabulary size>"
        if not (hp.get("work_dir") and os.path.isdir(hp["work_dir"])):
            vocab = int(hp["b200_vocab"])
            self.model.fs.load_state_dict({k: torch.from_numpy(v) for k, v in synth.fastspeech_state_dict(hp.get("seed", 1234), vocab).items()}, strict=False)
            self.model.mel_encoder.load_state_dict({k: torch.from_numpy(v) for k, v in synth.mel_encoder_state_dict(hp.get("seed", 1234)).items()})
        self.model.train()
        B, T = int(hp.get("max_sentences", 16)), int(hp.get("b200_frames", 1024))
        b = synth.synthetic_edit_batch(hp.get("seed", 1234), B, T, hp["audio_num_mel_bins"], vocab=int(hp.get("b200_vocab", 80)))
        sample = {k: torch.from_numpy(v).cuda() for k, v in b.items()}
        sample["mels"] = sample.pop("ref_mels")
        params = [p for p in self.model.denoise_fn.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(params, lr=float(hp.get("lr", 2e-4)), betas=(float(hp.get("optimizer_adam_beta1", 0.9)), float(hp.get("optimizer_adam_beta2", 0.98))),
                                weight_decay=float(hp.get("weight_decay", 0.0)))
        red = train.BucketedAllReduce(dict(self.model.denoise_fn.named_parameters()))
        step0 = 0
        if hp.get("work_dir"):                                       # resume as the trainer does (weights, optimizer state, step count)
            from .ckpt import SCHEDULE_BUFFERS, resume
            step0, _ = resume(self.model, hp["work_dir"], optimizer=opt, drop_keys=SCHEDULE_BUFFERS)
        log = []
        t0 = None
        for i in range(steps + 1):
            if i == 1:                                               # the first step pays the one-time allocations
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            total, losses = self._training_step(sample, i)
            opt.zero_grad(set_to_none=True)
            total.backward()
            red.finish() if red.active() else None
            opt.step()
            log.append({k: float(v) for k, v in losses.items() if isinstance(v, torch.Tensor)})
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / max(steps, 1)
        print(f"| B200 spec_denoiser training: {steps} steps of {B}x{T} frames, {dt * 1e3:.1f} ms / step ({B * T / dt:.0f} mel-frames/s); "
              f"losses {log[0]} -> {log[-1]}")
        if hp.get("work_dir"):                                       # model_ckpt_steps_<N>.ckpt in the trainer's layout (trainer.py:431-470)
            from .ckpt import save_ckpt
            save_ckpt(self.model, hp["work_dir"], step0 + steps + 1, optimizer=opt, num_ckpt_keep=int(hp.get("num_ckpt_keep", 3)))
        return log

    @torch.no_grad()
    def validation_step(self, sample: dict, batch_idx: int = 0):
        """SpeechDenoiserTask.validation_step (tasks/speech_editing/spec_denoiser.py:64-88): the losses of the training-branch forward
        (`run_model(infer=False)`: q_sample at a random t, one denoiser evaluation, masked mel losses) as python scalars with their sum
        and `nsamples`; for the first `num_valid_plots` batches also a full sampling run composited with the reference mel and, when a
        vocoder was built (validation_start), its waveform for the first item — the tensors the reference hands to its tensorboard
        logger (`plot_wav` / `plot_mel`, which are out of scope) are returned under `mel_out` / `wav_out` instead."""
        losses, _ = self.run_model(sample, infer=False)
        out = {"losses": {k: float(v) for k, v in losses.items()}}
        out["total_loss"] = float(sum(out["losses"].values()))
        out["nsamples"] = int(sample.get("nsamples", sample["txt_tokens"].shape[0]))
        if batch_idx < int(self.hparams.get("num_valid_plots", 0) or 0):
            model_out = self.run_model(sample, infer=True)
            out["mel_out"] = model_out["mel_out"]
            if self.vocoder is not None:
                out["wav_out"] = self.vocoder(model_out["mel_out"][:1])
        return out

    def validation_start(self):
        """SpeechBaseTask.validation_start (tasks/tts/speech_base.py:194-195)."""
        return self.build_vocoder()

    def validation_end(self, outputs):
        """BaseTask.validation_end (utils/commons/base_task.py:154-185): nsamples-weighted means of every loss and of the total,
        rounded to four decimals, as {'tb_log': {'val/<name>': mean}, 'val_loss': mean total}; empty outputs are skipped."""
        sums, counts = {"total_loss": 0.0}, {"total_loss": 0}
        for output in outputs:
            if not output:
                continue
            if isinstance(output, dict):
                if "losses" not in output:
                    raise AssertionError('Key "losses" should exist in validation output.')
                n = output.get("nsamples", 1)
                losses = {k: float(v) for k, v in output["losses"].items()}
                total = float(output.get("total_loss", sum(losses.values())))
            else:
                if len(output) != 2:
                    raise AssertionError("Validation output should only consist of two elements: (total_loss, losses)")
                n, (total, losses) = 1, output
                total, losses = float(total), {k: float(v) for k, v in losses.items()}
            for k, v in list(losses.items()) + [("total_loss", total)]:
                sums[k] = sums.get(k, 0.0) + v * n
                counts[k] = counts.get(k, 0) + n
        means = {k: round(sums[k] / counts[k], 4) if counts[k] else 0.0 for k in sums}
        print(f"| Validation results: {means}")
        return {"tb_log": {f"val/{k}": v for k, v in means.items()}, "val_loss": means["total_loss"]}

    @torch.no_grad()
    def test_step(self, sample: dict, batch_idx: int = 0):
        """speech_editing_base.py:151-192 minus file output: sample -> composite -> vocoder."""
        dev = torch.device("cuda")
        ref = sample["mels"].to(dev)
        mask = sample["time_mel_masks"].to(dev)
        seed = int(sample.get("seed", batch_idx))
        if "cond" in sample:                                       # a ready condition [B,T,H]
            mel = self.model.sample(sample["cond"].to(dev), None, seed, ref, mask)
        else:                                                      # the reference's sample dict: run_model(infer=True), :39-62
            g = lambda k: None if sample.get(k) is None else sample[k].to(dev)
            mel = self.model(g("txt_tokens"), mask.reshape(ref.shape[0], ref.shape[1], 1), g("mel2ph"), g("spk_embed"), ref, g("f0"), g("uv"),
                             infer=True, use_pred_pitch=bool(sample.get("use_pred_pitch", False)), seed=seed, composite=True)["mel_out"]
        wav = self.vocoder(mel)
        return {"mel_out": mel, "wav_out": wav}

    @torch.no_grad()
    def test(self, samples, gen_dir: Optional[str] = None):
        """The reference's test loop over a dataloader / iterable of sample dicts (Trainer.evaluate(test=True) ->
        SpeechEditingBaseTask.test_step, utils/commons/trainer.py:203-251, speech_editing_base.py:151-192): sample -> composite ->
        vocoder per batch; with `gen_dir` the generated mels / wavs are written as `<item_name>.npy` / `<item_name>.wav` (the
        reference writes wav + png through a process pool; plots are out of scope).  Returns the list of per-batch outputs."""
        outs = []
        if gen_dir:
            os.makedirs(gen_dir, exist_ok=True)
        for batch_idx, sample in enumerate(samples):
            out = self.test_step(sample, batch_idx)
            outs.append(out)
            if gen_dir:
                import numpy as np
                from scipy.io import wavfile
                names = sample.get("item_name") or [f"b{batch_idx:04d}_{i}" for i in range(out["mel_out"].shape[0])]
                sr = int(self.hparams.get("audio_sample_rate", 22050))
                for i, name in enumerate(names):
                    np.save(os.path.join(gen_dir, f"{name}.npy"), out["mel_out"][i].cpu().numpy())
                    wavfile.write(os.path.join(gen_dir, f"{name}.wav"), sr, out["wav_out"][i].cpu().numpy())
        return outs

    @classmethod
    def start(cls):
        """Without a dataset this is a SYNTHETIC THROUGHPUT RUN (one seeded batch of max_sentences x b200_frames), not the reference's
        trainer loop; with `-hp b200_test_samples=<file>` (a torch.save'd list of the reference's sample dicts) and `--infer` it
        runs `test()` over them, which is the part of `tasks/run.py --infer` this drop-in covers."""
        import time
        task = cls()
        hp = task.hparams
        task.build_model()
        if hp.get("work_dir") and os.path.isdir(hp["work_dir"]):
            from .ckpt import SCHEDULE_BUFFERS
            load_ckpt(task.model, hp["work_dir"], "model", force=True, strict=True, drop_keys=SCHEDULE_BUFFERS)
        else:
            task.model.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in synth.denoiser_state_dict(
                hp.get("seed", 1234), hp["audio_num_mel_bins"], hp["hidden_size"], hp["residual_channels"], hp["residual_layers"]).items()})
        if int(hp.get("b200_train_steps", 0) or 0) > 0:
            return task.train_synthetic(int(hp["b200_train_steps"]))
        task.build_vocoder()
        if hp.get("b200_test_samples"):
            samples = torch.load(hp["b200_test_samples"], map_location="cpu", weights_only=False)
            outs = task.test(samples, gen_dir=hp.get("b200_gen_dir") or None)
            print(f"| B200 spec_denoiser: test loop over {len(outs)} batches done")
            return outs
        B, T = int(hp.get("max_sentences", 16)), int(hp.get("b200_frames", 1024))
        batch = synth.synthetic_edit_batch(hp.get("seed", 1234), B, T, hp["audio_num_mel_bins"])
        sample = {"cond": torch.from_numpy(synth.synthetic_cond(hp.get("seed", 1234), B, T, hp["hidden_size"])),
                  "mels": torch.from_numpy(batch["ref_mels"]), "time_mel_masks": torch.from_numpy(batch["time_mel_masks"])}
        task.test_step(sample)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = task.test_step(sample, 1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        audio = B * T * hp.get("hop_size", 256) / hp.get("audio_sample_rate", 22050)
        print(f"| B200 spec_denoiser: {B}x{T} frames, {hp['timesteps']} steps + vocoder in {dt * 1e3:.1f} ms "
              f"({B * T / dt:.0f} mel-frames/s, RTF {dt / audio:.5f}); mel {tuple(out['mel_out'].shape)} wav {tuple(out['wav_out'].shape)}")
        return out


class CampNetTaskB200:
    """Inference side of tasks/speech_editing/campnet.py::CampNetTask for
    `--config egs/campnet.yaml -hp task_cls=speech_editing_toolkit_b200.plugin.CampNetTaskB200`:
    build_tts_model (:27-31) and run_model(infer=True) (:52-90: model forward, then mel_out = mel_out_fine * mask + mels * (1 - mask)).
    `start()` runs one synthetic batch of the configured shape (max_sentences x b200_frames) and reports throughput."""

    def __init__(self, ph_dict_size: Optional[int] = None, word_dict_size: Optional[int] = None):
        self.hparams = hparams
        self.ph_dict_size = int(ph_dict_size if ph_dict_size is not None else hparams.get("b200_vocab", 80))
        self.word_dict_size = word_dict_size
        self.model = None

    def build_tts_model(self):
        self.model = CampNetB200(self.ph_dict_size, self.word_dict_size, self.hparams)
        return self.model

    @torch.no_grad()
    def run_model(self, sample: dict, infer: bool = True, *args, **kwargs):
        if not infer:
            raise NotImplementedError("CampNetTaskB200 provides the inference side; training stays on the reference task")
        time_mel_masks = sample["time_mel_masks"][:, :, None] if sample["time_mel_masks"].dim() == 2 else sample["time_mel_masks"]
        output = self.model(sample["txt_tokens"], spk_embed=sample.get("spk_embed"), spk_id=sample.get("spk_ids"), mels=sample["mels"],
                            stutter_mel_masks=None, time_mel_masks=time_mel_masks, infer=True)
        output["mel_out"] = output["mel_out_fine"] * time_mel_masks + sample["mels"] * (1 - time_mel_masks)
        return output          # infer=True returns the output dict alone (tasks/speech_editing/campnet.py:77-88; test_step indexes it)

    @classmethod
    def start(cls):
        import time
        task = cls()
        hp = task.hparams
        model = task.build_tts_model().cuda().eval()
        if hp.get("work_dir") and os.path.isdir(hp["work_dir"]):
            load_ckpt(model, hp["work_dir"], "model", force=True, strict=True)
        else:
            model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.campnet_state_dict(hp.get("seed", 1234), task.ph_dict_size,
                                                                                             hp["hidden_size"]).items()}, strict=False)
        B, T = int(hp.get("max_sentences", 16)), int(hp.get("b200_frames", 1024))
        b = synth.synthetic_campnet_batch(hp.get("seed", 1234), B, T, vocab=task.ph_dict_size)
        sample = {k: torch.from_numpy(v).cuda() for k, v in b.items()}
        task.run_model(sample)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = task.run_model(sample)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"| B200 CampNet: {B}x{T} frames mask-predict forward in {dt * 1e3:.1f} ms ({B * T / dt:.0f} mel-frames/s); "
              f"mel_out {tuple(out['mel_out'].shape)} attn {tuple(out['attn'].shape)}")
        return out


def run_task():
    """tasks/run.py:9-14 — import hparams['task_cls'] by dotted path and call .start()."""
    assert hparams["task_cls"] != ""
    pkg, cls_name = hparams["task_cls"].rsplit(".", 1)
    getattr(importlib.import_module(pkg), cls_name).start()


if __name__ == "__main__":
    set_hparams()
    run_task()
