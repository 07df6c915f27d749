"""Run and time the UNMODIFIED reference modules of the hot path (imported by oracle/refshim.py from /root/reference, or on
the GPU box from oracle/_ref, its verbatim git-ignored copy — see oracle/build_ref.py).

TEST / MEASUREMENT INFRASTRUCTURE ONLY: bench.py's reference arm (`--impl reference`), `cpu_baseline` and
`eager_gpu_baseline` legs, and tests.  Never imported by the product package.

What runs is the reference's own code, stock path:
  * `GaussianDiffusion` (modules/speech_editing/spec_denoiser/spec_denoiser.py:14-196) holding the reference `FastSpeech`
    (fs.py), `MelEncoder` (mel_encoder.py) and `DiffNet` (diffnet.py:84-132), called through `forward(..., infer=True)`
    (:154-185) or, for a bounded sample of the 100-step loop, through the same three calls that forward makes:
    `fs(..., skip_decoder=True)` + `mel_encoder` (:159-164), then `p_sample(x, t, cond)` (:103-108, :181-182) K times.
  * `HifiGanGenerator.forward` (modules/vocoder/hifigan/hifigan.py:126-142) with the weight-norm hooks left in place, as the
    reference's wrapper does (tasks/tts/vocoder_infer/hifigan.py:13-21 never calls remove_weight_norm).
Weights are the seeded synthetic state_dicts of speech_editing_toolkit_b200/synth.py (pretrained checkpoints are Google-Drive
artefacts), loaded with the reference's own `load_state_dict`.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def available() -> bool:
    from oracle import refshim
    return refshim.available()


def _tt(sd):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def build_models(timesteps: int, device="cpu", vocab: int = 80, seed: int = 1234, vocoder: bool = True):
    """(GaussianDiffusion, HifiGanGenerator | None), reference classes with synthetic weights, eval mode, on `device`."""
    import torch
    from oracle import refshim
    from oracle.fluentspeech_oracle import HIFIGAN_V1
    from speech_editing_toolkit_b200 import synth
    hp = refshim.install("egs/spec_denoiser.yaml", overrides=f"timesteps={int(timesteps)}")
    from modules.speech_editing.spec_denoiser.diffnet import DiffNet
    from modules.speech_editing.spec_denoiser import spec_denoiser as sdmod
    from modules.vocoder.hifigan.hifigan import HifiGanGenerator
    net = DiffNet(hp["audio_num_mel_bins"])
    net.load_state_dict(_tt(synth.denoiser_state_dict(seed)), strict=True)
    model = sdmod.GaussianDiffusion(phone_encoder=list(range(vocab)), out_dims=80, denoise_fn=net, timesteps=int(timesteps),
                                    time_scale=hp["timescale"], loss_type=hp["diff_loss_type"], spec_min=hp["spec_min"],
                                    spec_max=hp["spec_max"])
    missing, unexpected = model.fs.load_state_dict(_tt(synth.fastspeech_state_dict(seed, vocab)), strict=False)
    assert not unexpected and all(k.startswith(("decoder.", "mel_out.")) for k in missing), (missing, unexpected)
    model.mel_encoder.load_state_dict(_tt(synth.mel_encoder_state_dict(seed)), strict=True)
    model = model.to(device).eval()
    gen = None
    if vocoder:
        gen = HifiGanGenerator(dict(HIFIGAN_V1))
        gen.load_state_dict(_tt(synth.hifigan_state_dict(seed)), strict=True)
        gen = gen.to(device).eval()
    return model, gen


def synthetic_inputs(B: int, T: int, device="cpu", seed: int = 1000, vocab: int = 80):
    import torch
    from speech_editing_toolkit_b200 import synth
    batch = synth.synthetic_edit_batch(seed, B, T, vocab=vocab)
    return {k: torch.from_numpy(v).to(device) for k, v in batch.items()}


def run_condition(model, tb):
    """The once-per-utterance part of GaussianDiffusion.forward (spec_denoiser.py:159-167), by the same calls."""
    ret = model.fs(tb["txt_tokens"], tb["time_mel_masks"][:, :, None], tb["mel2ph"], tb["spk_embed"], tb["f0"], tb["uv"],
                   skip_decoder=True, infer=True, use_pred_pitch=True)
    decoder_inp = ret["decoder_inp"]
    tgt_nonpadding = (tb["mel2ph"] > 0).float()[:, :, None]
    decoder_inp = decoder_inp + model.mel_encoder(tb["ref_mels"] * (1 - tb["time_mel_masks"][:, :, None])) * tgt_nonpadding
    return decoder_inp.transpose(1, 2)


def time_reference(B: int, T: int, timesteps: int, iters: int = 3, Bv: int = 1, Tv: int = 256, device="cpu", threads=None,
                   models=None):
    """Bounded sample of `timesteps`-step sampling + HiFi-GAN on the reference's own modules.
    Returns seconds per mel frame split into condition encoder / one loop iteration / vocoder, and the totals."""
    import torch
    cuda = str(device).startswith("cuda")
    if threads and not cuda:
        torch.set_num_threads(int(threads))
    model, gen = models if models is not None else build_models(timesteps, device)
    tb = synthetic_inputs(B, T, device)

    def sync():
        if cuda:
            torch.cuda.synchronize()

    def clock(fn, reps=1):
        sync()
        t0 = time.perf_counter()
        out = None
        for _ in range(reps):
            out = fn()
        sync()
        return (time.perf_counter() - t0) / reps, out

    with torch.no_grad():
        cond = run_condition(model, tb)                                       # warm-up (allocator, oneDNN / cuDNN heuristics)
        t_cond, cond = clock(lambda: run_condition(model, tb))
        x = torch.randn(B, 1, 80, T, device=device)                            # spec_denoiser.py:179-180
        tt = torch.full((B,), timesteps - 1, device=device, dtype=torch.long)
        x = model.p_sample(x, tt, cond)                                        # warm-up
        state = {"x": x}

        def loop():
            for i in range(iters):                                             # spec_denoiser.py:181-182, K of S iterations
                state["x"] = model.p_sample(state["x"], torch.full((B,), timesteps - 2 - i, device=device, dtype=torch.long), cond)
            return state["x"]
        t_loop, x = clock(loop)
        t_voc = 0.0
        if gen is not None:
            mel = torch.randn(Bv, 80, Tv, device=device) * 1.5 - 3.0
            gen(mel[:, :, :32])                                                # warm-up
            if cuda:
                gen(mel)
            t_voc, _ = clock(lambda: gen(mel))
    per_frame = {"cond": t_cond / (B * T), "iter": t_loop / (iters * B * T), "vocoder": t_voc / (Bv * Tv) if gen is not None else 0.0}
    sec_per_frame = per_frame["cond"] + timesteps * per_frame["iter"] + per_frame["vocoder"]
    return {"sec_per_frame": sec_per_frame, "frames_per_s": 1.0 / sec_per_frame, "per_frame": per_frame,
            "threads": (torch.get_num_threads() if not cuda else None),
            "sample": (f"reference modules ({'oracle/_ref' if 'oracle' in __import__('oracle.refshim', fromlist=['x']).REF_ROOT else '/root/reference'}): "
                       f"fs+mel_encoder B={B}xT={T} ({per_frame['cond'] * 1e6:.2f} us/frame) + {iters} of {timesteps} p_sample iterations "
                       f"B={B}xT={T} ({per_frame['iter'] * 1e6:.3f} us/frame-iteration)"
                       + (f" + HifiGanGenerator B={Bv}xT={Tv} ({per_frame['vocoder'] * 1e6:.2f} us/frame)" if gen is not None else "")
                       + ", extrapolated to cond + S x iteration + vocoder")}


def full_forward(B: int, T: int, timesteps: int, device="cpu"):
    """The reference's public call, un-cut: GaussianDiffusion.forward(infer=True) then HifiGanGenerator.forward. Small shapes."""
    import torch
    model, gen = build_models(timesteps, device)
    tb = synthetic_inputs(B, T, device)
    with torch.no_grad():
        t0 = time.perf_counter()
        ret = model(tb["txt_tokens"], tb["time_mel_masks"][:, :, None], tb["mel2ph"], tb["spk_embed"], tb["ref_mels"], tb["f0"], tb["uv"],
                    infer=True, use_pred_pitch=True)
        mel = ret["mel_out"] * tb["time_mel_masks"][:, :, None] + tb["ref_mels"] * (1 - tb["time_mel_masks"][:, :, None])
        wav = gen(mel.transpose(1, 2))
        if str(device).startswith("cuda"):
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return {"seconds": dt, "frames_per_s": B * T / dt, "mel": mel, "wav": wav}


if __name__ == "__main__":
    import json
    r = time_reference(2, 256, 100, iters=2, Bv=1, Tv=64)
    print(json.dumps({k: v for k, v in r.items()}, indent=1))
