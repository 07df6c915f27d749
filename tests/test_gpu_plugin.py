"""The reference-facing plugin layer on the GPU: same calls a user of the reference makes
(DiffNet.forward, GaussianDiffusion.forward(infer=True) / p_sample, vocoder.spec2wav, run_vocoder)."""
import numpy as np
import pytest

from conftest import golden, rel_l1

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

HP = dict(audio_num_mel_bins=80, hidden_size=192, residual_layers=20, residual_channels=256, dilation_cycle_length=1,
          timesteps=10, timescale=1, diff_loss_type="l1", spec_min=[], spec_max=[], keep_bins=80, schedule_type="vpsde",
          diff_decoder_type="wavenet_b200", b200_mode="tc_bf16")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


class StubFS(torch.nn.Module):
    """Stands in for the reference FastSpeech condition encoder: returns a fixed decoder_inp."""

    def __init__(self, cond):
        super().__init__()
        self.register_buffer("cond", cond)

    def forward(self, txt_tokens, time_mel_masks, mel2ph, spk_embed, f0, uv, energy, **kw):
        assert kw["skip_decoder"] and kw["infer"]
        return {"decoder_inp": self.cond.clone(), "dur": None, "mel2ph": mel2ph}


def _model(seed=1234, S=10):
    from speech_editing_toolkit_b200 import plugin, synth
    m = plugin.build_diffusion(dict(HP, timesteps=S))
    m.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in synth.denoiser_state_dict(seed).items()})
    return m.cuda().eval()


def test_diffnet_module_forward_matches_reference_fixture(lib_built):
    _need_gpu()
    from speech_editing_toolkit_b200 import synth
    g = golden("diffnet_step.npz")
    m = _model()
    cond = torch.from_numpy(synth.synthetic_cond(int(g["seed"]), int(g["B"]), int(g["T"]))).cuda()
    out = m.denoise_fn(torch.from_numpy(g["x"]).cuda()[:, None], torch.from_numpy(g["t"]).cuda(), cond.transpose(1, 2))
    assert out.shape == (int(g["B"]), 1, 80, int(g["T"]))
    assert rel_l1(out[:, 0].cpu().numpy(), g["x0"]) < 2e-2
    # changing a parameter invalidates the repacked weights
    with torch.no_grad():
        m.denoise_fn.output_projection.bias.add_(1.0)
    out2 = m.denoise_fn(torch.from_numpy(g["x"]).cuda()[:, None], torch.from_numpy(g["t"]).cuda(), cond.transpose(1, 2))
    assert np.allclose((out2 - out).cpu().numpy(), 1.0, atol=1e-5)


def test_gaussian_diffusion_forward_infer_matches_oracle(lib_built):
    _need_gpu()
    from oracle import fluentspeech_oracle as O
    from speech_editing_toolkit_b200 import synth
    g = golden("sample_c1.npz")
    seed, B, T, S = int(g["seed"]), int(g["B"]), int(g["T"]), int(g["S"])
    m = _model(S=S)
    cond = torch.from_numpy(synth.synthetic_cond(seed, B, T)).cuda()
    m.fs = StubFS(cond)
    with torch.no_grad():
        m.mel_encoder.fc_out.weight.zero_(); m.mel_encoder.fc_out.bias.zero_()
    noise = torch.from_numpy(synth.synthetic_noise(seed, S, B, T)).cuda()
    batch = synth.synthetic_edit_batch(1, B, T)
    ret = m(torch.from_numpy(batch["txt_tokens"]).cuda(), torch.from_numpy(batch["time_mel_masks"]).cuda()[:, :, None],
            torch.from_numpy(batch["mel2ph"]).cuda(), torch.from_numpy(batch["spk_embed"]).cuda(),
            torch.from_numpy(batch["ref_mels"]).cuda(), None, None, infer=True, noise=noise)
    assert set(ret) >= {"mel_out", "decoder_inp", "mel2ph"}
    assert rel_l1(ret["mel_out"].cpu().numpy(), g["mel_out"]) < 2e-2
    # p_sample = denoise + posterior, one step, against the oracle
    xt = noise[0][:, None]
    t = torch.full((B,), S - 1, dtype=torch.long, device="cuda")
    x1 = m.p_sample(xt, t, cond.transpose(1, 2), noise=noise[1][:, None])
    assert rel_l1(x1[:, 0].cpu().numpy(), g["x_after_first"]) < 2e-2
    # Philox path: seeded by torch.manual_seed through forward(seed=None)
    torch.manual_seed(5); a = m(None, torch.zeros(B, T, 1, device="cuda"), torch.ones(B, T, device="cuda"), None,
                               torch.zeros(B, T, 80, device="cuda"), None, None, infer=True)["mel_out"]
    torch.manual_seed(5); b = m(None, torch.zeros(B, T, 1, device="cuda"), torch.ones(B, T, device="cuda"), None,
                               torch.zeros(B, T, 80, device="cuda"), None, None, infer=True)["mel_out"]
    assert torch.equal(a, b) and torch.isfinite(a).all()


def test_vocoder_plugin_spec2wav_and_infer_class(lib_built, tmp_path):
    _need_gpu()
    import yaml
    from speech_editing_toolkit_b200 import plugin, synth
    from speech_editing_toolkit_b200.engine import HIFIGAN_V1
    from speech_editing_toolkit_b200.vocoder import get_vocoder_cls
    g = golden("hifigan_v1.npz")
    sd = synth.hifigan_state_dict(int(g["seed"]))
    yaml.safe_dump(dict(HIFIGAN_V1, audio_num_mel_bins=80), open(tmp_path / "config.yaml", "w"))
    torch.save({"state_dict": {"model_gen": {k: torch.from_numpy(v) for k, v in sd.items()}}}, tmp_path / "model_ckpt_steps_5.ckpt")
    voc = get_vocoder_cls("HifiGAN_B200")(str(tmp_path))
    wav = voc.spec2wav(g["mel"][0])                                  # numpy [T,80] -> numpy [T*256]
    assert wav.shape == (g["mel"].shape[1] * 256,) and wav.dtype == np.float32
    assert rel_l1(wav, g["wav"][0]) < 3e-2
    hp = dict(HP, timesteps=4, vocoder_ckpt=str(tmp_path), work_dir="")
    inf = plugin.SpecDenoiserInferB200(hp, vocoder=voc)
    inf.model.denoise_fn.load_state_dict({k: torch.from_numpy(v) for k, v in synth.denoiser_state_dict(1234).items()})
    B, T = 2, 64
    batch = synth.synthetic_edit_batch(2, B, T)
    wav_out, mel_out = inf.infer_once({"cond": torch.from_numpy(synth.synthetic_cond(2, B, T)), "ref_mels": torch.from_numpy(batch["ref_mels"]),
                                       "time_mel_masks": torch.from_numpy(batch["time_mel_masks"]), "seed": 3})
    assert wav_out.shape == (B, T * 256) and mel_out.shape == (B, T, 80) and np.isfinite(wav_out).all()
    m = batch["time_mel_masks"].astype(bool)
    assert np.array_equal(mel_out[~m], batch["ref_mels"][~m])


def _nccl_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from speech_editing_toolkit_b200 import dist as fdist, synth
        m = _model(S=2)
        B, T = 6, 128
        cond = torch.from_numpy(synth.synthetic_cond(4, B, T)).cuda()
        noise = torch.from_numpy(synth.synthetic_noise(4, 2, B, T)).cuda()
        full = m.sample(cond, noise)
        lo, hi = fdist.shard_bounds(B, rank, world)
        out = fdist.all_gather_batch(m.sample(cond[lo:hi], noise[:, lo:hi]), B)
        q.put((rank, bool(torch.equal(out, full))))
    finally:
        dist.destroy_process_group()


def test_batch_sharded_sampling_two_gpus_nccl(lib_built):
    _need_gpu()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=600) for _ in ps]
    [p.join(60) for p in ps]
    assert all(ok for _, ok in res)


def test_task_start_runs_the_inference_and_the_training_leg_from_the_reference_yaml(lib_built, capsys):
    """tasks/run.py contract: hparams['task_cls'].start() with the reference's own egs/spec_denoiser.yaml (from the reference copy that travels
    as oracle/_ref) and -hp overrides: the synthetic inference run, then `b200_train_steps=3` — three optimizer steps through
    _training_step / the native training chain, with finite, decreasing-or-equal-order losses reported."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import os
    from oracle import refshim
    if not refshim.available():
        pytest.skip("no reference tree (oracle/_ref) on this box")
    cfg = os.path.join(refshim.REF_ROOT, "egs", "spec_denoiser.yaml")
    from speech_editing_toolkit_b200 import plugin
    from speech_editing_toolkit_b200.hparams import set_hparams
    cwd = os.getcwd()
    os.chdir(refshim.REF_ROOT)                                        # base_config paths of the yaml are relative to the reference root
    try:
        set_hparams(cfg, hparams_str="task_cls=speech_editing_toolkit_b200.plugin.SpeechDenoiserTaskB200,timesteps=4,max_sentences=2,b200_frames=128,"
                                     "residual_layers=4,b200_vocab=80", print_hparams=False)
        out = plugin.SpeechDenoiserTaskB200.start()
        assert tuple(out["mel_out"].shape) == (2, 128, 80) and tuple(out["wav_out"].shape) == (2, 128 * 256) and bool(torch.isfinite(out["wav_out"]).all())
        set_hparams(cfg, hparams_str="task_cls=speech_editing_toolkit_b200.plugin.SpeechDenoiserTaskB200,timesteps=100,max_sentences=2,b200_frames=128,"
                                     "residual_layers=4,b200_vocab=80,b200_train_steps=3,b200_mode=tc_bf16", print_hparams=False)
        log = plugin.SpeechDenoiserTaskB200.start()
    finally:
        os.chdir(cwd)
    assert len(log) == 4 and all(np.isfinite(list(e.values())).all() for e in log) and set(log[0]) == {"l1_coarse", "ssim_coarse"}
    print("[margin] task start(): inference leg ok, training leg losses", log[0], "->", log[-1])


def test_task_validation_step_on_a_synthetic_batch(lib_built):
    """SpeechDenoiserTaskB200.validation_start / validation_step / validation_end (tasks/speech_editing/spec_denoiser.py:64-88,
    tasks/tts/speech_base.py:194-211, utils/commons/base_task.py:154-185) over a seeded synthetic editing batch with the reference's own
    yaml: finite scalar losses of the training-branch forward, the sampled + composited mel for the first num_valid_plots batches
    (unedited frames = the input mel, bit for bit), the vocoder's waveform of its first item."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import os
    from oracle import refshim
    if not refshim.available():
        pytest.skip("no reference tree (oracle/_ref) on this box")
    cfg = os.path.join(refshim.REF_ROOT, "egs", "spec_denoiser.yaml")
    from speech_editing_toolkit_b200 import plugin, synth
    from speech_editing_toolkit_b200.hparams import set_hparams
    cwd = os.getcwd()
    os.chdir(refshim.REF_ROOT)
    try:
        set_hparams(cfg, hparams_str="task_cls=speech_editing_toolkit_b200.plugin.SpeechDenoiserTaskB200,timesteps=4,max_sentences=2,b200_frames=128,"
                                     "residual_layers=4,b200_vocab=80,num_valid_plots=1", print_hparams=False)
    finally:
        os.chdir(cwd)
    task = plugin.SpeechDenoiserTaskB200()
    hp = task.hparams
    task.build_model()
    t_ = lambda sd: {k: torch.from_numpy(v) for k, v in sd.items()}
    task.model.denoise_fn.load_state_dict(t_(synth.denoiser_state_dict(1234, hp["audio_num_mel_bins"], hp["hidden_size"], hp["residual_channels"], hp["residual_layers"])))
    task.model.fs.load_state_dict(t_(synth.fastspeech_state_dict(1234, 80)), strict=False)
    task.model.mel_encoder.load_state_dict(t_(synth.mel_encoder_state_dict(1234)))
    task.validation_start()
    B, T = 2, 128
    b = synth.synthetic_edit_batch(1234, B, T, hp["audio_num_mel_bins"], vocab=80)
    sample = {k: torch.from_numpy(v).cuda() for k, v in b.items()}
    sample["mels"] = sample.pop("ref_mels")
    sample["nsamples"] = B
    torch.manual_seed(0)
    o0, o1 = task.validation_step(sample, 0), task.validation_step(sample, 1)
    for o in (o0, o1):
        assert set(o["losses"]) == {"l1_coarse", "ssim_coarse"} and np.isfinite(list(o["losses"].values())).all()
        assert abs(o["total_loss"] - sum(o["losses"].values())) < 1e-6 and o["nsamples"] == B
    assert "mel_out" not in o1 and tuple(o0["mel_out"].shape) == (B, T, 80) and tuple(o0["wav_out"].shape) == (1, T * 256)
    assert bool(torch.isfinite(o0["mel_out"]).all()) and bool(torch.isfinite(o0["wav_out"]).all())
    keep = sample["time_mel_masks"] == 0
    assert torch.equal(o0["mel_out"][keep], sample["mels"][keep])
    end = task.validation_end([o0, o1])
    assert np.isfinite(end["val_loss"]) and set(end["tb_log"]) == {"val/total_loss", "val/l1_coarse", "val/ssim_coarse"}
    print("[margin] task validation_step: losses", o0["losses"], o1["losses"], "val_loss", end["val_loss"])
