"""Training step: the native mel losses, and forward and gradients of `DiffNetB200` under autograd — native forward and
activation-gradient chain (fse_train_*), weight gradients by fse_wgrad (tcgen05, MN-major operands) — against tests/golden/diffnet_train.npz, the gradients
torch.autograd computes through the UNMODIFIED reference DiffNet (oracle/make_golden.py train).

Stated tolerances (relative L2 against the fp32 CPU autograd of the reference): FSE_MODE_SIMT_F32 <= 1e-4 on the output and on every
gradient.  The tensor-core modes: output <= 5e-3 (tf32) / 2e-2 (bf16); gradients <= 4e-2 (tf32) / 1.2e-1 (bf16).  Gradients are an
order of magnitude more sensitive than the forward (the gate's derivative factors s(1-s), 1-tanh^2 amplify pre-activation errors
where the gate saturates): `test_reference_gpu_path_gradient_deviation` measures what the reference's OWN GPU path (its modules
after .cuda(), cuDNN TF32 convolutions) deviates from the same fp32 fixture, which is the yardstick for the tf32 number."""
import numpy as np
import pytest

from conftest import golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

HP = dict(audio_num_mel_bins=80, hidden_size=192, residual_channels=256, dilation_cycle_length=1)
TOL = {"simt_f32": 1e-4, "tc_tf32": 5e-3, "tc_bf16": 2e-2}
GTOL = {"simt_f32": 1e-4, "tc_tf32": 4e-2, "tc_bf16": 1.2e-1}


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-30))


def _module(g, mode):
    from speech_editing_toolkit_b200 import synth
    from speech_editing_toolkit_b200.modules import DiffNetB200
    L = int(g["layers"])
    net = DiffNetB200(80, dict(HP, residual_layers=L, b200_mode=mode)).cuda().train()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.denoiser_state_dict(int(g["seed"]), layers=L).items()})
    return net


@pytest.mark.parametrize("mode", ["simt_f32", "tc_tf32", "tc_bf16"])
def test_diffnet_gradients_vs_reference_autograd_fixture(lib_built, mode):
    _need_gpu()
    from speech_editing_toolkit_b200 import synth
    g = golden("diffnet_train.npz")
    B, T = int(g["B"]), int(g["T"])
    net = _module(g, mode)
    cond = torch.from_numpy(synth.synthetic_cond(int(g["seed"]) + 2, B, T)).cuda().requires_grad_(True)
    x = torch.from_numpy(g["x"]).cuda()
    x0 = net(x[:, None], torch.from_numpy(g["t"]).cuda(), cond.transpose(1, 2))[:, 0]
    x0.backward(torch.from_numpy(g["dx0"]).cuda())
    tol = TOL[mode]
    e_fwd, e_cond = rel_l2(x0.detach().cpu().numpy(), g["x0"]), rel_l2(cond.grad.cpu().numpy(), g["dcond"])
    errs = {}
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        got = p.grad.detach().cpu().numpy().reshape(-1)
        want = g["g__" + name]
        got_s = got if got.size <= 4096 else got[::61]
        errs[name] = rel_l2(got_s, want)
        n = float(np.sqrt((got.astype(np.float64) ** 2).sum()))
        assert abs(n - float(g["gnorm__" + name])) <= 2 * GTOL[mode] * float(g["gnorm__" + name]) + 1e-12, (name, n, float(g["gnorm__" + name]))
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print(f"[margin] DiffNet training {mode}: x0 rel-L2 {e_fwd:.3e}, dcond {e_cond:.3e}, median parameter-gradient rel-L2 "
          f"{float(np.median(list(errs.values()))):.3e}, worst: " + ", ".join(f"{k} {v:.2e}" for k, v in top))
    assert e_fwd < tol and e_cond < GTOL[mode]
    assert top[0][1] < GTOL[mode], top


def test_reference_gpu_path_gradient_deviation(lib_built):
    """Context for the tolerances above: the unmodified reference DiffNet on this GPU (eager torch, cuDNN TF32 convolutions as torch
    defaults) against the same fp32 CPU fixture.  Skipped where no reference copy exists."""
    _need_gpu()
    from oracle import refshim
    if not refshim.available():
        pytest.skip("no reference tree (oracle/_ref) on this box")
    from speech_editing_toolkit_b200 import synth
    g = golden("diffnet_train.npz")
    L, B, T = int(g["layers"]), int(g["B"]), int(g["T"])
    hp = refshim.install("egs/spec_denoiser.yaml", overrides=f"timesteps=100,residual_layers={L}")
    from utils.commons.hparams import hparams as ref_hparams
    ref_hparams["residual_layers"] = L
    from modules.speech_editing.spec_denoiser.diffnet import DiffNet
    net = DiffNet(hp["audio_num_mel_bins"]).train()
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in synth.denoiser_state_dict(int(g["seed"]), layers=L).items()}, strict=True)
    net = net.cuda()
    cond = torch.from_numpy(synth.synthetic_cond(int(g["seed"]) + 2, B, T)).cuda().requires_grad_(True)
    x0 = net(torch.from_numpy(g["x"]).cuda()[:, None], torch.from_numpy(g["t"]).cuda(), cond.transpose(1, 2))[:, 0]
    x0.backward(torch.from_numpy(g["dx0"]).cuda())
    errs = {}
    for name, p in net.named_parameters():
        got = p.grad.detach().cpu().numpy().reshape(-1)
        errs[name] = rel_l2(got if got.size <= 4096 else got[::61], g["g__" + name])
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:3]
    print(f"[margin] reference modules on the GPU (cudnn.allow_tf32={torch.backends.cudnn.allow_tf32}): x0 rel-L2 "
          f"{rel_l2(x0.detach().cpu().numpy(), g['x0']):.3e}, dcond {rel_l2(cond.grad.cpu().numpy(), g['dcond']):.3e}, median parameter-gradient "
          f"{float(np.median(list(errs.values()))):.3e}, worst: " + ", ".join(f"{k} {v:.2e}" for k, v in top))


def test_train_step_decreases_the_loss_and_is_deterministic(lib_built):
    """A few AdamW steps of the denoiser branch of the training step (train.train_step: q_sample, DiffNet, masked l1 + ssim, backward,
    optimizer) on one synthetic batch with fixed t / noise: the loss goes down, and two runs from the same state give the same losses."""
    _need_gpu()
    from speech_editing_toolkit_b200 import schedule, synth, train
    from speech_editing_toolkit_b200.modules import DiffNetB200
    L, B, T = 4, 2, 256
    sched = {k: torch.from_numpy(v).cuda() for k, v in schedule.diffusion_buffers(100).items()}
    batch = synth.synthetic_edit_batch(5, B, T)
    data = {"ref_mels": torch.from_numpy(batch["ref_mels"]).cuda(), "time_mel_masks": torch.from_numpy(batch["time_mel_masks"]).cuda(),
            "cond": torch.from_numpy(synth.synthetic_cond(5, B, T)).cuda()}
    t = torch.tensor([10, 60], device="cuda")
    noise = torch.randn(B, 1, 80, T, generator=torch.Generator().manual_seed(0)).cuda()
    runs = []
    for _ in range(2):
        net = DiffNetB200(80, dict(HP, residual_layers=L, b200_mode="tc_bf16")).cuda().train()
        net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.denoiser_state_dict(9, layers=L).items()})
        opt = torch.optim.AdamW(net.parameters(), lr=2e-3, betas=(0.9, 0.98), weight_decay=0.0)
        runs.append([train.train_step(net, sched, data, opt, t=t, noise=noise)["total"] for _ in range(8)])
    assert runs[0][-1] < 0.9 * runs[0][0], runs[0]
    assert np.allclose(runs[0], runs[1], rtol=1e-4), (runs[0], runs[1])
    print(f"[margin] train_step: loss {runs[0][0]:.4f} -> {runs[0][-1]:.4f} in 8 steps")


def test_mel_loss_kernels_vs_reference_fixture(lib_built):
    """fse_mel_loss_forward / fse_mel_loss_backward against the reference's own l1_loss / ssim_loss and their autograd gradient
    (tests/golden/mel_loss.npz).  Stated tolerance: losses 2e-6 absolute, gradient rel-L2 1e-4 against the fp64 run of the reference code
    (the reference's own fp32 run is printed beside it: sigma = E[a^2] - mu^2 on values near 6 cancels ~5 digits in either)."""
    _need_gpu()
    from speech_editing_toolkit_b200 import train
    g = golden("mel_loss.npz")
    a = torch.from_numpy(g["mel_out"]).cuda().requires_grad_(True)
    out = train.mel_losses(a, torch.from_numpy(g["target"]).cuda())
    (out["l1"] + out["ssim"]).backward()
    e_l1, e_ss = abs(float(out["l1"]) - float(g["l1_f64"])), abs(float(out["ssim"]) - float(g["ssim_f64"]))
    e_g = rel_l2(a.grad.cpu().numpy(), g["grad_f64"])
    print(f"[margin] mel losses: l1 abs err {e_l1:.2e}, ssim abs err {e_ss:.2e}, gradient rel-L2 {e_g:.2e} vs the reference in fp64 "
          f"(the reference's fp32 run: {abs(float(g['l1_f32']) - float(g['l1_f64'])):.2e}, {abs(float(g['ssim_f32']) - float(g['ssim_f64'])):.2e}, "
          f"{rel_l2(g['grad_f32'], g['grad_f64']):.2e})")
    assert e_l1 < 2e-6 and e_ss < 2e-6 and e_g < 1e-4


def test_mel_loss_kernels_at_training_shape_are_deterministic_and_match_the_torch_restatement(lib_built):
    """32 x 1024 x 80 (ragged: padded tails, masked spans) against oracle.train_oracle.mel_losses_torch in fp64 on the same device;
    separate upstream weights for the two terms; two runs must agree bit for bit (no floating-point atomics)."""
    _need_gpu()
    from oracle.train_oracle import mel_losses_torch
    from speech_editing_toolkit_b200 import train
    B, T, M = 32, 1024, 80
    gen = torch.Generator().manual_seed(11)
    target = (torch.randn(B, T, M, generator=gen) * 1.5 - 3.0).clamp(-6.0, 1.5)
    out = target + 0.3 * torch.randn(B, T, M, generator=gen)
    mask = torch.zeros(B, T, 1)
    for b in range(B):
        lo = (37 * b) % 500
        mask[b, lo:lo + 200 + 11 * b] = 1
    mask[3] = 0                                                      # an utterance without any speech frame in the mask
    out, target = (out * mask).cuda(), (target * mask).cuda()
    runs = []
    for _ in range(2):
        a = out.clone().requires_grad_(True)
        o = train.mel_losses(a, target)
        (0.7 * o["l1"] + 1.3 * o["ssim"]).backward()
        runs.append((o["l1"].item(), o["ssim"].item(), a.grad.clone()))
    assert runs[0][0] == runs[1][0] and runs[0][1] == runs[1][1] and torch.equal(runs[0][2], runs[1][2])
    a64 = out.double().requires_grad_(True)
    o64 = mel_losses_torch(a64, target.double())
    (0.7 * o64["l1"] + 1.3 * o64["ssim"]).backward()
    e_g = rel_l2(runs[0][2].cpu().numpy(), a64.grad.cpu().numpy())
    print(f"[margin] mel losses 32 x 1024: l1 {runs[0][0]:.6f} vs {o64['l1'].item():.6f}, ssim {runs[0][1]:.6f} vs {o64['ssim'].item():.6f}, gradient rel-L2 {e_g:.2e}")
    assert abs(runs[0][0] - o64["l1"].item()) < 2e-6 and abs(runs[0][1] - o64["ssim"].item()) < 2e-6 and e_g < 1e-4


def test_model_training_branch_vs_reference_model_on_the_gpu(lib_built):
    """GaussianDiffusionB200.forward(infer=False) -> the task's mel losses -> backward, against the UNMODIFIED reference
    GaussianDiffusion.forward(infer=False) (spec_denoiser.py:168-176) + torch restatement of its losses + torch.autograd, same weights,
    same torch seed (both draw t with randint and the noise with randn_like, in that order), both in fp32 (cuDNN TF32 off for the
    reference, FSE_MODE_SIMT_F32 here).  Stated tolerance: mel_out max-abs 2e-4, losses 1e-5, denoiser gradients rel-L2 1e-3."""
    _need_gpu()
    from oracle import refshim
    if not refshim.available():
        pytest.skip("no reference tree (oracle/_ref) on this box")
    from oracle import ref_runner
    from oracle.train_oracle import mel_losses_torch
    from speech_editing_toolkit_b200 import plugin
    from speech_editing_toolkit_b200.modules import GaussianDiffusionB200
    B, T = 2, 192
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref_model, _ = ref_runner.build_models(100, device="cuda", vocoder=False)
        tb = ref_runner.synthetic_inputs(B, T, device="cuda")
        tmm = tb["time_mel_masks"][:, :, None]
        args = (tb["txt_tokens"], tmm, tb["mel2ph"], tb["spk_embed"], tb["ref_mels"], tb["f0"], tb["uv"])
        ours = GaussianDiffusionB200.from_reference(ref_model, mode="simt_f32").train()
        ref_model.eval()                                             # no dropout draws between the seed and t / noise
        torch.manual_seed(77)
        r = ref_model(*args, infer=False)
        lr = mel_losses_torch(r["mel_out"] * tmm, tb["ref_mels"] * tmm)
        ref_model.zero_grad()
        (lr["l1"] + lr["ssim"]).backward()
        task = plugin.SpeechDenoiserTaskB200.__new__(plugin.SpeechDenoiserTaskB200)
        task.hparams, task.model, task.vocoder = {"mel_losses": "l1:0.5|ssim:0.5", "use_spk_id": False}, ours, None
        sample = dict(txt_tokens=tb["txt_tokens"], mels=tb["ref_mels"], mel2ph=tb["mel2ph"], f0=tb["f0"], uv=tb["uv"],
                      time_mel_masks=tb["time_mel_masks"], spk_embed=tb["spk_embed"])
        torch.manual_seed(77)
        total, losses = task._training_step(sample, 0)
        total.backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    _, out = None, None
    e_l1, e_ss = abs(float(losses["l1_coarse"]) - float(lr["l1"])), abs(float(losses["ssim_coarse"]) - float(lr["ssim"]))
    errs = {}
    ref_grads = dict(ref_model.denoise_fn.named_parameters())
    for name, p in ours.denoise_fn.named_parameters():
        assert p.grad is not None, name
        errs[name] = rel_l2(p.grad.cpu().numpy(), ref_grads[name].grad.cpu().numpy())
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:3]
    print(f"[margin] model training branch (fp32): l1 err {e_l1:.2e}, ssim err {e_ss:.2e}, denoiser gradients median rel-L2 "
          f"{float(np.median(list(errs.values()))):.2e}, worst " + ", ".join(f"{k} {v:.1e}" for k, v in top))
    assert e_l1 < 1e-5 and e_ss < 1e-5 and top[0][1] < 1e-3, top


@pytest.mark.parametrize("mode", ["tc_bf16", "tc_tf32"])
def test_weight_gradient_gemm_vs_float64(lib_built, mode):
    """fse_wgrad (tcgen05, MN-major operands) against a float64 evaluation over the same rounded operands: the three taps of a conv with the
    zero padding at utterance ends (ragged T: partial K chunks), a strided P view (one layer's slice of a wider buffer), narrow M / N
    (80 columns: the input / output projections), and the flat [1, B*T] form of a plain GEMM.  Two calls must agree bit for bit (the frame
    slices are combined in a fixed order).  Stated tolerance: 5e-5 of the largest output element (fp32 accumulation of exact products; the measured error is printed)."""
    _need_gpu()
    from speech_editing_toolkit_b200 import train
    wg = train.WeightGradGemm(mode)
    dt = torch.bfloat16 if mode == "tc_bf16" else torch.float32
    gen = torch.Generator().manual_seed(3)

    def operand(*shape):
        x = torch.randn(*shape, generator=gen)
        if mode == "tc_tf32":                                         # what the tensor core keeps of an fp32 container: 10 mantissa bits
            x = (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
        return x.to(dt).cuda()

    def ref(P, Q, offs):
        P64, Q64 = P.double(), Q.double()
        B, T = P.shape[:2]
        out = torch.zeros(P.shape[2], Q.shape[2], len(offs), dtype=torch.float64, device=P.device)
        for j, off in enumerate(offs):
            lo, hi = max(0, -off), min(T, T - off)
            if hi > lo:
                out[:, :, j] = torch.einsum("btm,btn->mn", P64[:, lo:hi], Q64[:, lo + off:hi + off])
        return out

    cases = []
    wide = operand(3, 203, 5 * 512)
    cases.append(("conv taps, strided P", wide[:, :, 2 * 512:3 * 512], operand(3, 203, 256), (-1, 0, 1)))
    cases.append(("dilated taps", operand(2, 130, 256), operand(2, 130, 192), (-4, 0, 4)))
    cases.append(("narrow M", operand(2, 96, 80), operand(2, 96, 256), (0,)))
    cases.append(("narrow N", operand(1, 77, 256), operand(1, 77, 80), (0,)))
    cases.append(("flat rows", operand(1, 4099, 512), operand(1, 4099, 256), (0,)))
    worst = 0.0
    for name, P, Q, offs in cases:
        outs = []
        for _ in range(2):
            out = torch.full((P.shape[2], Q.shape[2], len(offs)), float("nan"), device="cuda")
            wg(P, Q, out, offs)
            outs.append(out)
        assert torch.equal(outs[0], outs[1]), name
        want = ref(P, Q, offs)
        err = float((outs[0].double() - want).abs().max() / want.abs().max())
        worst = max(worst, err)
        assert err < 5e-5, (name, err)
    # torch-layout output with strides: [M, N, taps] viewed from a [M, N * taps] buffer and a transposed 2-D output
    P, Q = cases[0][1], cases[0][2]
    buf = torch.zeros(512, 256 * 3 + 5, device="cuda")
    wg(P, Q, buf[:, :768].view(512, 256, 3), (-1, 0, 1))
    assert float((buf[:, :768].view(512, 256, 3).double() - ref(P, Q, (-1, 0, 1))).abs().max()) < 2e-5 * float(ref(P, Q, (-1, 0, 1)).abs().max()) and float(buf[:, 768:].abs().max()) == 0.0
    # four GEMMs over one (B, T) grid in one launch (the weight gradients of a residual layer) = the same four launched alone, bit for bit
    Bq, Tq = 4, 300
    dy, hin, cnd, dres, dSk, u = operand(Bq, Tq, 512), operand(Bq, Tq, 256), operand(Bq, Tq, 192), operand(Bq, Tq, 256), operand(Bq, Tq, 256), operand(Bq, Tq, 256)
    probs = [(dy, hin, (-1, 0, 1), (512, 256, 3)), (dy, cnd, (0,), (512, 192)), (dres, u, (0,), (256, 256)), (dSk, u, (0,), (256, 256))]
    alone = [wg(P, Q, torch.empty(*shape, device="cuda"), offs) for P, Q, offs, shape in probs]
    together = wg.group([(P, Q, torch.empty(*shape, device="cuda"), offs) for P, Q, offs, shape in probs])
    for a, b, (P, Q, offs, shape) in zip(alone, together, probs):
        want = ref(P, Q, offs).reshape(shape)
        assert float((b.double() - want).abs().max() / want.abs().max()) < 5e-5
        assert float((a.double() - want).abs().max() / want.abs().max()) < 5e-5
    print(f"[margin] fse_wgrad {mode}: worst error {worst:.2e} of the largest element over {len(cases)} shapes, bit-identical repeats, grouped launch ok")


def test_training_kernels_edge_shapes(lib_built):
    """Edge cases of the two training kernels: images shorter than the SSIM window, a single utterance, no speech frame at all (the
    reference divides 0 by 0 there: NaN, reproduced), one-frame and sub-chunk GEMMs, 16-column operands."""
    _need_gpu()
    from oracle.train_oracle import mel_losses_torch
    from speech_editing_toolkit_b200 import train
    gen = torch.Generator().manual_seed(9)
    for B, T in ((1, 1), (1, 5), (2, 11), (3, 17)):
        tgt = (torch.randn(B, T, 80, generator=gen) * 1.5 - 3).cuda()
        out = (tgt + 0.3 * torch.randn(B, T, 80, generator=gen).cuda()).requires_grad_(True)
        got = train.mel_losses(out, tgt)
        (got["l1"] + got["ssim"]).backward()
        o64 = out.detach().double().requires_grad_(True)
        want = mel_losses_torch(o64, tgt.double())
        (want["l1"] + want["ssim"]).backward()
        assert abs(got["l1"].item() - want["l1"].item()) < 2e-6 and abs(got["ssim"].item() - want["ssim"].item()) < 2e-6, (B, T)
        assert rel_l2(out.grad.cpu().numpy(), o64.grad.cpu().numpy()) < 1e-4, (B, T)
    z = torch.zeros(1, 20, 80, device="cuda")
    nan = train.mel_losses(z.clone().requires_grad_(True), z)
    assert torch.isnan(nan["l1"]) and torch.isnan(nan["ssim"])                 # weights.sum() == 0: 0 / 0, as the reference
    for mode in ("tc_bf16", "tc_tf32"):
        wg = train.WeightGradGemm(mode)
        dt = torch.bfloat16 if mode == "tc_bf16" else torch.float32
        for B, T, M, N, offs in ((1, 1, 128, 64, (0,)), (2, 3, 16, 16, (-1, 0, 1)), (1, 40, 256, 512, (0,)), (5, 9, 64, 32, (-2, 2))):
            P = torch.randn(B, T, M, generator=gen).to(dt).cuda()
            Q = torch.randn(B, T, N, generator=gen).to(dt).cuda()
            if mode == "tc_tf32":
                P, Q = [(x.view(torch.int32) & ~0x1FFF).view(torch.float32) for x in (P, Q)]
            got = wg(P, Q, torch.full((M, N, len(offs)), float("nan"), device="cuda"), offs)
            want = torch.zeros(M, N, len(offs), dtype=torch.float64, device="cuda")
            for j, off in enumerate(offs):
                lo, hi = max(0, -off), min(T, T - off)
                if hi > lo:
                    want[:, :, j] = torch.einsum("btm,btn->mn", P.double()[:, lo:hi], Q.double()[:, lo + off:hi + off])
            assert float((got.double() - want).abs().max()) <= 5e-5 * max(float(want.abs().max()), 1.0), (mode, B, T, M, N, offs)
