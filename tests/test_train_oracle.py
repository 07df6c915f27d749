"""CPU pins of the training path: the buffer-for-buffer torch restatement of the native forward / backward
(oracle/train_oracle.py) under the product's own weight-gradient code (speech_editing_toolkit_b200.train.DiffNetFunction) must
reproduce the gradients torch.autograd computes through the unmodified reference DiffNet (tests/golden/diffnet_train.npz), and the
loss helpers must equal the reference's (live reference, build container only)."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import refshim
from oracle.train_oracle import TorchTrainer, mel_losses_torch
from speech_editing_toolkit_b200 import synth, train
from speech_editing_toolkit_b200.modules import DiffNetB200

HP = dict(audio_num_mel_bins=80, hidden_size=192, residual_channels=256, dilation_cycle_length=1)


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-30))


def test_native_training_algebra_and_weight_gradient_code_vs_reference_autograd_fixture(monkeypatch):
    torch.set_num_threads(4)
    g = golden("diffnet_train.npz")
    L, B, T = int(g["layers"]), int(g["B"]), int(g["T"])
    net = DiffNetB200(80, dict(HP, residual_layers=L, b200_mode="simt_f32")).train()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.denoiser_state_dict(int(g["seed"]), layers=L).items()})
    net._trainer = TorchTrainer(80, 192, 256, L, 1, "simt_f32")          # stands in for the C-ABI handle
    monkeypatch.setattr(train, "_need_cuda", lambda *a: None)
    cond = torch.from_numpy(synth.synthetic_cond(int(g["seed"]) + 2, B, T)).requires_grad_(True)
    x0 = train.diffnet_train_forward(net, torch.from_numpy(g["x"])[:, None], torch.from_numpy(g["t"]), cond.transpose(1, 2))[:, 0]
    x0.backward(torch.from_numpy(g["dx0"]))
    assert rel_l2(x0.detach().numpy(), g["x0"]) < 1e-5
    assert rel_l2(cond.grad.numpy(), g["dcond"]) < 1e-5
    for name, p in net.named_parameters():
        got = p.grad.numpy().reshape(-1)
        got_s = got if got.size <= 4096 else got[::61]
        assert rel_l2(got_s, g["g__" + name]) < 2e-5, name
        assert abs(np.sqrt((got.astype(np.float64) ** 2).sum()) - float(g["gnorm__" + name])) < 1e-4 * float(g["gnorm__" + name]) + 1e-12


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_mel_losses_equal_the_reference_losses():
    """oracle.train_oracle.mel_losses_torch (the expected value of the native loss kernels) against SpeechBaseTask.l1_loss / ssim_loss (tasks/tts/speech_base.py:219-257) — the task module cannot be
    imported (matplotlib, librosa...), so the two methods are cut out of the source and run on the reference's own ssim()."""
    import ast
    import os
    refshim.install("egs/spec_denoiser.yaml")
    from utils.metrics.ssim import ssim
    from utils.nn.seq_utils import weights_nonzero_speech
    import torch.nn.functional as F
    src = open(os.path.join(refshim.REF_ROOT, "tasks", "tts", "speech_base.py")).read() if os.path.exists(
        os.path.join(refshim.REF_ROOT, "tasks", "tts", "speech_base.py")) else None
    if src is None:
        pytest.skip("tasks/ not in the vendored copy")
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "SpeechBaseTask")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("l1_loss", "ssim_loss")]
    ns = {"F": F, "ssim": ssim, "weights_nonzero_speech": weights_nonzero_speech, "torch": torch}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "speech_base.py", "exec"), ns)
    rs = np.random.RandomState(3)
    a = torch.from_numpy(rs.standard_normal((2, 50, 80)).astype(np.float32))
    b = torch.from_numpy(rs.standard_normal((2, 50, 80)).astype(np.float32))
    m = torch.zeros(2, 50, 1); m[:, 10:30] = 1
    ours = mel_losses_torch(a * m, b * m)
    assert abs(float(ours["l1"]) - 0.5 * float(ns["l1_loss"](None, a * m, b * m))) < 1e-6
    assert abs(float(ours["ssim"]) - 0.5 * float(ns["ssim_loss"](None, a * m, b * m))) < 1e-5


def test_mel_loss_oracles_vs_reference_fixture():
    """tests/golden/mel_loss.npz holds the reference's own l1_loss / ssim_loss and their autograd gradient (oracle/make_golden.py mel_loss).
    Both restatements must reproduce it: the torch one (expected value of the GPU test) and the closed-form algebra of csrc/mel_loss.cu."""
    from oracle.train_oracle import mel_loss_native_algebra
    g = golden("mel_loss.npz")
    a = torch.from_numpy(g["mel_out"]).requires_grad_(True)
    out = mel_losses_torch(a, torch.from_numpy(g["target"]))
    (out["l1"] + out["ssim"]).backward()
    assert abs(float(out["l1"]) - float(g["l1_f32"])) < 1e-6 and abs(float(out["ssim"]) - float(g["ssim_f32"])) < 1e-6
    assert rel_l2(a.grad.numpy(), g["grad_f32"]) < 1e-5
    l1, ss, grad = mel_loss_native_algebra(g["mel_out"], g["target"])
    # the reference's fp64 run still uses the fp32-rounded window (create_window builds a float tensor): 1e-7 is that rounding
    assert abs(l1 - float(g["l1_f64"])) < 1e-9 and abs(ss - float(g["ssim_f64"])) < 1e-7
    assert rel_l2(grad, g["grad_f64"]) < 1e-6
