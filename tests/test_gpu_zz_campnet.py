"""Parity of the native CampNet mask-predict forward (fse_campnet_*; CampNetB200) on the GPU against
  * tests/golden/campnet.npz — outputs of the unmodified reference CampNet.forward (oracle/make_golden.py campnet), and
  * oracle/campnet_oracle.py on a larger ragged batch (several query / key tiles in the attention kernel).
Stated tolerances: FSE_MODE_SIMT_F32 max-abs <= 2e-3 on the mel outputs (values up to ~4) and 1e-3 on the attention
probabilities; FSE_MODE_TC_BF16 relative L1 over the masked (predicted) region <= 3e-2 vs the fp32 reference and <= 1.5e-2 vs
the bf16-operand oracle; the unmasked region is the input mel bit for bit; padded frames are exactly zero."""
import numpy as np
import pytest

from conftest import golden, rel_l1

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

HP = dict(hidden_size=192, dec_ffn_kernel_size=9, audio_num_mel_bins=80)


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _module(sd, vocab, mode):
    from speech_editing_toolkit_b200.modules import CampNetB200
    net = CampNetB200(vocab, 100, dict(HP, b200_mode=mode)).cuda()
    missing, unexpected = net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not unexpected
    return net


@pytest.mark.parametrize("mode", ["simt_f32", "tc_bf16"])
def test_campnet_forward_vs_reference_fixture(lib_built, mode):
    _need_gpu()
    from speech_editing_toolkit_b200 import synth
    g = golden("campnet.npz")
    seed, B, T, vocab = int(g["seed"]), int(g["B"]), int(g["T"]), int(g["vocab"])
    b = synth.synthetic_campnet_batch(seed, B, T, vocab=vocab, pad_items=[(1, 4)])
    net = _module(synth.campnet_state_dict(seed, vocab), vocab, mode)
    ret = net(cu(b["txt_tokens"]), mels=cu(b["mels"]), time_mel_masks=cu(b["time_mel_masks"]), infer=True)
    assert set(ret) == {"mel_out_coarse", "mel_out_fine", "attn"}
    out = {k: v.cpu().numpy() for k, v in ret.items()}
    m = b["time_mel_masks"]
    assert np.isfinite(out["mel_out_fine"]).all() and np.isfinite(out["mel_out_coarse"]).all() and np.isfinite(out["attn"]).all()
    assert np.array_equal(out["mel_out_fine"] * (1 - m), b["mels"] * (1 - m))            # unmasked frames: the input, bit for bit
    assert np.abs(out["mel_out_coarse"][1, -32:]).max() == 0.0                            # padded frames
    assert np.abs(out["attn"].sum(-1) - 1).max() < 1e-4 and np.abs(out["attn"][1, :, -4:]).max() == 0.0
    if mode == "simt_f32":
        assert np.abs(out["mel_out_coarse"] - g["mel_out_coarse"]).max() < 2e-3
        assert np.abs(out["mel_out_fine"] - g["mel_out_fine"]).max() < 2e-3
        assert np.abs(out["attn"] - g["attn"].astype(np.float32)).max() < 1e-3
    else:
        assert rel_l1(out["mel_out_coarse"] * m, g["mel_out_coarse"] * m) < 3e-2
        assert rel_l1(out["mel_out_fine"] * m, g["mel_out_fine"] * m) < 3e-2
        assert np.abs(out["attn"] - g["attn"].astype(np.float32)).max() < 5e-2
    assert net.engine().last_launches > 100


def test_campnet_multi_tile_ragged_batch_vs_oracle(lib_built):
    """3 items x 330 frames / 82 tokens: six query tiles and two key tiles per (item, head), ragged tails, both contracts."""
    _need_gpu()
    from oracle import campnet_oracle as KO
    from speech_editing_toolkit_b200 import synth
    vocab, B, T = 60, 3, 330
    sd = synth.campnet_state_dict(7, vocab)
    b = synth.synthetic_campnet_batch(8, B, T, vocab=vocab, frames_per_phone=4, pad_items=[(1, 20), (2, 5)])
    m = b["time_mel_masks"]
    ref32 = KO.campnet_forward(sd, b["txt_tokens"], b["mels"], m)
    refbf = KO.campnet_forward(sd, b["txt_tokens"], b["mels"], m, gemm_dtype="bf16", attn_dtype="bf16")
    for mode in ("simt_f32", "tc_bf16"):
        eng = _module(sd, vocab, mode).engine()
        out = {k: v.cpu().numpy() for k, v in eng.forward(cu(b["txt_tokens"]), cu(b["mels"]), cu(m), need_encoder_out=True).items()}
        if mode == "simt_f32":
            assert np.abs(out["encoder_out"] - ref32["encoder_out"]).max() < 1e-3
            for k in ("mel_out_coarse", "mel_out_fine"):
                assert np.abs(out[k] - ref32[k]).max() < 2e-3, k
            assert np.abs(out["attn"] - ref32["attn"]).max() < 1e-3
        else:
            assert rel_l1(out["encoder_out"], ref32["encoder_out"]) < 3e-2
            for k in ("mel_out_coarse", "mel_out_fine"):
                assert rel_l1(out[k] * m, ref32[k] * m) < 3e-2, k
                assert rel_l1(out[k] * m, refbf[k] * m) < 1.5e-2, k


def test_tensor_core_attention_matches_cuda_core_attention(lib_built, monkeypatch):
    """The tcgen05 attention kernel (attention_tc.cuh: S = Q K^T and O = P V as tcgen05.mma, softmax out of TMEM) against the
    CUDA-core flash kernel on the same bf16 q / k / v, through the whole forward: 3 query tiles x 3 key tiles per (item, head)
    in the decoder, masked keys in the encoder / cross attention, ragged tails."""
    _need_gpu()
    from oracle import campnet_oracle as KO
    from speech_editing_toolkit_b200 import synth
    vocab, B, T = 60, 2, 330
    sd = synth.campnet_state_dict(17, vocab)
    b = synth.synthetic_campnet_batch(18, B, T, vocab=vocab, frames_per_phone=4, pad_items=[(1, 9)])
    m = b["time_mel_masks"]
    outs = {}
    for sel in ("simt", "tc"):
        monkeypatch.setenv("FSE_CAMP_ATTN", sel)
        eng = _module(sd, vocab, "tc_bf16").engine()
        outs[sel] = {k: v.cpu().numpy() for k, v in eng.forward(cu(b["txt_tokens"]), cu(b["mels"]), cu(m), need_encoder_out=True).items()}
        assert all(np.isfinite(v).all() for v in outs[sel].values()), sel
    ref = KO.campnet_forward(sd, b["txt_tokens"], b["mels"], m)
    # two different kernels really ran: the tensor-core path rounds P to bf16, so the results agree closely but not bit for bit
    assert not np.array_equal(outs["tc"]["encoder_out"], outs["simt"]["encoder_out"])
    for k in ("encoder_out", "mel_out_coarse", "mel_out_fine"):
        w = m if k != "encoder_out" else 1.0
        assert rel_l1(outs["tc"][k] * w, outs["simt"][k] * w) < 1.5e-2, k          # P is rounded to bf16 on the tensor-core path
        assert rel_l1(outs["tc"][k] * w, ref[k] * w) < 3e-2, k
    assert np.abs(outs["tc"]["attn"] - ref["attn"]).max() < 5e-2


def test_two_tile_attention_schedule_matches(lib_built, monkeypatch):
    """FSE_CAMP_ATTN=tc2: two query tiles per CTA, two softmax warp groups sharing the K / V^T tiles, mask-free fast path and
    ex2.approx (attention_tc.cuh, second schedule) against the CUDA-core kernel and the oracle; odd tile counts on both axes
    (330 queries = 2 CTAs of 256 with a ragged second tile; 3 key tiles; 82-key cross attention with padded keys)."""
    _need_gpu()
    from oracle import campnet_oracle as KO
    from speech_editing_toolkit_b200 import synth
    vocab, B, T = 60, 2, 330
    sd = synth.campnet_state_dict(17, vocab)
    b = synth.synthetic_campnet_batch(18, B, T, vocab=vocab, frames_per_phone=4, pad_items=[(1, 9)])
    m = b["time_mel_masks"]
    outs = {}
    for sel in ("simt", "tc2"):
        monkeypatch.setenv("FSE_CAMP_ATTN", sel)
        eng = _module(sd, vocab, "tc_bf16").engine()
        outs[sel] = {k: v.cpu().numpy() for k, v in eng.forward(cu(b["txt_tokens"]), cu(b["mels"]), cu(m), need_encoder_out=True).items()}
        assert all(np.isfinite(v).all() for v in outs[sel].values()), sel
    ref = KO.campnet_forward(sd, b["txt_tokens"], b["mels"], m)
    assert not np.array_equal(outs["tc2"]["encoder_out"], outs["simt"]["encoder_out"])
    for k in ("encoder_out", "mel_out_coarse", "mel_out_fine"):
        w = m if k != "encoder_out" else 1.0
        assert rel_l1(outs["tc2"][k] * w, outs["simt"][k] * w) < 1.5e-2, k
        assert rel_l1(outs["tc2"][k] * w, ref[k] * w) < 3e-2, k
