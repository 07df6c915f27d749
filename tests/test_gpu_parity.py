"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI (ctypes);
the oracle is only the checker.

Stated tolerances
  * FSE_MODE_SIMT_F32  (fp32 CUDA-core arithmetic = the reference's fp32 contract): max |err| <= 2e-4 on O(1) mels.
  * FSE_MODE_TC_BF16   (tcgen05, bf16 operands / fp32 accumulate / tanh.approx gate): relative mel-L1
    <= 2e-2 against the fp32 reference, and <= 8e-3 against the oracle's bf16-operand restatement.
  * FSE_MODE_TC_TF32   (tcgen05 kind::tf32: fp32 operands, 10-bit-mantissa multiplies, fp32 accumulate = the arithmetic of the
    reference's own GPU path, cuDNN TF32 convolutions): relative mel-L1 <= 3e-3 against the fp32 (CPU) reference.
  * integer / index work and the t == 0 posterior: bit exact.
"""
import numpy as np
import pytest

from conftest import golden, rel_l1

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL_TC_VS_F32 = 2e-2
TOL_TC_VS_BF16_ORACLE = 8e-3
TOL_SIMT_BF16_VS_ORACLE = 8e-3
TOL_F32_ABS = 2e-4
TOL_TF32_VS_F32 = 3e-3


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


_cache = {}


def make_denoiser(mode, S=None, seed=1234):
    from speech_editing_toolkit_b200 import schedule, synth
    from speech_editing_toolkit_b200.engine import Denoiser
    key = (mode, S, seed)
    if key not in _cache:
        d = Denoiser(mode=mode)
        d.load_state_dict(synth.denoiser_state_dict(seed))
        if S is not None:
            b = schedule.diffusion_buffers(S)
            d.set_schedule(b["posterior_mean_coef1"], b["posterior_mean_coef2"], b["posterior_log_variance_clipped"])
        _cache[key] = d
    return _cache[key]


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(_dev())


# ------------------------------------------------------------------ one DiffNet step
@pytest.mark.parametrize("mode", ["simt_f32", "simt_bf16", "tc_bf16", "tc_tf32"])
def test_denoise_step_vs_reference_fixture(lib_built, mode):
    from oracle import fluentspeech_oracle as O
    from speech_editing_toolkit_b200 import synth
    g = golden("diffnet_step.npz")
    B, T = int(g["B"]), int(g["T"])
    cond = synth.synthetic_cond(int(g["seed"]), B, T)                 # [B,T,H] physical layout
    d = make_denoiser(mode)
    x0 = d.denoise_step(cu(g["x"]), cu(cond), cu(g["t"])).cpu().numpy()
    assert np.isfinite(x0).all()
    if mode == "simt_f32":
        assert np.abs(x0 - g["x0"]).max() < TOL_F32_ABS
    elif mode == "tc_tf32":
        print(f"[margin] DiffNet step tc_tf32: rel-L1 {rel_l1(x0, g['x0']):.3e}, max-abs {np.abs(x0 - g['x0']).max():.3e}")
        assert rel_l1(x0, g["x0"]) < TOL_TF32_VS_F32
    else:
        ob = O.diffnet_forward(synth.denoiser_state_dict(1234), g["x"], g["t"], cond.transpose(0, 2, 1), gemm_dtype="bf16")
        assert rel_l1(x0, ob) < (TOL_SIMT_BF16_VS_ORACLE if mode == "simt_bf16" else TOL_TC_VS_BF16_ORACLE)
        assert rel_l1(x0, g["x0"]) < TOL_TC_VS_F32


@pytest.mark.parametrize("B,T", [(1, 1), (1, 127), (2, 129), (3, 300)])
def test_tc_matches_simt_bf16_on_ragged_shapes(lib_built, B, T):
    """Tile edges, batch boundaries and TMA out-of-bounds zero fill: tensor-core path vs the CUDA-core
    path over the SAME bf16 operands (differences: accumulation order + tanh.approx)."""
    from speech_editing_toolkit_b200 import synth
    rs = np.random.RandomState(B * 1000 + T)
    x = rs.standard_normal((B, 80, T)).astype(np.float32)
    cond = synth.synthetic_cond(T, B, T)
    t = rs.randint(0, 100, size=(B,)).astype(np.int64)
    a = make_denoiser("tc_bf16").denoise_step(cu(x), cu(cond), cu(t)).cpu().numpy()
    b = make_denoiser("simt_bf16").denoise_step(cu(x), cu(cond), cu(t)).cpu().numpy()
    assert np.isfinite(a).all()
    assert rel_l1(a, b) < TOL_TC_VS_BF16_ORACLE
    # the tf32 kind over fp32 operands (streamed pair kernel incl. a dummy tile when the tile count is odd) vs exact fp32 CUDA cores
    c = make_denoiser("tc_tf32").denoise_step(cu(x), cu(cond), cu(t)).cpu().numpy()
    e = make_denoiser("simt_f32").denoise_step(cu(x), cu(cond), cu(t)).cpu().numpy()
    assert np.isfinite(c).all()
    assert rel_l1(c, e) < TOL_TF32_VS_F32


def test_dilation_cycle_edges_match_oracle(lib_built):
    """dilation_cycle_length > 1 (dilation 1,2,4,...): the fp32 timestep bias needs the zero-padding edge
    correction for t < dil and t >= T - dil."""
    from oracle import fluentspeech_oracle as O
    from speech_editing_toolkit_b200 import synth
    from speech_editing_toolkit_b200.engine import Denoiser
    sd = synth.denoiser_state_dict(99, layers=4)
    B, T = 2, 40
    rs = np.random.RandomState(11)
    x = rs.standard_normal((B, 80, T)).astype(np.float32)
    cond = synth.synthetic_cond(5, B, T)
    t = np.array([3, 50], dtype=np.int64)
    ref = O.diffnet_forward(sd, x, t, cond.transpose(0, 2, 1), dilation_cycle_length=3)
    for mode, tol in (("simt_f32", 1e-3), ("tc_bf16", TOL_TC_VS_F32), ("tc_tf32", TOL_TF32_VS_F32)):
        d = Denoiser(layers=4, dilation_cycle_length=3, mode=mode)
        d.load_state_dict(sd)
        out = d.denoise_step(cu(x), cu(cond), cu(t)).cpu().numpy()
        assert rel_l1(out, ref) < tol, mode


# ------------------------------------------------------------------ posterior
def test_posterior_step_matches_oracle_and_t0_is_exact(lib_built):
    from oracle import fluentspeech_oracle as O
    S = 100
    d = make_denoiser("simt_f32", S)
    rs = np.random.RandomState(2)
    B, T = 3, 50
    x0, xt, z = (rs.standard_normal((B, 80, T)).astype(np.float32) for _ in range(3))
    t = np.array([0, 57, 99], dtype=np.int64)
    out = d.posterior_step(cu(x0), cu(xt), cu(t), cu(z)).cpu().numpy()
    ref = O.posterior_sample(O.make_schedule(S), x0, xt, t, z)
    assert np.array_equal(out[0], x0[0])            # coef1[0] = 1, coef2[0] = 0, no noise at t = 0
    assert np.allclose(out, ref, rtol=2e-6, atol=1e-6)


def test_philox_noise_is_standard_normal_and_seeded(lib_built):
    S = 100
    d = make_denoiser("simt_f32", S)
    B, T = 4, 512
    zeros = torch.zeros(B, 80, T, device=_dev())
    t = torch.full((B,), 99, dtype=torch.long, device=_dev())
    from speech_editing_toolkit_b200 import schedule
    sigma = float(np.exp(0.5 * schedule.diffusion_buffers(S)["posterior_log_variance_clipped"][99]))
    a = d.posterior_step(zeros, zeros, t, None, seed=7, step=3).cpu().numpy() / sigma
    b = d.posterior_step(zeros, zeros, t, None, seed=7, step=3).cpu().numpy() / sigma
    c = d.posterior_step(zeros, zeros, t, None, seed=8, step=3).cpu().numpy() / sigma
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(a.mean()) < 0.01 and abs(a.std() - 1.0) < 0.01
    assert abs((a ** 3).mean()) < 0.03 and abs((a ** 4).mean() - 3.0) < 0.1
    assert abs(np.corrcoef(a[:, :, :-1].ravel(), a[:, :, 1:].ravel())[0, 1]) < 0.01


# ------------------------------------------------------------------ sampling loop, config C1
@pytest.mark.parametrize("mode", ["simt_f32", "simt_bf16", "tc_bf16", "tc_tf32"])
def test_sample_c1_vs_reference_fixture(lib_built, mode):
    from speech_editing_toolkit_b200 import synth
    g = golden("sample_c1.npz")
    seed, B, T, S = int(g["seed"]), int(g["B"]), int(g["T"]), int(g["S"])
    d = make_denoiser(mode, S)
    cond = synth.synthetic_cond(seed, B, T)
    noise = synth.synthetic_noise(seed, S, B, T)
    mel, xs = d.sample(cu(cond), cu(noise), trace=True)
    mel, xs = mel.cpu().numpy(), xs.cpu().numpy()
    assert mel.shape == (B, T, 80) and np.isfinite(mel).all()
    assert np.array_equal(mel, xs[-1].transpose(0, 2, 1))          # mel_out is x[:,0].transpose(1,2)
    if mode == "simt_f32":
        assert np.abs(xs[0] - g["x_after_first"]).max() < TOL_F32_ABS
        assert np.abs(mel - g["mel_out"]).max() < 5e-4
    else:
        tol = TOL_TF32_VS_F32 if mode == "tc_tf32" else TOL_TC_VS_F32
        print(f"[margin] C1 loop {mode}: mel rel-L1 {rel_l1(mel, g['mel_out']):.3e}")
        assert rel_l1(xs[0], g["x_after_first"]) < tol
        assert rel_l1(mel, g["mel_out"]) < tol
    assert d.last_launches > 0


def test_sample_composite_is_bit_exact_outside_mask(lib_built):
    from speech_editing_toolkit_b200 import synth
    S, B, T = 4, 2, 96
    d = make_denoiser("tc_bf16", S)
    batch = synth.synthetic_edit_batch(3, B, T)
    cond = synth.synthetic_cond(3, B, T)
    noise = synth.synthetic_noise(3, S, B, T)
    plain = d.sample(cu(cond), cu(noise)).cpu().numpy()
    comp = d.sample(cu(cond), cu(noise), ref_mel=cu(batch["ref_mels"]), mask=cu(batch["time_mel_masks"])).cpu().numpy()
    m = batch["time_mel_masks"].astype(bool)
    assert m.any() and (~m).any()
    assert np.array_equal(comp[~m], batch["ref_mels"][~m])        # integer frame masking: exact
    assert np.array_equal(comp[m], plain[m])


def test_sample_host_roundtrip_equals_device_call(lib_built):
    from speech_editing_toolkit_b200 import synth
    S, B, T = 4, 2, 64
    d = make_denoiser("tc_bf16", S)
    cond = synth.synthetic_cond(9, B, T)
    noise = synth.synthetic_noise(9, S, B, T)
    a = d.sample(cu(cond), cu(noise)).cpu().numpy()
    b = d.sample_host(cond, noise)
    assert np.array_equal(a, b)
    c = d.sample_host(cond, None, seed=5)
    e = d.sample_host(cond, None, seed=5)
    assert np.isfinite(c).all() and np.array_equal(c, e)


def test_full_size_batch_items_are_independent(lib_built):
    """BASELINE config 2 shape (B=32, T=1024): utterances never interact, so item b of the batched run
    equals a B=1 run on that item (checks tile/batch indexing at full size)."""
    from speech_editing_toolkit_b200 import synth
    S, B, T = 2, 32, 1024
    d = make_denoiser("tc_bf16", S)
    cond = synth.synthetic_cond(21, B, T)
    noise = synth.synthetic_noise(21, S, B, T)
    full = d.sample(cu(cond), cu(noise)).cpu().numpy()
    assert np.isfinite(full).all()
    for b in (0, 17, 31):
        one = d.sample(cu(cond[b:b + 1]), cu(noise[:, b:b + 1])).cpu().numpy()
        assert np.array_equal(one[0], full[b])


def test_errors_are_loud(lib_built):
    from speech_editing_toolkit_b200 import FseError
    from speech_editing_toolkit_b200.engine import Denoiser
    d = Denoiser(mode="tc_bf16")
    with pytest.raises(FseError):                       # weights not loaded
        d.denoise_step(torch.zeros(1, 80, 8, device=_dev()), torch.zeros(1, 8, 192, device=_dev()),
                       torch.zeros(1, dtype=torch.long, device=_dev()))
    with pytest.raises(FseError):                       # missing tensors
        d.load_state_dict({"input_projection.weight": np.zeros((256, 80, 1), np.float32)})
    with pytest.raises(FseError):                       # CPU tensors are refused, no fallback
        make_denoiser("tc_bf16").denoise_step(torch.zeros(1, 80, 8), torch.zeros(1, 8, 192), torch.zeros(1, dtype=torch.long))


# ------------------------------------------------------------------ HiFi-GAN
@pytest.mark.parametrize("mode", ["simt_f32", "simt_bf16", "tc_bf16", "tc_tf32"])
def test_hifigan_vs_reference_fixture(lib_built, mode):
    from oracle import fluentspeech_oracle as O
    from speech_editing_toolkit_b200 import synth
    from speech_editing_toolkit_b200.engine import Vocoder
    g = golden("hifigan_v1.npz")
    v = Vocoder(mode=mode)
    sd = synth.hifigan_state_dict(int(g["seed"]))
    v.load_state_dict(sd)
    wav = v.forward(cu(g["mel"])).cpu().numpy()
    assert wav.shape == g["wav"].shape and np.isfinite(wav).all()
    if mode == "simt_f32":
        assert np.abs(wav - g["wav"]).max() < TOL_F32_ABS
    elif mode == "tc_tf32":
        print(f"[margin] HiFi-GAN T=24 tc_tf32: wav rel-L1 {rel_l1(wav, g['wav']):.3e}")
        assert rel_l1(wav, g["wav"]) < 5e-3
    else:
        ob = O.hifigan_forward(sd, O.HIFIGAN_V1, g["mel"].transpose(0, 2, 1), gemm_dtype="bf16")[:, 0]
        assert rel_l1(wav, ob) < 1e-2
        assert rel_l1(wav, g["wav"]) < 3e-2
    assert np.array_equal(v.forward_host(g["mel"]), wav)


def test_hifigan_tc_matches_simt_on_batch(lib_built):
    from speech_editing_toolkit_b200 import synth
    from speech_editing_toolkit_b200.engine import Vocoder
    sd = synth.hifigan_state_dict(5)
    rs = np.random.RandomState(4)
    mel = np.clip(rs.standard_normal((2, 37, 80)) * 1.5 - 3, -6, 1.5).astype(np.float32)
    outs = {}
    for mode in ("tc_bf16", "simt_bf16"):
        v = Vocoder(mode=mode)
        v.load_state_dict(sd)
        outs[mode] = v.forward(cu(mel)).cpu().numpy()
    assert outs["tc_bf16"].shape == (2, 37 * 256)
    assert rel_l1(outs["tc_bf16"], outs["simt_bf16"]) < 5e-3


@pytest.mark.gpu
@pytest.mark.parametrize("kernels", [[3], [3, 5]])
def test_hifigan_other_resblock_counts(lib_built, kernels):
    """Generators with one or two parallel resblocks per stage (hifigan.py:131-137: xs = mean over the blocks): the
    running-sum epilogue variants (first / middle / last block) must also hold when there is no middle or no sum at all."""
    from oracle import fluentspeech_oracle as O
    from speech_editing_toolkit_b200 import synth
    from speech_editing_toolkit_b200.engine import Vocoder
    cfg = dict(upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], upsample_initial_channel=128, resblock="1",
               resblock_kernel_sizes=kernels, resblock_dilation_sizes=[[1, 3, 5]] * len(kernels))
    sd = synth.hifigan_state_dict(11, cfg)
    rs = np.random.RandomState(12)
    mel = np.clip(rs.standard_normal((2, 150, 80)) * 1.5 - 3, -6, 1.5).astype(np.float32)
    ref = O.hifigan_forward(sd, cfg, mel.transpose(0, 2, 1), gemm_dtype="f32")[:, 0]
    v32 = Vocoder(cfg, mode="simt_f32"); v32.load_state_dict(sd)
    assert np.abs(v32.forward(cu(mel)).cpu().numpy() - ref).max() < TOL_F32_ABS
    vtc = Vocoder(cfg, mode="tc_bf16"); vtc.load_state_dict(sd)
    wav = vtc.forward(cu(mel)).cpu().numpy()
    assert wav.shape == (2, 150 * 8) and np.isfinite(wav).all()
    assert rel_l1(wav, ref) < 3e-2


@pytest.mark.gpu
def test_hifigan_resblock2_generator(lib_built):
    """config['resblock'] == '2' (hifigan.py:67-88: two dilated convs per block, each `x = c(lrelu(x)) + x`; HiFi-GAN V3-style
    config) against the oracle, which tests/test_oracles_vs_live_reference.py pins to the live reference generator."""
    from oracle import fluentspeech_oracle as O
    from speech_editing_toolkit_b200 import synth
    from speech_editing_toolkit_b200.engine import Vocoder
    cfg = dict(upsample_rates=[8, 8, 4], upsample_kernel_sizes=[16, 16, 8], upsample_initial_channel=256, resblock="2",
               resblock_kernel_sizes=[3, 5, 7], resblock_dilation_sizes=[[1, 2], [2, 6], [3, 12]])
    sd = synth.hifigan_state_dict(17, cfg)
    assert any(".convs.1." in k for k in sd) and not any("convs1" in k for k in sd)
    mel = np.clip(np.random.RandomState(18).standard_normal((2, 45, 80)) * 1.5 - 3.0, -6, 1.5).astype(np.float32)
    ref = O.hifigan_forward(sd, cfg, mel.transpose(0, 2, 1))[:, 0]
    for mode, tol in (("simt_f32", None), ("tc_bf16", 3e-2), ("tc_tf32", 5e-3)):
        v = Vocoder(cfg, mode=mode)
        v.load_state_dict(sd)
        wav = v.forward(cu(mel)).cpu().numpy()
        assert wav.shape == ref.shape == (2, 45 * 256) and np.isfinite(wav).all()
        if tol is None:
            assert np.abs(wav - ref).max() < TOL_F32_ABS
        else:
            assert rel_l1(wav, ref) < tol, mode


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["simt_f32", "tc_bf16", "tc_tf32"])
def test_mel_encoder_vs_reference_fixture(lib_built, mode):
    """fse_mel_encoder_forward against the unmodified reference MelEncoder (tests/golden/mel_encoder.npz), plain and with
    the call site's `decoder_inp + out * tgt_nonpadding` fused into the last epilogue (spec_denoiser.py:162-164)."""
    from speech_editing_toolkit_b200 import synth
    from speech_editing_toolkit_b200.engine import MelEncoderKernel
    g = golden("mel_encoder.npz")
    B, T = int(g["B"]), int(g["T"])
    enc = MelEncoderKernel(80, 192, mode=mode)
    enc.load_state_dict(synth.mel_encoder_state_dict(int(g["seed"])))
    ref, mask = synth.synthetic_ref_and_mask(int(g["seed"]), B, T)
    x = cu(ref * (1 - mask))
    out = enc.forward(x).cpu().numpy()
    cond = enc.forward(x, cu(g["decoder_inp"]), cu(g["nonpad"])).cpu().numpy()
    assert enc.last_launches == (4 if mode == "tc_bf16" else 3)
    if mode == "simt_f32":
        assert np.abs(out - g["out"]).max() < TOL_F32_ABS
        assert np.abs(cond - g["cond"]).max() < TOL_F32_ABS
    else:
        tol = 2e-3 if mode == "tc_tf32" else 1e-2
        assert rel_l1(out, g["out"]) < tol
        assert rel_l1(cond, g["cond"]) < tol
    assert np.array_equal(cond[1, 60:], g["decoder_inp"][1, 60:])      # padding frames: decoder_inp bit for bit


@pytest.mark.gpu
def test_mel_encoder_module_matches_oracle_on_batch(lib_built):
    """The nn.Module drop-in (same state_dict keys as the reference MelEncoder) on a ragged multi-tile batch."""
    from oracle import fluentspeech_oracle as O
    from speech_editing_toolkit_b200 import synth
    from speech_editing_toolkit_b200.modules import MelEncoderB200
    sd = synth.mel_encoder_state_dict(5)
    m = MelEncoderB200(80, 192).cuda()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    assert sorted(m.state_dict()) == sorted(sd)
    ref, mask = synth.synthetic_ref_and_mask(6, 3, 333)
    x = ref * (1 - mask)
    out = m(cu(x)).cpu().numpy()
    assert out.shape == (3, 333, 192)
    assert rel_l1(out, O.mel_encoder_forward(sd, x)) < 1e-2
