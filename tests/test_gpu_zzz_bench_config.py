"""Parity AT THE BENCHMARKED SHAPES (BASELINE configs[1] / [2] / [3]): T = 1024 frames, S = 100 diffusion iterations, batch 32.

Fixtures `sample_s100_t1024.npz`, `hifigan_t1024.npz`, `campnet_t1024.npz` are outputs of the unmodified reference
(`oracle/make_golden.py bench_config`): its own `p_sample` loop with the 101 normal draws injected, `HifiGanGenerator.forward`
and `CampNet.forward` on ONE item of the full length.  No op of the path mixes batch items, so the B=32 runs
are pinned by (a) the B=1 comparison with the reference and (b) bit-exact item independence of the batched kernels, which the
tests below assert at S=100 / T=1024 / B=32.  Every assertion prints the measured margin (pytest -s / -rP shows it).

Stated tolerances: as tests/test_gpu_parity.py — fp32 CUDA-core mode max-abs (mel 5e-4 after 100 steps, wav 2e-4), tensor-core
modes relative L1 <= 2e-2 (mel) / 3e-2 (wav, CampNet) against the fp32 reference."""
import numpy as np
import pytest

from conftest import golden, rel_l1

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = {"tc_bf16": 2e-2, "tc_tf32": 4e-3}


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _denoiser(mode, S):
    from speech_editing_toolkit_b200 import schedule, synth
    from speech_editing_toolkit_b200.engine import Denoiser
    d = Denoiser(mode=mode)
    d.load_state_dict(synth.denoiser_state_dict(1234))
    b = schedule.diffusion_buffers(S)
    d.set_schedule(b["posterior_mean_coef1"], b["posterior_mean_coef2"], b["posterior_log_variance_clipped"])
    return d


@pytest.mark.parametrize("mode", ["simt_f32", "tc_bf16", "tc_tf32"])
def test_sample_s100_t1024_vs_reference_fixture(lib_built, mode):
    """spec_denoiser.py:177-185 at the benchmark's T and S: one item, the reference's own loop as the expected value."""
    _need_gpu()
    from speech_editing_toolkit_b200 import synth
    g = golden("sample_s100_t1024.npz")
    seed, B, T, S = int(g["seed"]), int(g["B"]), int(g["T"]), int(g["S"])
    assert (B, T, S) == (1, 1024, 100)
    d = _denoiser(mode, S)
    cond, noise = synth.synthetic_cond(seed, B, T), synth.synthetic_noise(seed, S, B, T)
    mel, xs = d.sample(cu(cond), cu(noise), trace=True)
    mel, xs = mel.cpu().numpy(), xs.cpu().numpy()
    assert np.isfinite(mel).all()
    mid = {t: rel_l1(xs[S - 1 - t][:, :, :64], g[f"x_t{t}"]) for t in (75, 50, 25)}       # xs[k] = x after iteration k (t = S-1-k)
    err_abs, err_rel = float(np.abs(mel - g["mel_out"]).max()), rel_l1(mel, g["mel_out"])
    print(f"[margin] S=100 T=1024 {mode}: mel max-abs {err_abs:.3e}, rel-L1 {err_rel:.3e}; trace rel-L1 at t=75/50/25 "
          f"{mid[75]:.2e}/{mid[50]:.2e}/{mid[25]:.2e}")
    if mode == "simt_f32":
        assert err_abs < 5e-4
    else:
        assert err_rel < TOL[mode]
        assert max(mid.values()) < TOL[mode]          # the error does not build up along the chain


@pytest.mark.parametrize("mode", ["tc_bf16", "tc_tf32"])
def test_b32_s100_items_match_the_pinned_single_item_run(lib_built, mode):
    """BASELINE configs[1] in full (B=32, T=1024, S=100): the pinned item sits in slot 5 of a batch of 32 different utterances;
    its output must equal the B=1 run bit for bit (tile / batch indexing, the streamed kernel's item dealing and flags)."""
    _need_gpu()
    from speech_editing_toolkit_b200 import synth
    g = golden("sample_s100_t1024.npz")
    seed, T, S, B, slot = int(g["seed"]), 1024, 100, 32, 5
    d = _denoiser(mode, S)
    cond1, noise1 = synth.synthetic_cond(seed, 1, T), synth.synthetic_noise(seed, S, 1, T)
    one = d.sample(cu(cond1), cu(noise1)).cpu().numpy()
    gen = torch.Generator(device="cuda").manual_seed(7)
    cond = torch.randn(B, T, 192, device="cuda", generator=gen) * 0.5
    noise = torch.randn(S + 1, B, 80, T, device="cuda", generator=gen)
    cond[slot] = cu(cond1)[0]
    noise[:, slot] = cu(noise1)[:, 0]
    full = d.sample(cond, noise).cpu().numpy()
    assert np.isfinite(full).all()
    assert np.array_equal(full[slot], one[0])
    err = rel_l1(full[slot], g["mel_out"][0])
    print(f"[margin] B=32 S=100 T=1024 {mode}: item {slot} rel-L1 vs reference {err:.3e}")
    assert err < TOL[mode]


def test_stream_kernel_is_deterministic_at_bench_shape(lib_built):
    """20 runs of fse_sample at 32 x 1024 (S=2) must be bit-identical: the cheapest detector of a missed acquire in the streamed
    kernel's done[l][u] protocol (a stale halo row would change the result of some run)."""
    _need_gpu()
    from speech_editing_toolkit_b200 import synth
    S, B, T = 2, 32, 1024
    d = _denoiser("tc_bf16", S)
    cond, noise = cu(synth.synthetic_cond(31, B, T)), cu(synth.synthetic_noise(31, S, B, T))
    first = d.sample(cond, noise).clone()
    for i in range(20):
        again = d.sample(cond, noise)
        assert torch.equal(first, again), f"run {i} differs"


@pytest.mark.parametrize("mode", ["tc_bf16", "tc_tf32"])
def test_stream_dependent_launch_equals_plain_launch(lib_built, mode, monkeypatch):
    """The streamed kernel as a programmatic dependent launch with alternating publication counters (the default) against the
    memset + plain launch it replaced (FSE_STREAM_PDL=0), eagerly and as a replayed graph, also with the batch walked in chunks
    (FSE_BATCH_CHUNK: several streamed launches per evaluation, odd tile counts): schedules differ, bits must not."""
    _need_gpu()
    from speech_editing_toolkit_b200 import synth
    S, B, T = 6, 5, 700                                   # 6 tiles per item, 30 tiles: 15 CTA pairs' worth of units, ragged last tile
    cond, noise = cu(synth.synthetic_cond(77, B, T)), cu(synth.synthetic_noise(77, S, B, T))
    outs = {}
    for pdl, chunk in (("0", "0"), ("1", "0"), ("1", "2"), ("0", "2")):
        monkeypatch.setenv("FSE_STREAM_PDL", pdl)
        monkeypatch.setenv("FSE_BATCH_CHUNK", chunk)
        d = _denoiser(mode, S)
        runs = [d.sample(cond, noise).clone() for _ in range(4)]     # eager, capture, two replays
        for i, r in enumerate(runs[1:]):
            assert torch.equal(runs[0], r), f"pdl={pdl} chunk={chunk}: call {i + 1} differs from the eager call"
        outs[(pdl, chunk)] = runs[0]
        del d
    ref = outs[("0", "0")]
    for k, v in outs.items():
        assert torch.equal(ref, v), f"FSE_STREAM_PDL={k[0]} FSE_BATCH_CHUNK={k[1]} changes the result"
    print(f"[margin] {mode}: dependent launch / plain launch / chunked batch: bit-identical over {len(outs)} variants x 4 calls")


@pytest.mark.parametrize("mode", ["simt_f32", "tc_bf16", "tc_tf32"])
def test_hifigan_t1024_vs_reference_fixture(lib_built, mode):
    """hifigan.py:126-142 at T=1024 (stage-4 tensors of 262 144 rows: the multi-sub-tile / many-wave regime)."""
    _need_gpu()
    from speech_editing_toolkit_b200 import synth
    from speech_editing_toolkit_b200.engine import Vocoder
    g = golden("hifigan_t1024.npz")
    T = int(g["T"])
    mel = np.clip(np.random.RandomState(int(g["seed"])).standard_normal((1, T, 80)) * 1.5 - 3.0, -6, 1.5).astype(np.float32)
    v = Vocoder(mode=mode)
    v.load_state_dict(synth.hifigan_state_dict(1234))
    wav = v.forward(cu(mel)).cpu().numpy()
    assert wav.shape == g["wav"].shape and np.isfinite(wav).all()
    err_abs, err_rel = float(np.abs(wav - g["wav"]).max()), rel_l1(wav, g["wav"])
    print(f"[margin] HiFi-GAN T=1024 {mode}: wav max-abs {err_abs:.3e}, rel-L1 {err_rel:.3e}")
    if mode == "simt_f32":
        assert err_abs < 2e-4
    else:
        assert err_rel < (3e-2 if mode == "tc_bf16" else 5e-3)
    if mode != "simt_f32":
        # the same item inside the benchmarked batch of 32 (stage-4 tensors of 8.4 M rows, many waves at every stage): bit-identical
        rs = np.random.RandomState(3)
        batch = np.clip(rs.standard_normal((32, T, 80)) * 1.5 - 3.0, -6, 1.5).astype(np.float32)
        batch[21] = mel[0]
        wb = v.forward(cu(batch))
        assert bool(torch.isfinite(wb).all())
        assert np.array_equal(wb[21].cpu().numpy(), wav[0])


@pytest.mark.parametrize("mode", ["simt_f32", "tc_bf16"])
def test_campnet_t1024_vs_reference_fixture(lib_built, mode):
    """campnet.py:40-69 on one item of 1024 frames / 128 tokens (8 query tiles x 8 key tiles per head in self-attention)."""
    _need_gpu()
    from speech_editing_toolkit_b200 import synth
    from speech_editing_toolkit_b200.modules import CampNetB200
    g = golden("campnet_t1024.npz")
    seed, T, vocab = int(g["seed"]), int(g["T"]), int(g["vocab"])
    b = synth.synthetic_campnet_batch(seed, 1, T, vocab=vocab)
    net = CampNetB200(vocab, 100, dict(hidden_size=192, dec_ffn_kernel_size=9, audio_num_mel_bins=80, b200_mode=mode)).cuda()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.campnet_state_dict(1234, vocab).items()}, strict=False)
    ret = net(cu(b["txt_tokens"]), mels=cu(b["mels"]), time_mel_masks=cu(b["time_mel_masks"]), infer=True)
    m = b["time_mel_masks"]
    fine, coarse = ret["mel_out_fine"].cpu().numpy(), ret["mel_out_coarse"].cpu().numpy()
    assert np.array_equal(fine * (1 - m), b["mels"] * (1 - m))
    e_abs = float(np.abs(fine - g["mel_out_fine"]).max())
    e_rel = rel_l1(fine * m, g["mel_out_fine"] * m)
    print(f"[margin] CampNet T=1024 {mode}: fine max-abs {e_abs:.3e}, rel-L1 over the predicted region {e_rel:.3e}")
    if mode == "simt_f32":
        assert e_abs < 2e-3 and np.abs(coarse - g["mel_out_coarse"]).max() < 2e-3
    else:
        assert e_rel < 3e-2 and rel_l1(coarse * m, g["mel_out_coarse"] * m) < 3e-2
        # the same item inside a batch of 64 x 1024 frames x 128 tokens (BASELINE configs[3]): no op of the forward mixes items, so its
        # mels must come out bit for bit as in the single-item run (the head-averaged attention weights are summed with atomics)
        nb = 64
        bb = synth.synthetic_campnet_batch(seed + 1, nb, T, vocab=vocab)
        for k in ("txt_tokens", "mels", "time_mel_masks"):
            bb[k][37] = b[k][0]
        rb = net(cu(bb["txt_tokens"]), mels=cu(bb["mels"]), time_mel_masks=cu(bb["time_mel_masks"]), infer=True)
        assert bool(torch.isfinite(rb["mel_out_fine"]).all())
        assert np.array_equal(rb["mel_out_fine"][37].cpu().numpy(), fine[0]) and np.array_equal(rb["mel_out_coarse"][37].cpu().numpy(), coarse[0])
        print(f"[margin] CampNet {mode}: item 37 of a 64 x {T} batch equals the single-item run bit for bit")
