"""Pins oracle/campnet_oracle.py (numpy restatement of CampNet.forward, campnet.py:40-69) against tests/golden/campnet.npz —
outputs of the unmodified reference CampNet on a ragged batch (oracle/make_golden.py campnet)."""
import numpy as np

from conftest import golden, rel_l1
from oracle import campnet_oracle as KO
from speech_editing_toolkit_b200 import synth


def _case():
    g = golden("campnet.npz")
    seed, B, T, vocab = int(g["seed"]), int(g["B"]), int(g["T"]), int(g["vocab"])
    return g, synth.campnet_state_dict(seed, vocab), synth.synthetic_campnet_batch(seed, B, T, vocab=vocab, pad_items=[(1, 4)])


def test_positional_table_known_answers():
    tab = KO.sinusoid_table(2000, 192)
    assert tab.shape == (2000, 192) and np.abs(tab[0]).max() == 0.0                # padding row
    assert abs(tab[1, 0] - np.sin(1.0)) < 1e-6 and abs(tab[1, 96] - np.cos(1.0)) < 1e-6 and abs(tab[1, 95] - 1e-4) < 1e-8
    pos = KO.make_positions(np.array([[5, 7, 0, 9, 0]]))
    assert np.array_equal(pos, [[1, 2, 0, 3, 0]])


def test_campnet_forward_matches_reference_fixture():
    g, sd, b = _case()
    assert (b["txt_tokens"][1, -4:] == 0).all() and np.abs(b["mels"][1, -32:]).max() == 0.0          # ragged second item
    out = KO.campnet_forward(sd, b["txt_tokens"], b["mels"], b["time_mel_masks"])
    assert np.abs(out["encoder_out"] - g["encoder_out"]).max() < 5e-5
    assert np.abs(out["mel_out_coarse"] - g["mel_out_coarse"]).max() < 2e-4
    assert np.abs(out["mel_out_fine"] - g["mel_out_fine"]).max() < 2e-4
    assert np.abs(out["attn"] - g["attn"].astype(np.float32)).max() < 1e-3                            # fixture stores fp16
    assert np.abs(out["attn"].sum(-1) - 1).max() < 1e-5 and np.abs(out["attn"][1, :, -4:]).max() == 0.0   # padded keys get no mass
    # unmasked frames pass through unchanged, padded frames stay zero
    m = b["time_mel_masks"]
    assert np.array_equal((out["mel_out_fine"] * (1 - m)), (b["mels"] * (1 - m)))
    assert np.abs(out["mel_out_coarse"][1, -32:]).max() == 0.0


def test_bf16_contract_within_stated_tolerance():
    g, sd, b = _case()
    out = KO.campnet_forward(sd, b["txt_tokens"], b["mels"], b["time_mel_masks"], gemm_dtype="bf16", attn_dtype="bf16")
    m = b["time_mel_masks"]
    assert rel_l1(out["mel_out_coarse"] * m, g["mel_out_coarse"] * m) < 3e-2
    assert rel_l1(out["mel_out_fine"] * m, g["mel_out_fine"] * m) < 3e-2
