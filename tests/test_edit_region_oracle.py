"""Pins oracle/edit_region_oracle.py (the region surgery of inference/tts/spec_denoiser.py:88-131) against
tests/golden/edit_region.npz — tensors recorded from the reference's OWN forward_model code (oracle/make_golden.py edit_region)."""
import numpy as np
import pytest

from conftest import golden
from oracle import edit_region_oracle as EO
from speech_editing_toolkit_b200 import synth


def case(g, i):
    kw = g[f"c{i}_kw"]
    item = synth.synthetic_edit_item(int(kw[0]), n_words=int(kw[1]), vocab=int(g["vocab"]), edit_span=(int(kw[2]), int(kw[3])),
                                     new_span_phones=tuple(int(x) for x in kw[5:5 + int(kw[4])]))
    return item, {k[len(f"c{i}_"):]: g[k] for k in g.files if k.startswith(f"c{i}_")}


@pytest.mark.parametrize("i", [0, 1, 2])
def test_region_surgery_matches_reference_code(i):
    item, want = case(golden("edit_region.npz"), i)
    md, mm, tm = EO.prepare(item["mel2ph"], item["mel2word"], item["ph2word"], item["dur"], len(item["edited_ph2word"]), item["words_region"][0])
    assert np.array_equal(md, want["masked_dur"]) and np.array_equal(mm, want["masked_mel2ph"])
    assert np.array_equal(tm, want["time_mel_masks_orig"].astype(np.float32))
    out = EO.assemble(item["mel2ph"], item["mel2word"], item["edited_ph2word"], want["edited_mel2ph_pred"], item["words_region"][0],
                      item["edited_words_region"][0], item["mel"], item["f0"], item["uv"])
    assert np.array_equal(out["mel2ph"], want["mel2ph"])                                   # integer: bit-exact
    assert np.array_equal(out["time_mel_masks"], want["time_mel_masks"][:, 0])
    for k in ("ref_mels", "f0", "uv"):
        assert np.array_equal(out[k], want[k]), k                                          # copies: bit-exact
    Tn, head, tail, le = out["plan"]
    assert Tn == len(item["mel2ph"]) + le and 0 <= head <= tail <= Tn and int(out["time_mel_masks"].sum()) == tail - head


def test_fixture_covers_head_tail_and_growth():
    g = golden("edit_region.npz")
    plans = []
    for i in range(3):
        item, want = case(g, i)
        plans.append((len(item["mel2ph"]), len(want["mel2ph"]), int(want["time_mel_masks"].argmax()), int(want["time_mel_masks"].sum())))
    assert plans[1][2] + plans[1][3] == plans[1][1]          # case 1: the edit runs to the end (no tail)
    assert plans[2][2] == 0                                  # case 2: the edit starts at frame 0 (no head)
    assert plans[0][2] > 0 and plans[0][2] + plans[0][3] < plans[0][1]
