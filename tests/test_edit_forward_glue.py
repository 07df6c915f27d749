"""SpecDenoiserInferB200.edit_forward — the host glue around the region surgery — exercised on the CPU with stand-ins for the
device calls (the engine's edit_* functions are replaced by the numpy oracle, the condition encoder by a stub that returns the
alignment the reference predicted), and checked against the tensors the reference's own forward_model handed to its model
(tests/golden/edit_region.npz): argument order, shapes, region handling and the final model call are the product's code."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import edit_region_oracle as EO
from test_edit_region_oracle import case


class _Eng:
    def dur_input(self, enc, style, txt):
        return enc


class _FS:
    def __init__(self, want):
        self.want, self.seen = want, {}

    def encoder(self, txt):
        return torch.zeros(txt.shape[0], txt.shape[1], 4)

    def forward_style_embed(self, spk, spk_id=None):
        return torch.zeros(spk.shape[0], 1, 4)

    def engine(self):
        return _Eng()

    def forward_dur(self, dur_inp, masks, mel2ph, txt, ret, masked_dur=None, use_pred_mel2ph=False):
        self.seen = dict(masks=masks, mel2ph=mel2ph, masked_dur=masked_dur, use_pred_mel2ph=use_pred_mel2ph)
        ret["dur"] = torch.zeros(txt.shape, dtype=torch.float32)
        return torch.from_numpy(self.want["edited_mel2ph_pred"])[None]


class _Model:
    def __init__(self, want):
        self.fs, self.kw = _FS(want), None

    def __call__(self, txt, **kw):
        self.kw = dict(kw, txt=txt)
        return {"mel_out": kw["ref_mels"].clone()}


@pytest.mark.parametrize("i", [0, 1, 2])
def test_edit_forward_hands_the_model_what_the_reference_does(i, monkeypatch):
    from speech_editing_toolkit_b200 import engine as E, plugin
    item, want = case(golden("edit_region.npz"), i)

    def fake_prepare(mel2ph, mel2word, ph2word, dur, regions, n_edited, T_len=None, Tp_len=None, Tpe_len=None):
        r = regions[0].tolist()
        md, mm, mo = EO.prepare(mel2ph[0].numpy(), mel2word[0].numpy(), ph2word[0].numpy(), dur[0].numpy(), n_edited, (r[0], r[1]))
        return torch.from_numpy(md)[None], torch.from_numpy(mm)[None], torch.from_numpy(mo)[None]

    def fake_assemble(mel2ph, mel2word, edited_ph2word, edited_mel2ph, regions, mel, f0, uv, T_len=None, Tpe_len=None, Te_len=None):
        r = regions[0].tolist()
        o = EO.assemble(mel2ph[0].numpy(), mel2word[0].numpy(), edited_ph2word[0].numpy(), edited_mel2ph[0].numpy(), (r[0], r[1]), (r[2], r[3]),
                        mel[0].numpy(), f0[0].numpy(), uv[0].numpy())
        out = {k: torch.from_numpy(o[k])[None] for k in ("mel2ph", "ref_mels", "f0", "uv", "time_mel_masks")}
        out["plan"] = torch.tensor([[*o["plan"], 0, 0, 0, 0]])
        return out

    monkeypatch.setattr(E, "edit_prepare", fake_prepare)
    monkeypatch.setattr(E, "edit_assemble", fake_assemble)
    inf = object.__new__(plugin.SpecDenoiserInferB200)
    inf.device, inf.model = torch.device("cpu"), _Model(want)
    inf.run_vocoder = lambda c: torch.zeros(c.shape[0], c.shape[1] * 256)
    sample = {k: torch.from_numpy(item[k])[None] for k in ("mel", "mel2ph", "mel2word", "dur", "ph2word", "edited_ph2word", "f0", "uv", "spk_embed")}
    sample.update(edited_txt_tokens=torch.from_numpy(item["edited_ph_token"])[None], words_region=item["words_region"],
                  edited_words_region=item["edited_words_region"], seed=1)
    wav, mel, aux = inf.forward_model(sample)
    fs, kw = inf.model.fs, inf.model.kw
    # what forward_dur received (inference/tts/spec_denoiser.py:98)
    assert fs.seen["use_pred_mel2ph"] is True
    assert np.array_equal(fs.seen["masked_dur"][0].numpy(), want["masked_dur"])
    assert np.array_equal(fs.seen["mel2ph"][0].numpy(), want["masked_mel2ph"])
    assert np.array_equal(fs.seen["masks"][0].numpy(), want["time_mel_masks_orig"].astype(np.float32))
    # what the model received (:133-135)
    assert kw["infer"] is True and kw["use_pred_pitch"] is True and kw["composite"] is True
    assert np.array_equal(kw["txt"][0].numpy(), item["edited_ph_token"])
    assert np.array_equal(kw["mel2ph"][0].numpy(), want["mel2ph"])
    assert np.array_equal(kw["time_mel_masks"][0].numpy(), want["time_mel_masks"]) and kw["time_mel_masks"].shape[-1] == 1
    for k in ("ref_mels", "f0", "uv"):
        assert np.array_equal(kw[k][0].numpy(), want[k]), k
    Tn = len(want["mel2ph"])
    assert mel.shape == (1, Tn, 80) and wav.shape == (1, Tn * 256) and int(aux["plan"][0, 0]) == Tn
    # tensor-form regions (a batch of them) are accepted as well
    sample2 = dict(sample, words_region=torch.tensor([item["words_region"][0]]), edited_words_region=torch.tensor([item["edited_words_region"][0]]))
    inf.forward_model(sample2)
    assert np.array_equal(inf.model.kw["mel2ph"][0].numpy(), want["mel2ph"])
