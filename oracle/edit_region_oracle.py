"""CPU oracle of the region surgery of the FluentSpeech inference script: the integer / index
work that turns the ORIGINAL utterance's alignment and the predicted alignment of the EDITED text into the model inputs.

TEST INFRASTRUCTURE — NOT PRODUCT CODE (same rules as oracle/fluentspeech_oracle.py).

numpy restatement of inference/tts/spec_denoiser.py:88-131 (reference tree, Zain-Jiang/Speech-Editing-Toolkit @ a8d5bf33),
one utterance (the reference runs batch 1):
  :88-91   masked_dur: durations of the phones before / after the edited word span
  :94-97   masked mel2ph / time mask of the original frames inside the span
  :99      # Deliberate divergence for PADDING: a predicted phone index 0 (a padded frame of a B > 1 batch; the reference runs batch 1, where
    # LengthRegulator emits no zeros) maps to word 0, i.e. to no word, instead of Python's negative-index wrap to the LAST word
    # (`edited_ph2word[-1]`), which would pull padding frames into the edited span whenever that span ends the utterance.  For
    # every index >= 1 this is the reference expression edited_ph2word[p - 1] (:99); the device core does the same.
    edited_mel2word = np.where(edited_mel2ph >= 1, edited_ph2word[np.maximum(edited_mel2ph, 1) - 1], 0)
  :100-110 length_edited, head_idx, tail_idx, edited_mel2ph_ (head copy, edited span, tail re-based by
           - min(tail) + max(edited span) + 2)
  :117-131 ref_mels / f0 / uv head + tail copies, time_mel_masks

Parity status: PINNED to the reference's own code: oracle/make_golden.py edit_region extracts `SpecDenoiserInfer.forward_model`
from the reference source (the module itself cannot be imported: it needs inference_acl, resemblyzer, ...), runs it with the real
reference FastSpeech and records the tensors it hands to the model (tests/golden/edit_region.npz).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def prepare(mel2ph, mel2word, ph2word, dur, n_edited_phones, words_region):
    """:88-97 -> masked_dur [Tpe], masked_mel2ph [T], time_mel_masks_orig [T] (int64, int64, float32)."""
    w0, w1 = words_region
    masked_dur = np.zeros(n_edited_phones, dtype=np.int64)
    n_head = int((ph2word < w0).sum())
    masked_dur[:n_head] = dur[:n_head]
    if ph2word.max() > w1:
        n_tail = int((ph2word > w1).sum())
        masked_dur[len(masked_dur) - n_tail:] = dur[len(dur) - n_tail:]
    region = (mel2word >= w0) & (mel2word <= w1)
    masked_mel2ph = np.where(region, 0, mel2ph).astype(np.int64)
    return masked_dur, masked_mel2ph, region.astype(F32)


def assemble(mel2ph, mel2word, edited_ph2word, edited_mel2ph, words_region, edited_words_region, mel, f0, uv):
    """:99-131 -> dict(mel2ph [Tn] int64, ref_mels [Tn,M], f0 [Tn], uv [Tn], time_mel_masks [Tn], plan=(Tn, head_idx, tail_idx, length_edited))."""
    w0, w1 = words_region
    c0, c1 = edited_words_region
    # Deliberate divergence for PADDING: a predicted phone index 0 (a padded frame of a B > 1 batch; the reference runs batch 1, where
    # LengthRegulator emits no zeros) maps to word 0, i.e. to no word, instead of Python's negative-index wrap to the LAST word
    # (`edited_ph2word[-1]`), which would pull padding frames into the edited span whenever that span ends the utterance.  For
    # every index >= 1 this is the reference expression edited_ph2word[p - 1] (:99); the device core does the same.
    edited_mel2word = np.where(edited_mel2ph >= 1, edited_ph2word[np.maximum(edited_mel2ph, 1) - 1], 0)
    sel_edit = (edited_mel2word >= c0) & (edited_mel2word <= c1)
    region = (mel2word >= w0) & (mel2word <= w1)
    length_edited = int(sel_edit.sum()) - int(region.sum())
    head_idx = int((mel2word < w0).sum())
    tail_idx = int((mel2word <= w1).sum()) + length_edited
    Tn = len(mel2ph) + length_edited
    out = np.zeros(Tn, dtype=np.int64)
    out[:head_idx] = mel2ph[:head_idx]
    out[head_idx:tail_idx] = edited_mel2ph[sel_edit]
    tail = mel2word > w1
    if mel2word.max() > w1:
        out[tail_idx:] = mel2ph[tail] - mel2ph[tail].min() + edited_mel2ph[sel_edit].max() + 2
    ref = np.zeros((Tn, mel.shape[1]), dtype=F32)
    ref[:head_idx] = mel[:head_idx]
    ref[tail_idx:] = mel[tail]
    ef0, euv = np.zeros(Tn, dtype=F32), np.zeros(Tn, dtype=F32)
    ef0[:head_idx], euv[:head_idx] = f0[:head_idx], uv[:head_idx]
    ef0[tail_idx:], euv[tail_idx:] = f0[tail], uv[tail]
    mask = np.zeros(Tn, dtype=F32)
    mask[head_idx:tail_idx] = 1.0
    return dict(mel2ph=out, ref_mels=ref, f0=ef0, uv=euv, time_mel_masks=mask, plan=(Tn, head_idx, tail_idx, length_edited))
