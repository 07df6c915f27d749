// Region surgery of the FluentSpeech inference script (inference/tts/spec_denoiser.py:88-131 of the reference): the integer /
// index work that turns the original utterance's alignment plus the predicted alignment of the edited text into the model's
// inputs.  Per-item logic as plain functions usable on the host and on the device (FSE_HD): the CUDA kernels of
// edit_region.cu are thin loops around them, and tests/tools/edit_region_host.cpp compiles the very same functions with g++
// so that the logic is checked bit-exactly on the CPU-only build container as well.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FSE_HD __host__ __device__ inline
#else
#define FSE_HD inline
#endif

namespace fse {
namespace edit {

// one utterance; regions are 1-based inclusive word indices as in the reference (words_region[0], edited_words_region[0])
struct Item {
  const int64_t* mel2ph;          // [T]   1-based phone index of every frame
  const int64_t* mel2word;        // [T]   1-based word index of every frame
  int T;
  const int64_t* ph2word;         // [Tp]  word index of every phone of the original text
  const int64_t* dur;             // [Tp]  frames per phone of the original utterance
  int Tp;
  const int64_t* edited_ph2word;  // [Tpe] word index of every phone of the edited text
  int Tpe;
  int64_t w0, w1;                 // edited span in the original words
  int64_t c0, c1;                 // the span that replaces it, in the edited words
};

enum { kPlanTn = 0, kPlanHead = 1, kPlanTail = 2, kPlanLenEdited = 3, kPlanNEdit = 4, kPlanNTail = 5, kPlanTailShift = 6, kPlanHasTail = 7, kPlanSize = 8 };

// :88-97  masked_dur [Tpe_stride] (zero past Tpe), masked_mel2ph [T], time_mel_masks_orig [T]
FSE_HD void prepare_item(const Item& it, int64_t* masked_dur, int Tpe_stride, int64_t* masked_mel2ph, float* mask_orig) {
  for (int i = 0; i < Tpe_stride; ++i) masked_dur[i] = 0;
  int n_head = 0, n_tail = 0;
  int64_t wmax = 0;
  for (int i = 0; i < it.Tp; ++i) {
    if (it.ph2word[i] < it.w0) ++n_head;
    if (it.ph2word[i] > it.w1) ++n_tail;
    if (i == 0 || it.ph2word[i] > wmax) wmax = it.ph2word[i];
  }
  for (int i = 0; i < n_head && i < it.Tpe; ++i) masked_dur[i] = it.dur[i];
  if (it.Tp > 0 && wmax > it.w1)
    for (int i = 0; i < n_tail; ++i) {
      const int dst = it.Tpe - n_tail + i, src = it.Tp - n_tail + i;
      if (dst >= 0 && src >= 0) masked_dur[dst] = it.dur[src];
    }
  for (int t = 0; t < it.T; ++t) {
    const bool in = it.mel2word[t] >= it.w0 && it.mel2word[t] <= it.w1;
    masked_mel2ph[t] = in ? 0 : it.mel2ph[t];
    mask_orig[t] = in ? 1.f : 0.f;
  }
}

// :99-110  the plan of the assembled sequence and the (order-preserving) selections it copies from:
//   sel_edit[k] = index into edited_mel2ph of the k-th frame whose word lies in [c0, c1]
//   sel_tail[k] = index into the original frames of the k-th frame whose word lies after w1
FSE_HD void plan_item(const Item& it, const int64_t* edited_mel2ph, int Te, int32_t* sel_edit, int32_t* sel_tail, int64_t* plan) {
  int n_edit = 0, n_region = 0, n_before = 0, n_upto = 0, n_tail = 0;
  int64_t edit_max = 0, tail_min = 0, wmax = 0;
  for (int i = 0; i < Te; ++i) {
    const int64_t ph = edited_mel2ph[i];
    // edited_mel2word (:99).  ph == 0 is a PADDED frame of a B > 1 batch (the reference runs batch 1 and never sees one): it belongs
    // to no word here, whereas Python's edited_ph2word[p - 1] would wrap to the last word; the oracle restates the same rule.
    const int64_t w = (ph >= 1 && ph <= it.Tpe) ? it.edited_ph2word[ph - 1] : 0;
    if (w >= it.c0 && w <= it.c1) {
      if (n_edit == 0 || ph > edit_max) edit_max = ph;
      sel_edit[n_edit++] = i;
    }
  }
  for (int t = 0; t < it.T; ++t) {
    const int64_t w = it.mel2word[t];
    if (w >= it.w0 && w <= it.w1) ++n_region;
    if (w < it.w0) ++n_before;
    if (w <= it.w1) ++n_upto;
    if (t == 0 || w > wmax) wmax = w;
    if (w > it.w1) {
      if (n_tail == 0 || it.mel2ph[t] < tail_min) tail_min = it.mel2ph[t];
      sel_tail[n_tail++] = t;
    }
  }
  const int64_t le = static_cast<int64_t>(n_edit) - n_region;
  plan[kPlanLenEdited] = le;
  plan[kPlanHead] = n_before;
  plan[kPlanTail] = n_upto + le;
  plan[kPlanTn] = it.T + le;
  plan[kPlanNEdit] = n_edit;
  plan[kPlanNTail] = n_tail;
  plan[kPlanTailShift] = -tail_min + edit_max + 2;          // mel2ph[tail] - min(mel2ph[tail]) + max(edited span) + 2 (:108)
  plan[kPlanHasTail] = (it.T > 0 && wmax > it.w1) ? 1 : 0;
}

// :103-131 one output frame i < Tn: the phone index, the source frame of the mel / f0 / uv copies (-1: zeros) and the mask
FSE_HD void assemble_frame(const Item& it, const int64_t* plan, const int64_t* edited_mel2ph, const int32_t* sel_edit, const int32_t* sel_tail,
                           int i, int64_t* mel2ph_out, int* src_frame, float* mask) {
  const int64_t head = plan[kPlanHead], tail = plan[kPlanTail];
  if (i < head) {
    *mel2ph_out = it.mel2ph[i]; *src_frame = i; *mask = 0.f;
  } else if (i < tail) {
    const int64_t k = i - head;
    *mel2ph_out = k < plan[kPlanNEdit] ? edited_mel2ph[sel_edit[k]] : 0;
    *src_frame = -1; *mask = 1.f;
  } else {
    const int64_t k = i - tail;
    const bool ok = k < plan[kPlanNTail];
    *src_frame = ok ? sel_tail[k] : -1;                      // ref_mels / f0 / uv tails are copied unconditionally (:121,127,129)
    *mel2ph_out = (ok && plan[kPlanHasTail]) ? it.mel2ph[sel_tail[k]] + plan[kPlanTailShift] : 0;
    *mask = 0.f;
  }
}

}  // namespace edit
}  // namespace fse
