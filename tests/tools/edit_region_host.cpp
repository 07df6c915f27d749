// TEST INFRASTRUCTURE: compiles the per-item functions of speech_editing_toolkit_b200/csrc/edit_region_core.h — the very
// code the CUDA kernels of edit_region.cu loop over — with g++, so that tests/test_edit_region_core.py can check the logic
// bit-exactly against the oracle and the reference fixture on the CPU-only build container.  Not part of the shipped library.
#include <cstring>
#include <vector>

#include "../../speech_editing_toolkit_b200/csrc/edit_region_core.h"

using namespace fse::edit;

extern "C" {

void er_prepare(const int64_t* mel2ph, const int64_t* mel2word, int T, const int64_t* ph2word, const int64_t* dur, int Tp, int Tpe, int Tpe_stride,
                const int64_t* region, int64_t* masked_dur, int64_t* masked_mel2ph, float* mask_orig) {
  Item it{mel2ph, mel2word, T, ph2word, dur, Tp, nullptr, Tpe, region[0], region[1], region[2], region[3]};
  prepare_item(it, masked_dur, Tpe_stride, masked_mel2ph, mask_orig);
}

// plan + assembly of one item; outputs sized by the caller from a first call with out_mel2ph == nullptr (returns Tn)
long long er_assemble(const int64_t* mel2ph, const int64_t* mel2word, int T, const int64_t* edited_ph2word, int Tpe, const int64_t* region,
                      const int64_t* edited_mel2ph, int Te, const float* mel, const float* f0, const float* uv, int M, int64_t* plan_out,
                      int64_t* out_mel2ph, float* out_ref, float* out_f0, float* out_uv, float* out_mask) {
  Item it{mel2ph, mel2word, T, nullptr, nullptr, 0, edited_ph2word, Tpe, region[0], region[1], region[2], region[3]};
  std::vector<int32_t> sel_edit(Te > 0 ? Te : 1), sel_tail(T > 0 ? T : 1);
  int64_t plan[kPlanSize];
  plan_item(it, edited_mel2ph, Te, sel_edit.data(), sel_tail.data(), plan);
  if (plan_out) std::memcpy(plan_out, plan, sizeof(plan));
  if (!out_mel2ph) return plan[kPlanTn];
  for (int i = 0; i < plan[kPlanTn]; ++i) {
    int64_t ph; int src; float m;
    assemble_frame(it, plan, edited_mel2ph, sel_edit.data(), sel_tail.data(), i, &ph, &src, &m);
    out_mel2ph[i] = ph; out_mask[i] = m;
    out_f0[i] = src >= 0 ? f0[src] : 0.f;
    out_uv[i] = src >= 0 ? uv[src] : 0.f;
    for (int c = 0; c < M; ++c) out_ref[static_cast<size_t>(i) * M + c] = src >= 0 ? mel[static_cast<size_t>(src) * M + c] : 0.f;
  }
  return plan[kPlanTn];
}

}  // extern "C"
