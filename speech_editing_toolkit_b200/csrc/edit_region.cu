// Batched region surgery of the FluentSpeech inference script (inference/tts/spec_denoiser.py:88-131) behind the C ABI of
// include/fse_b200.h: fse_edit_prepare / fse_edit_plan / fse_edit_assemble.  The reference does this work for ONE utterance
// with host-side tensor slicing; here every item of a padded batch is handled on the device (per-item lengths, per-item
// regions), so that editing can run with B > 1.  Integer / index work only: bit-exact by construction.  The per-item logic
// lives in edit_region_core.h (host + device); these kernels are thin loops around it:
//   prepare, plan   one thread per item (a serial pass over a few hundred phones / a few thousand frames, once per batch)
//   assemble        one warp per output frame: lane-uniform index computation, lanes copy the mel row
#include <cuda_runtime.h>

#include "edit_region_core.h"
#include "fse_common.cuh"

namespace fse {
namespace {

struct Batch {
  const int64_t *mel2ph, *mel2word, *T_len;             // [B,T], [B,T], [B] or null (= T)
  const int64_t *ph2word, *dur, *Tp_len;                // [B,Tp], [B,Tp], [B] or null
  const int64_t *edited_ph2word, *Tpe_len;              // [B,Tpe], [B] or null
  const int64_t* regions;                               // [B,4]: w0, w1, c0, c1
  int B, T, Tp, Tpe;
};

__device__ __forceinline__ edit::Item make_item(const Batch& bt, int b) {
  edit::Item it;
  it.mel2ph = bt.mel2ph + static_cast<size_t>(b) * bt.T;
  it.mel2word = bt.mel2word + static_cast<size_t>(b) * bt.T;
  it.T = bt.T_len ? static_cast<int>(bt.T_len[b]) : bt.T;
  it.T = it.T < 0 ? 0 : (it.T > bt.T ? bt.T : it.T);
  it.ph2word = bt.ph2word ? bt.ph2word + static_cast<size_t>(b) * bt.Tp : nullptr;
  it.dur = bt.dur ? bt.dur + static_cast<size_t>(b) * bt.Tp : nullptr;
  it.Tp = bt.ph2word ? (bt.Tp_len ? static_cast<int>(bt.Tp_len[b]) : bt.Tp) : 0;
  it.Tp = it.Tp < 0 ? 0 : (it.Tp > bt.Tp ? bt.Tp : it.Tp);
  it.edited_ph2word = bt.edited_ph2word ? bt.edited_ph2word + static_cast<size_t>(b) * bt.Tpe : nullptr;
  it.Tpe = bt.Tpe_len ? static_cast<int>(bt.Tpe_len[b]) : bt.Tpe;
  it.Tpe = it.Tpe < 0 ? 0 : (it.Tpe > bt.Tpe ? bt.Tpe : it.Tpe);
  it.w0 = bt.regions[4 * b]; it.w1 = bt.regions[4 * b + 1]; it.c0 = bt.regions[4 * b + 2]; it.c1 = bt.regions[4 * b + 3];
  return it;
}

__global__ void edit_prepare_kernel(Batch bt, int64_t* __restrict__ masked_dur, int64_t* __restrict__ masked_mel2ph, float* __restrict__ mask_orig) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bt.B) return;
  const edit::Item it = make_item(bt, b);
  int64_t* mm = masked_mel2ph + static_cast<size_t>(b) * bt.T;
  float* mo = mask_orig + static_cast<size_t>(b) * bt.T;
  edit::prepare_item(it, masked_dur + static_cast<size_t>(b) * bt.Tpe, bt.Tpe, mm, mo);
  for (int t = it.T; t < bt.T; ++t) { mm[t] = 0; mo[t] = 0.f; }           // padding frames of a ragged batch
}

__global__ void edit_plan_kernel(Batch bt, const int64_t* __restrict__ edited_mel2ph, const int64_t* __restrict__ Te_len, int Te,
                                 int32_t* __restrict__ sel_edit, int32_t* __restrict__ sel_tail, int64_t* __restrict__ plan) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bt.B) return;
  const edit::Item it = make_item(bt, b);
  int te = Te_len ? static_cast<int>(Te_len[b]) : Te;
  te = te < 0 ? 0 : (te > Te ? Te : te);
  edit::plan_item(it, edited_mel2ph + static_cast<size_t>(b) * Te, te, sel_edit + static_cast<size_t>(b) * Te,
                  sel_tail + static_cast<size_t>(b) * bt.T, plan + static_cast<size_t>(b) * edit::kPlanSize);
}

__global__ void __launch_bounds__(256) edit_assemble_kernel(Batch bt, const int64_t* __restrict__ plan, const int64_t* __restrict__ edited_mel2ph, int Te,
                                                            const int32_t* __restrict__ sel_edit, const int32_t* __restrict__ sel_tail,
                                                            const float* __restrict__ mel, const float* __restrict__ f0, const float* __restrict__ uv,
                                                            int64_t* __restrict__ out_mel2ph, float* __restrict__ out_ref, float* __restrict__ out_f0,
                                                            float* __restrict__ out_uv, float* __restrict__ out_mask, int Tn, int M) {
  const size_t row = static_cast<size_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<size_t>(bt.B) * Tn) return;
  const int b = static_cast<int>(row / Tn), i = static_cast<int>(row % Tn);
  const int64_t* pl = plan + static_cast<size_t>(b) * edit::kPlanSize;
  int64_t ph = 0;
  int src = -1;
  float m = 0.f;
  if (i < pl[edit::kPlanTn]) {
    const edit::Item it = make_item(bt, b);
    edit::assemble_frame(it, pl, edited_mel2ph + static_cast<size_t>(b) * Te, sel_edit + static_cast<size_t>(b) * Te,
                         sel_tail + static_cast<size_t>(b) * bt.T, i, &ph, &src, &m);
  }
  if (lane == 0) {
    out_mel2ph[row] = ph;
    out_mask[row] = m;
    out_f0[row] = (src >= 0 && f0) ? f0[static_cast<size_t>(b) * bt.T + src] : 0.f;
    out_uv[row] = (src >= 0 && uv) ? uv[static_cast<size_t>(b) * bt.T + src] : 0.f;
  }
  const float* srow = src >= 0 ? mel + (static_cast<size_t>(b) * bt.T + src) * M : nullptr;
  for (int c = lane; c < M; c += 32) out_ref[row * M + c] = srow ? srow[c] : 0.f;
}

int check_device() {
  int dev = 0;
  FSE_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  FSE_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(FSE_ECUDA, "device is sm_%d%d; this library is built for sm_100a only (no fallback)", prop.major, prop.minor);
  return FSE_OK;
}

}  // namespace
}  // namespace fse

using namespace fse;

extern "C" {

int fse_edit_prepare(const int64_t* mel2ph, const int64_t* mel2word, const int64_t* T_len, const int64_t* ph2word, const int64_t* dur,
                     const int64_t* Tp_len, const int64_t* Tpe_len, const int64_t* regions, int64_t* masked_dur, int64_t* masked_mel2ph,
                     float* time_mel_masks_orig, int32_t B, int32_t T, int32_t Tp, int32_t Tpe, void* stream) {
  if (!mel2ph || !mel2word || !ph2word || !dur || !regions || !masked_dur || !masked_mel2ph || !time_mel_masks_orig) return fail(FSE_EINVAL, "null argument");
  if (B <= 0 || T <= 0 || Tp <= 0 || Tpe <= 0) return fail(FSE_EINVAL, "B, T, Tp and Tpe must be positive");
  FSE_TRY(check_device());
  Batch bt{mel2ph, mel2word, T_len, ph2word, dur, Tp_len, nullptr, Tpe_len, regions, B, T, Tp, Tpe};
  edit_prepare_kernel<<<(B + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(bt, masked_dur, masked_mel2ph, time_mel_masks_orig);
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

int fse_edit_plan(const int64_t* mel2ph, const int64_t* mel2word, const int64_t* T_len, const int64_t* edited_ph2word, const int64_t* Tpe_len,
                  const int64_t* regions, const int64_t* edited_mel2ph, const int64_t* Te_len, int32_t* sel_edit, int32_t* sel_tail, int64_t* plan,
                  int32_t B, int32_t T, int32_t Tpe, int32_t Te, void* stream) {
  if (!mel2ph || !mel2word || !edited_ph2word || !regions || !edited_mel2ph || !sel_edit || !sel_tail || !plan) return fail(FSE_EINVAL, "null argument");
  if (B <= 0 || T <= 0 || Tpe <= 0 || Te <= 0) return fail(FSE_EINVAL, "B, T, Tpe and Te must be positive");
  FSE_TRY(check_device());
  Batch bt{mel2ph, mel2word, T_len, nullptr, nullptr, nullptr, edited_ph2word, Tpe_len, regions, B, T, 0, Tpe};
  edit_plan_kernel<<<(B + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(bt, edited_mel2ph, Te_len, Te, sel_edit, sel_tail, plan);
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

int fse_edit_assemble(const int64_t* mel2ph, const int64_t* T_len, const int64_t* regions, const int64_t* plan, const int64_t* edited_mel2ph,
                      const int32_t* sel_edit, const int32_t* sel_tail, const float* mel, const float* f0, const float* uv, int64_t* out_mel2ph,
                      float* out_ref_mels, float* out_f0, float* out_uv, float* out_time_mel_masks, int32_t B, int32_t T, int32_t Te, int32_t Tn,
                      int32_t n_mels, void* stream) {
  if (!mel2ph || !regions || !plan || !edited_mel2ph || !sel_edit || !sel_tail || !mel || !out_mel2ph || !out_ref_mels || !out_f0 || !out_uv ||
      !out_time_mel_masks)
    return fail(FSE_EINVAL, "null argument");
  if (B <= 0 || T <= 0 || Te <= 0 || Tn <= 0 || n_mels <= 0) return fail(FSE_EINVAL, "B, T, Te, Tn and n_mels must be positive");
  FSE_TRY(check_device());
  Batch bt{mel2ph, mel2ph /*mel2word is not needed once the plan exists*/, T_len, nullptr, nullptr, nullptr, nullptr, nullptr, regions, B, T, 0, 1};
  const size_t rows = static_cast<size_t>(B) * Tn;
  edit_assemble_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      bt, plan, edited_mel2ph, Te, sel_edit, sel_tail, mel, f0, uv, out_mel2ph, out_ref_mels, out_f0, out_uv, out_time_mel_masks, Tn, n_mels);
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

}  // extern "C"
