// Condition encoder of FluentSpeech — FastSpeech.forward(skip_decoder=True) — behind the C ABI of include/fse_b200.h.
// Reference (file:line relative to the reference tree):
//   modules/speech_editing/spec_denoiser/fs.py:83-189      FastSpeech.forward / forward_style_embed / forward_dur / forward_pitch
//   modules/commons/conv.py:24-139                         ResidualBlock (LN -> Conv k5 -> *k^-0.5 -> GELU -> Conv 1x1), ConvBlocks, TextConvEncoder
//   modules/commons/nar_tts_modules.py:8-100               DurationPredictor, LengthRegulator, PitchPredictor
//   modules/tts/commons/align_ops.py:21-25                 expand_states
//   utils/audio/align.py:71-90, utils/audio/pitch/utils.py:17-28,71-82   mel2token_to_dur, f0_to_coarse, denorm_f0
//
// Layout: everything is channels-last [B, T, C] (a frame / token is a GEMM row).  Every Conv1d is one launch of the
// conv-as-shifted-GEMM primitive (conv_gemm.cuh: TMA -> smem ring -> tcgen05.mma -> TMEM -> epilogue functor) with the
// pointwise tail of the layer in its epilogue; LayerNorm is a warp-per-row kernel (warp-shuffle reductions) that writes the
// next GEMM's operand; embedding gathers, the 192->1 / 192->2 heads and the integer ops (duration histogram, length
// regulator, pitch bins) are small fp32 / int64 CUDA-core kernels.  The path runs once per batch in front of the
// sampling loop (~2 MFLOP per frame against 25 MFLOP per frame and diffusion step), so the goal here is that no torch
// arithmetic is left between the text tokens and `cond`, not peak throughput.
#include "rowwise.cuh"

namespace fse {
namespace {

// x = embed_scale * embed_tokens(txt) (conv.py:138) and the ConvBlocks-level nonpadding (|x|.sum(c) > 0, conv.py:104)
__global__ void __launch_bounds__(256) cond_embed_tokens_kernel(const int64_t* __restrict__ txt, const float* __restrict__ table,
                                                                float* __restrict__ x, float* __restrict__ mask0, int rows, int C,
                                                                int vocab, float scale) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  long long tok = txt[row];
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
  const float* e = table + static_cast<size_t>(tok) * C;
  float sa = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = __fmul_rn(scale, __ldg(e + c));
    x[static_cast<size_t>(row) * C + c] = v;
    sa += fabsf(v);
  }
  sa = warp_sum(sa);
  if (lane == 0) mask0[row] = sa > 0.f ? 1.f : 0.f;
}

// dur_inp = (encoder_out + style) * (txt > 0)   (fs.py:90)
__global__ void __launch_bounds__(256) cond_dur_input_kernel(const float* __restrict__ enc, const float* __restrict__ style,
                                                             const int64_t* __restrict__ txt, float* __restrict__ out, int rows,
                                                             int Tt, int C) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / Tt;
  const float m = txt[row] > 0 ? 1.f : 0.f;
  for (int c = lane; c < C; c += 32) {
    const float s = style ? __ldg(style + static_cast<size_t>(b) * C + c) : 0.f;
    out[static_cast<size_t>(row) * C + c] = __fmul_rn(__fadd_rn(enc[static_cast<size_t>(row) * C + c], s), m);
  }
}

// operand of the first duration-predictor conv: dur_inp + dur_embed(masked_dur)  (fs.py:139/141); keep[row] = (txt != 0)
template <typename TOp>
__global__ void __launch_bounds__(256) cond_dur_embed_add_kernel(const float* __restrict__ dur_inp, const int64_t* __restrict__ idx,
                                                                 const float* __restrict__ table, int n_table,
                                                                 const int64_t* __restrict__ txt, TOp* __restrict__ out,
                                                                 float* __restrict__ keep, int rows, int C) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  long long d = idx[row];
  d = d < 0 ? 0 : (d >= n_table ? n_table - 1 : d);
  const float* e = table + static_cast<size_t>(d) * C;
  for (int c = lane; c < C; c += 32)
    store_op(out + static_cast<size_t>(row) * C + c, __fadd_rn(dur_inp[static_cast<size_t>(row) * C + c], __ldg(e + c)));
  if (lane == 0) keep[row] = txt[row] != 0 ? 1.f : 0.f;
}

// the predictor heads: y[row, o] = <x[row, :], w[o, :]> + b[o], o < O (O = 1: duration, 2: pitch), fp32 CUDA cores.
// kDur: Softplus(beta 1, threshold 20) then * keep[row]   (nar_tts_modules.py:22,31-33)
template <bool kDur>
__global__ void __launch_bounds__(256) cond_head_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, const float* __restrict__ keep,
                                                        float* __restrict__ out, int rows, int C, int O) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<size_t>(row) * C;
  for (int o = 0; o < O; ++o) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(xr[c], __ldg(w + static_cast<size_t>(o) * C + c), s);
    s = warp_sum(s) + __ldg(bias + o);
    if (kDur) {
      s = s > 20.f ? s : log1pf(expf(s));
      s = __fmul_rn(s, keep[row]);
    }
    if (lane == 0) out[static_cast<size_t>(row) * O + o] = s;
  }
}

// style_embed = spk_embed_proj(spk_embed)   (fs.py:117-118): one block per item, one thread per output channel
__global__ void cond_style_kernel(const float* __restrict__ spk, const float* __restrict__ w, const float* __restrict__ bias,
                                  float* __restrict__ out, int D, int C) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < D; ++k) s = fmaf(__ldg(spk + static_cast<size_t>(b) * D + k), __ldg(w + static_cast<size_t>(c) * D + k), s);
    out[static_cast<size_t>(b) * C + c] = s + __ldg(bias + c);
  }
}

// ---------------------------------------------------------------------------------------------- integer kernels
// masked_dur[b, i] += 1 for every frame with mel2ph * (1 - mask).long() == i + 1, times (txt != 0)   (fs.py:136-138,
// utils/audio/align.py:82: scatter_add of ones; bucket 0 = padding is dropped).  The output is zeroed by the caller.
__global__ void __launch_bounds__(256) cond_masked_dur_kernel(const int64_t* __restrict__ mel2ph, const float* __restrict__ mask,
                                                              const int64_t* __restrict__ txt, int64_t* __restrict__ out, int B,
                                                              int T, int Tt) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * T) return;
  const int b = static_cast<int>(i / T);
  const long long keep = mask ? static_cast<long long>(1.0f - mask[i]) : 1;
  const long long idx = mel2ph[i] * keep;
  if (idx >= 1 && idx <= Tt && txt[static_cast<size_t>(b) * Tt + idx - 1] != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(out + static_cast<size_t>(b) * Tt + idx - 1), 1ULL);
}

// LengthRegulator (nar_tts_modules.py:63-67): dur = round(dur).long() * (1 - padding); cumsum.  One thread per item
// (Tt is a few hundred tokens and this runs once per batch).
__global__ void cond_length_cumsum_kernel(const float* __restrict__ dur, const int64_t* __restrict__ txt, int64_t* __restrict__ cumsum,
                                          int64_t* __restrict__ totals, int B, int Tt) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  long long acc = 0;
  for (int i = 0; i < Tt; ++i) {
    long long d = static_cast<long long>(rintf(dur[static_cast<size_t>(b) * Tt + i]));   // torch.round: half to even
    if (txt && txt[static_cast<size_t>(b) * Tt + i] == 0) d = 0;
    acc += d;
    cumsum[static_cast<size_t>(b) * Tt + i] = acc;
  }
  totals[b] = acc;
}

// mel2ph[b, pos] = i + 1 for the token i with cumsum[i-1] <= pos < cumsum[i], 0 past the item's end (:69-72)
__global__ void __launch_bounds__(256) cond_length_fill_kernel(const int64_t* __restrict__ cumsum, int64_t* __restrict__ mel2ph, int B,
                                                               int Tt, int Tmax) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * Tmax) return;
  const int b = static_cast<int>(i / Tmax);
  const long long pos = static_cast<long long>(i % Tmax);
  const int64_t* cs = cumsum + static_cast<size_t>(b) * Tt;
  int lo = 0, hi = Tt;                       // first index with cs[idx] > pos
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cs[mid] > pos) hi = mid; else lo = mid + 1;
  }
  mel2ph[i] = lo < Tt ? lo + 1 : 0;
}

// ---------------------------------------------------------------------------------------------- pitch helpers
struct PitchConst {
  float mel_min;   // 1127 ln(1 + 50/700), float64 -> fp32 as torch does with the numpy scalar
  float mel_rng;   // (mel_max - mel_min), float64 difference -> fp32
};

// denorm_f0 (utils/audio/pitch/utils.py:71-82), pitch_norm = 'log': 2 ** f0, clamp [50, 900], zero where unvoiced / padded.
// exp2 / log are evaluated in double and rounded once so the result does not depend on the fast-math flavour of the build.
__device__ __forceinline__ float denorm_f0_dev(float f0, bool use_uv, float uv, bool pad) {
  float y = static_cast<float>(exp2(static_cast<double>(f0)));
  y = fminf(fmaxf(y, 50.f), 900.f);
  if (use_uv && uv > 0.f) y = 0.f;
  if (pad) y = 0.f;
  return y;
}
// f0_to_coarse (utils/audio/pitch/utils.py:17-28): 1127 ln(1 + f0/700) mapped to bins 1..255, truncation after + 0.5
__device__ __forceinline__ long long f0_to_coarse_dev(float f0, PitchConst pc) {
  float m = __fmul_rn(1127.f, static_cast<float>(log(static_cast<double>(__fadd_rn(1.f, __fdiv_rn(f0, 700.f))))));
  if (m > 0.f) m = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(m, pc.mel_min), 254.f), pc.mel_rng), 1.f);
  if (m <= 1.f) m = 1.f;
  if (m > 255.f) m = 255.f;
  return static_cast<long long>(__fadd_rn(m, 0.5f));
}

// expand_states (align_ops.py:21-25) + the pitch predictor's input (fs.py:96, 159-166):
//   dec[row] = mel2ph > 0 ? encoder_out[b, mel2ph - 1] : 0                                   -> dec32 (fp32)
//   pitch_inp = (dec + style) * (mel2ph > 0) + pitch_embed(f0_to_coarse(denorm_f0(f0 (1-m), uv (1-m), pad)))   -> operand
template <typename TOp>
__global__ void __launch_bounds__(256) cond_frames_prepare_kernel(const float* __restrict__ enc, const float* __restrict__ style,
                                                                  const int64_t* __restrict__ mel2ph, const float* __restrict__ mask,
                                                                  const float* __restrict__ f0, const float* __restrict__ uv,
                                                                  const float* __restrict__ pitch_table, float* __restrict__ dec32,
                                                                  TOp* __restrict__ pitch_inp, int rows, int T, int Tt, int C,
                                                                  int use_pitch, int use_uv, PitchConst pc) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / T;
  long long idx = mel2ph[row];
  idx = idx < 0 ? 0 : (idx > Tt ? Tt : idx);
  const bool nonpad = idx > 0;
  const float np = nonpad ? 1.f : 0.f;
  const float* er = enc + (static_cast<size_t>(b) * Tt + (nonpad ? idx - 1 : 0)) * C;
  const float* pe = nullptr;
  if (use_pitch) {
    const float one_m = mask ? __fsub_rn(1.f, mask[row]) : 1.f;
    const float mf0 = __fmul_rn(f0[row], one_m), muv = __fmul_rn(uv[row], one_m);
    const long long bin = f0_to_coarse_dev(denorm_f0_dev(mf0, use_uv != 0, muv, !nonpad), pc);
    pe = pitch_table + static_cast<size_t>(bin) * C;
  }
  for (int c = lane; c < C; c += 32) {
    const float e = nonpad ? er[c] : 0.f;
    dec32[static_cast<size_t>(row) * C + c] = e;
    if (use_pitch) {
      const float s = style ? __ldg(style + static_cast<size_t>(b) * C + c) : 0.f;
      store_op(pitch_inp + static_cast<size_t>(row) * C + c, __fadd_rn(__fmul_rn(__fadd_rn(e, s), np), __ldg(pe + c)));
    }
  }
}

// the rest of forward_pitch (fs.py:172-189) and the decoder input (fs.py:99-102):
//   use_pred_pitch: f0 = f0 (1-m) + pred_f0 m, uv = uv (1-m) + (pred_uv > 0) m, no padding mask; else the inputs as given
//   f0_denorm = denorm_f0(f0, uv, pad); pitch = f0_to_coarse(f0_denorm); f0_denorm_pred = denorm_f0(pred_f0, pred_uv > 0, pad)
//   decoder_inp = ((dec + pitch_embed(pitch)) + style) * (mel2ph > 0)
__global__ void __launch_bounds__(256) cond_frames_finish_kernel(const float* __restrict__ dec32, const float* __restrict__ style,
                                                                 const int64_t* __restrict__ mel2ph, const float* __restrict__ mask,
                                                                 const float* __restrict__ f0, const float* __restrict__ uv,
                                                                 const float* __restrict__ pitch_pred, const float* __restrict__ pitch_table,
                                                                 float* __restrict__ decoder_inp, float* __restrict__ f0_denorm,
                                                                 float* __restrict__ f0_denorm_pred, int64_t* __restrict__ pitch_out,
                                                                 int rows, int T, int C, int use_pitch, int use_uv, int use_pred_pitch,
                                                                 PitchConst pc) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / T;
  const bool nonpad = mel2ph[row] > 0;
  const float np = nonpad ? 1.f : 0.f;
  const float* pe = nullptr;
  if (use_pitch) {
    const float p0 = pitch_pred[static_cast<size_t>(row) * 2], p1 = pitch_pred[static_cast<size_t>(row) * 2 + 1];
    const float puv = p1 > 0.f ? 1.f : 0.f;
    float rf0 = f0[row], ruv = uv[row];
    bool pad = !nonpad;
    if (use_pred_pitch) {
      const float m = mask ? mask[row] : 0.f;
      const float one_m = __fsub_rn(1.f, m);
      rf0 = __fadd_rn(__fmul_rn(rf0, one_m), __fmul_rn(p0, m));
      ruv = __fadd_rn(__fmul_rn(ruv, one_m), __fmul_rn(puv, m));
      pad = false;                                         // pitch_padding = None (fs.py:174)
    }
    const float fd = denorm_f0_dev(rf0, use_uv != 0, ruv, pad);
    const long long bin = f0_to_coarse_dev(fd, pc);
    if (lane == 0) {
      f0_denorm[row] = fd;
      f0_denorm_pred[row] = denorm_f0_dev(p0, use_uv != 0, puv, pad);
      if (pitch_out) pitch_out[row] = bin;
    }
    pe = pitch_table + static_cast<size_t>(bin) * C;
  }
  for (int c = lane; c < C; c += 32) {
    float v = dec32[static_cast<size_t>(row) * C + c];
    if (use_pitch) v = __fadd_rn(v, __ldg(pe + c));
    const float s = style ? __ldg(style + static_cast<size_t>(b) * C + c) : 0.f;
    decoder_inp[static_cast<size_t>(row) * C + c] = __fmul_rn(__fadd_rn(v, s), np);
  }
}

}  // namespace
}  // namespace fse

using namespace fse;

struct fse_cond_encoder {
  fse_cond_encoder_config cfg{};
  LayerCtx ctx;
  bool loaded = false;
  float* embed_tokens = nullptr;
  ConvBlocksW enc;                      // TextConvEncoder's ConvBlocks
  float *spk_w = nullptr, *spk_b = nullptr;
  float* dur_embed = nullptr; int n_dur = 0;
  std::vector<ConvW> dur_conv; std::vector<LNW> dur_ln;
  float *dur_lin_w = nullptr, *dur_lin_b = nullptr;
  float* pitch_embed = nullptr; int n_pitch = 0;
  std::vector<ConvW> pitch_conv; std::vector<LNW> pitch_ln;
  float *pitch_lin_w = nullptr, *pitch_lin_b = nullptr;
  PitchConst pc{};
};

namespace {

// conv -> ReLU -> LayerNorm [-> * keep] stack of the two predictors; the last layer's output goes to y32 (fp32) for the head
template <typename TOp>
int predictor_stack(fse_cond_encoder* h, const std::vector<ConvW>& convs, const std::vector<LNW>& lns, const RowBufs& w, const float* keep,
                    int B, int T, cudaStream_t st) {
  const size_t rows = static_cast<size_t>(B) * T;
  const int L = static_cast<int>(convs.size());
  void* in = w.opA;
  void* out = w.opB;
  for (int i = 0; i < L; ++i) {
    EpiReluF32 epi{convs[i].bias, w.tmp32, convs[i].N, T};
    FSE_TRY((run_conv<TOp>(&h->ctx, convs[i], in, B, T, epi, st)));
    const bool last = i == L - 1;
    FSE_TRY((layer_norm<TOp>(&h->ctx, w.tmp32, lns[i], nullptr, keep, nullptr, last ? nullptr : out, last ? w.y32 : nullptr, rows, st)));
    void* t = in; in = out; out = t;
  }
  return FSE_OK;
}

template <typename TOp>
int text_encoder_impl(fse_cond_encoder* h, const int64_t* txt, float* enc_out, int B, int Tt, void* ws, cudaStream_t st) {
  const size_t rows = static_cast<size_t>(B) * Tt;
  const RowBufs w = carve_rows(h->ctx, ws, rows);
  const int H = h->cfg.hidden;
  cond_embed_tokens_kernel<<<row_blocks(rows), 256, 0, st>>>(txt, h->embed_tokens, w.x32, w.m0, static_cast<int>(rows), H, h->cfg.vocab,
                                                            sqrtf(static_cast<float>(H)));
  FSE_CUDA(cudaGetLastError());
  ++h->ctx.launches;
  EpiBiasMaskF32 e3{h->enc.post.bias, w.m0, enc_out, H, Tt};
  return conv_blocks_forward<TOp>(&h->ctx, h->enc, w, B, Tt, e3, st);
}

template <typename TOp>
int duration_impl(fse_cond_encoder* h, const float* dur_inp, const int64_t* masked_dur, const int64_t* txt, float* dur, int B, int Tt,
                  void* ws, cudaStream_t st) {
  const size_t rows = static_cast<size_t>(B) * Tt;
  const RowBufs w = carve_rows(h->ctx, ws, rows);
  cond_dur_embed_add_kernel<TOp><<<row_blocks(rows), 256, 0, st>>>(dur_inp, masked_dur, h->dur_embed, h->n_dur, txt,
                                                                  static_cast<TOp*>(w.opA), w.m0, static_cast<int>(rows), h->cfg.hidden);
  FSE_CUDA(cudaGetLastError());
  ++h->ctx.launches;
  FSE_TRY((predictor_stack<TOp>(h, h->dur_conv, h->dur_ln, w, w.m0, B, Tt, st)));
  cond_head_kernel<true><<<row_blocks(rows), 256, 0, st>>>(w.y32, h->dur_lin_w, h->dur_lin_b, w.m0, dur, static_cast<int>(rows),
                                                          h->cfg.hidden, 1);
  FSE_CUDA(cudaGetLastError());
  ++h->ctx.launches;
  return FSE_OK;
}

template <typename TOp>
int frames_impl(fse_cond_encoder* h, const float* enc, const float* style, const int64_t* mel2ph, const float* mask, const float* f0,
                const float* uv, int use_pred_pitch, float* decoder_inp, float* pitch_pred, float* f0_denorm, float* f0_denorm_pred,
                int64_t* pitch, int B, int Tt, int T, void* ws, cudaStream_t st) {
  const size_t rows = static_cast<size_t>(B) * T;
  const RowBufs w = carve_rows(h->ctx, ws, rows);
  const int H = h->cfg.hidden, up = h->cfg.use_pitch_embed, uu = h->cfg.use_uv;
  cond_frames_prepare_kernel<TOp><<<row_blocks(rows), 256, 0, st>>>(enc, style, mel2ph, mask, f0, uv, h->pitch_embed, w.x32,
                                                                   static_cast<TOp*>(w.opA), static_cast<int>(rows), T, Tt, H, up, uu, h->pc);
  FSE_CUDA(cudaGetLastError());
  ++h->ctx.launches;
  if (up) {
    FSE_TRY((predictor_stack<TOp>(h, h->pitch_conv, h->pitch_ln, w, nullptr, B, T, st)));
    cond_head_kernel<false><<<row_blocks(rows), 256, 0, st>>>(w.y32, h->pitch_lin_w, h->pitch_lin_b, nullptr, pitch_pred,
                                                             static_cast<int>(rows), H, 2);
    FSE_CUDA(cudaGetLastError());
    ++h->ctx.launches;
  }
  cond_frames_finish_kernel<<<row_blocks(rows), 256, 0, st>>>(w.x32, style, mel2ph, mask, f0, uv, pitch_pred, h->pitch_embed, decoder_inp,
                                                             f0_denorm, f0_denorm_pred, pitch, static_cast<int>(rows), T, H, up, uu,
                                                             use_pred_pitch, h->pc);
  FSE_CUDA(cudaGetLastError());
  ++h->ctx.launches;
  return FSE_OK;
}

int check_ws(const fse_cond_encoder* h, void* ws, int64_t ws_bytes, size_t rows) {
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 1023)) return fail(FSE_EINVAL, "workspace must be non-null and 1024-byte aligned");
  if (ws_bytes < static_cast<int64_t>(carve_rows(h->ctx, nullptr, rows).bytes)) return fail(FSE_EINVAL, "workspace too small");
  return FSE_OK;
}
int check_ready(const fse_cond_encoder* h) {
  if (!h) return fail(FSE_EINVAL, "null handle");
  if (!h->loaded) return fail(FSE_ESTATE, "weights not loaded");
  return FSE_OK;
}

}  // namespace

extern "C" {

int fse_cond_encoder_create(const fse_cond_encoder_config* cfg, fse_cond_encoder** out) {
  if (!cfg || !out) return fail(FSE_EINVAL, "null argument");
  if (cfg->mode < 0 || cfg->mode > 3) return fail(FSE_EINVAL, "unknown mode %d", cfg->mode);
  if (cfg->hidden <= 0 || cfg->hidden % 64 != 0 || cfg->hidden > 32 * kMaxPerLane)
    return fail(FSE_EINVAL, "hidden must be a multiple of 64, <= %d", 32 * kMaxPerLane);
  if (cfg->vocab <= 0) return fail(FSE_EINVAL, "vocab must be positive");
  if (cfg->enc_layers < 1 || cfg->enc_layers > 8 || cfg->layers_in_block < 1) return fail(FSE_EINVAL, "enc_layers must be 1..8, layers_in_block >= 1");
  for (int i = 0; i < cfg->enc_layers; ++i)
    if (cfg->enc_dilations[i] < 1) return fail(FSE_EINVAL, "enc_dilations[%d] must be >= 1", i);
  if (cfg->dur_predictor_layers < 1 || (cfg->use_pitch_embed && cfg->pitch_predictor_layers < 1)) return fail(FSE_EINVAL, "predictor layer counts must be >= 1");
  if (cfg->spk_embed_dim < 0) return fail(FSE_EINVAL, "spk_embed_dim must be >= 0");
  int dev = 0;
  FSE_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  FSE_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(FSE_ECUDA, "device is sm_%d%d; this library is built for sm_100a only (no fallback)", prop.major, prop.minor);
  auto* h = new fse_cond_encoder();
  h->cfg = *cfg;
  h->ctx.mode = cfg->mode;
  h->ctx.bf16 = mode_is_bf16(cfg->mode);
  h->ctx.hidden = cfg->hidden;
  const double mel_min = 1127.0 * std::log(1.0 + 50.0 / 700.0), mel_max = 1127.0 * std::log(1.0 + 900.0 / 700.0);
  h->pc.mel_min = static_cast<float>(mel_min);
  h->pc.mel_rng = static_cast<float>(mel_max - mel_min);
  *out = h;
  return FSE_OK;
}

void fse_cond_encoder_destroy(fse_cond_encoder* h) {
  if (!h) return;
  h->ctx.release();
  delete h;
}

int fse_cond_encoder_load_weights(fse_cond_encoder* h, const fse_tensor* tensors, int32_t n) {
  if (!h || !tensors || n <= 0) return fail(FSE_EINVAL, "null argument");
  if (h->loaded) return fail(FSE_ESTATE, "weights already loaded");
  TensorTable tt(tensors, n);
  const auto& c = h->cfg;
  const int H = c.hidden;
  LayerCtx* ctx = &h->ctx;
  FSE_TRY(load_vec(ctx, tt, "encoder.embed_tokens.weight", static_cast<int64_t>(c.vocab) * H, &h->embed_tokens));
  FSE_TRY(load_conv_blocks(ctx, tt, "encoder.", c.enc_layers, c.enc_dilations, c.layers_in_block, c.enc_kernel_size, c.enc_post_net_kernel, h->enc));
  if (c.spk_embed_dim > 0) {
    FSE_TRY(load_vec(ctx, tt, "spk_embed_proj.weight", static_cast<int64_t>(H) * c.spk_embed_dim, &h->spk_w));
    FSE_TRY(load_vec(ctx, tt, "spk_embed_proj.bias", H, &h->spk_b));
  }
  FSE_TRY(load_table(ctx, tt, "dur_embed.weight", H, &h->dur_embed, &h->n_dur));
  h->dur_conv.resize(c.dur_predictor_layers); h->dur_ln.resize(c.dur_predictor_layers);
  for (int i = 0; i < c.dur_predictor_layers; ++i) {
    const std::string pre = "dur_predictor.conv." + std::to_string(i) + ".";
    FSE_TRY(pack_conv(ctx, tt, pre + "0", H, H, c.dur_predictor_kernel, 1, h->dur_conv[i]));
    FSE_TRY(load_ln(ctx, tt, pre + "2", H, h->dur_ln[i]));
  }
  FSE_TRY(load_vec(ctx, tt, "dur_predictor.linear.0.weight", H, &h->dur_lin_w));
  FSE_TRY(load_vec(ctx, tt, "dur_predictor.linear.0.bias", 1, &h->dur_lin_b));
  if (c.use_pitch_embed) {
    FSE_TRY(load_table(ctx, tt, "pitch_embed.weight", H, &h->pitch_embed, &h->n_pitch));
    if (h->n_pitch < 256) return fail(FSE_EINVAL, "pitch_embed has %d rows; f0_to_coarse produces bins up to 255", h->n_pitch);
    h->pitch_conv.resize(c.pitch_predictor_layers); h->pitch_ln.resize(c.pitch_predictor_layers);
    for (int i = 0; i < c.pitch_predictor_layers; ++i) {
      const std::string pre = "pitch_predictor.conv." + std::to_string(i) + ".";
      FSE_TRY(pack_conv(ctx, tt, pre + "0", H, H, c.predictor_kernel, 1, h->pitch_conv[i]));
      FSE_TRY(load_ln(ctx, tt, pre + "2", H, h->pitch_ln[i]));
    }
    FSE_TRY(load_vec(ctx, tt, "pitch_predictor.linear.weight", 2 * static_cast<int64_t>(H), &h->pitch_lin_w));
    FSE_TRY(load_vec(ctx, tt, "pitch_predictor.linear.bias", 2, &h->pitch_lin_b));
  }
  h->loaded = true;
  return FSE_OK;
}

int64_t fse_cond_encoder_workspace_bytes(const fse_cond_encoder* h, int32_t B, int32_t Tt, int32_t T) {
  if (!h || B <= 0 || (Tt <= 0 && T <= 0)) return 0;
  const size_t rows = static_cast<size_t>(B) * static_cast<size_t>(Tt > T ? Tt : T);
  return static_cast<int64_t>(carve_rows(h->ctx, nullptr, rows).bytes);
}

int64_t fse_cond_encoder_last_launches(const fse_cond_encoder* h) { return h ? h->ctx.launches : 0; }

int fse_cond_text_encoder(fse_cond_encoder* h, const int64_t* txt, float* encoder_out, int32_t B, int32_t Tt, void* workspace,
                          int64_t workspace_bytes, void* stream) {
  FSE_TRY(check_ready(h));
  if (!txt || !encoder_out) return fail(FSE_EINVAL, "null argument");
  if (B <= 0 || Tt <= 0) return fail(FSE_EINVAL, "B and Tt must be positive");
  FSE_TRY(check_ws(h, workspace, workspace_bytes, static_cast<size_t>(B) * Tt));
  h->ctx.launches = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return h->ctx.bf16 ? text_encoder_impl<__nv_bfloat16>(h, txt, encoder_out, B, Tt, workspace, st)
                 : text_encoder_impl<float>(h, txt, encoder_out, B, Tt, workspace, st);
}

int fse_cond_style_embed(fse_cond_encoder* h, const float* spk_embed, float* style, int32_t B, void* stream) {
  FSE_TRY(check_ready(h));
  if (!spk_embed || !style) return fail(FSE_EINVAL, "null argument");
  if (B <= 0) return fail(FSE_EINVAL, "B must be positive");
  if (h->cfg.spk_embed_dim <= 0) return fail(FSE_ESTATE, "handle was created without a speaker embedding (spk_embed_dim = 0)");
  h->ctx.launches = 1;
  cond_style_kernel<<<B, 192, 0, static_cast<cudaStream_t>(stream)>>>(spk_embed, h->spk_w, h->spk_b, style, h->cfg.spk_embed_dim, h->cfg.hidden);
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

int fse_cond_dur_input(fse_cond_encoder* h, const float* encoder_out, const float* style, const int64_t* txt, float* dur_inp, int32_t B,
                       int32_t Tt, void* stream) {
  FSE_TRY(check_ready(h));
  if (!encoder_out || !txt || !dur_inp) return fail(FSE_EINVAL, "null argument");
  if (B <= 0 || Tt <= 0) return fail(FSE_EINVAL, "B and Tt must be positive");
  const size_t rows = static_cast<size_t>(B) * Tt;
  h->ctx.launches = 1;
  cond_dur_input_kernel<<<row_blocks(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(encoder_out, style, txt, dur_inp,
                                                                                         static_cast<int>(rows), Tt, h->cfg.hidden);
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

int fse_cond_masked_dur(fse_cond_encoder* h, const int64_t* mel2ph, const float* mask, const int64_t* txt, int64_t* masked_dur, int32_t B,
                        int32_t T, int32_t Tt, void* stream) {
  FSE_TRY(check_ready(h));
  if (!mel2ph || !txt || !masked_dur) return fail(FSE_EINVAL, "null argument");
  if (B <= 0 || T <= 0 || Tt <= 0) return fail(FSE_EINVAL, "B, T and Tt must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FSE_CUDA(cudaMemsetAsync(masked_dur, 0, static_cast<size_t>(B) * Tt * sizeof(int64_t), st));
  const size_t n = static_cast<size_t>(B) * T;
  h->ctx.launches = 1;
  cond_masked_dur_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(mel2ph, mask, txt, masked_dur, B, T, Tt);
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

int fse_cond_duration(fse_cond_encoder* h, const float* dur_inp, const int64_t* masked_dur, const int64_t* txt, float* dur, int32_t B,
                      int32_t Tt, void* workspace, int64_t workspace_bytes, void* stream) {
  FSE_TRY(check_ready(h));
  if (!dur_inp || !masked_dur || !txt || !dur) return fail(FSE_EINVAL, "null argument");
  if (B <= 0 || Tt <= 0) return fail(FSE_EINVAL, "B and Tt must be positive");
  FSE_TRY(check_ws(h, workspace, workspace_bytes, static_cast<size_t>(B) * Tt));
  h->ctx.launches = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return h->ctx.bf16 ? duration_impl<__nv_bfloat16>(h, dur_inp, masked_dur, txt, dur, B, Tt, workspace, st)
                 : duration_impl<float>(h, dur_inp, masked_dur, txt, dur, B, Tt, workspace, st);
}

int fse_cond_length_cumsum(fse_cond_encoder* h, const float* dur, const int64_t* txt, int64_t* cumsum, int64_t* totals, int32_t B,
                           int32_t Tt, void* stream) {
  if (!h) return fail(FSE_EINVAL, "null handle");
  if (!dur || !cumsum || !totals) return fail(FSE_EINVAL, "null argument");
  if (B <= 0 || Tt <= 0) return fail(FSE_EINVAL, "B and Tt must be positive");
  h->ctx.launches = 1;
  cond_length_cumsum_kernel<<<(B + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(dur, txt, cumsum, totals, B, Tt);
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

int fse_cond_length_fill(fse_cond_encoder* h, const int64_t* cumsum, int64_t* mel2ph, int32_t B, int32_t Tt, int32_t Tmax, void* stream) {
  if (!h) return fail(FSE_EINVAL, "null handle");
  if (!cumsum || !mel2ph) return fail(FSE_EINVAL, "null argument");
  if (B <= 0 || Tt <= 0 || Tmax <= 0) return fail(FSE_EINVAL, "B, Tt and Tmax must be positive");
  const size_t n = static_cast<size_t>(B) * Tmax;
  h->ctx.launches = 1;
  cond_length_fill_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(cumsum, mel2ph, B, Tt, Tmax);
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

int fse_cond_frames(fse_cond_encoder* h, const float* encoder_out, const float* style, const int64_t* mel2ph, const float* mask,
                    const float* f0, const float* uv, int32_t use_pred_pitch, float* decoder_inp, float* pitch_pred, float* f0_denorm,
                    float* f0_denorm_pred, int64_t* pitch, int32_t B, int32_t Tt, int32_t T, void* workspace, int64_t workspace_bytes,
                    void* stream) {
  FSE_TRY(check_ready(h));
  if (!encoder_out || !mel2ph || !decoder_inp) return fail(FSE_EINVAL, "null argument");
  if (B <= 0 || Tt <= 0 || T <= 0) return fail(FSE_EINVAL, "B, Tt and T must be positive");
  if (h->cfg.use_pitch_embed && (!f0 || !uv || !pitch_pred || !f0_denorm || !f0_denorm_pred))
    return fail(FSE_EINVAL, "use_pitch_embed: f0, uv, pitch_pred, f0_denorm and f0_denorm_pred are required");
  FSE_TRY(check_ws(h, workspace, workspace_bytes, static_cast<size_t>(B) * T));
  h->ctx.launches = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return h->ctx.bf16 ? frames_impl<__nv_bfloat16>(h, encoder_out, style, mel2ph, mask, f0, uv, use_pred_pitch, decoder_inp, pitch_pred, f0_denorm,
                                              f0_denorm_pred, pitch, B, Tt, T, workspace, st)
                 : frames_impl<float>(h, encoder_out, style, mel2ph, mask, f0, uv, use_pred_pitch, decoder_inp, pitch_pred, f0_denorm,
                                      f0_denorm_pred, pitch, B, Tt, T, workspace, st);
}

}  // extern "C"
