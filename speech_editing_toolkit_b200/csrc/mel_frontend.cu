// Mel front-end of the inference entry point: wav -> log10-mel (utils/audio/__init__.py:34-81 `librosa_wav2spec`, called at
// inference/tts/spec_denoiser.py:258) behind the C ABI of include/fse_b200.h.
//   stft (n_fft = R * hop, periodic Hann, center = True with zero padding) -> |.| -> Slaney mel filterbank -> log10(max(eps, .))
// The waveform [B, n] is viewed as rows of `hop` samples ([B, n/hop, hop], no copy): the window of frame t covers rows
// t - R/2 .. t + R/2 - 1, i.e. the windowed DFT is an R-tap conv-GEMM over that row sequence (weights = window x cos / sin,
// real and imaginary parts of a bin in adjacent columns so that the epilogue forms the magnitude), rows outside the signal
// read zeros — exactly librosa's center padding.  The mel projection is a second GEMM with the log10 in its epilogue.
// Both run the library's fp32 CUDA-core GEMM (conv_gemm_simt_kernel<float>): log-mel of quiet frames needs fp32 operands
// (bf16 tensor-core operands lose the low-energy bins), and the work is ~2 MFLOP per frame, once per utterance.
#include <cmath>

#include "mel_frontend_weights.h"
#include "rowwise.cuh"

namespace fse {
namespace {

// |re + i im| of adjacent column pairs -> mag[row, n/2]
struct EpiMagnitude {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  float* out;   // [B*T, N/2]
  int N, T;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    float* dst = out + (static_cast<size_t>(b) * T + t) * (N / 2) + n0 / 2;
#pragma unroll
    for (int i = 0; i < NV / 2; ++i) dst[i] = sqrtf(fmaf(acc[2 * i], acc[2 * i], acc[2 * i + 1] * acc[2 * i + 1]));
  }
};

// mel = log10(max(eps, basis @ mag))
struct EpiLog10 {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  float* out;   // [B*T, N]
  int N, T;
  float eps;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = log10f(fmaxf(eps, acc[i]));
    st_vec<NV>(out + (static_cast<size_t>(b) * T + t) * N + n0, v);
  }
};

}  // namespace
}  // namespace fse

using namespace fse;

struct fse_mel_frontend {
  fse_mel_frontend_config cfg{};
  LayerCtx ctx;
  ConvW dft, mel;
  int taps = 0, nbins = 0, ndft = 0;      // R, n_fft/2+1, padded 2*nbins (columns of the DFT GEMM)
  long long launches = 0;
};

extern "C" {

int fse_mel_frontend_create(const fse_mel_frontend_config* cfg, fse_mel_frontend** out) {
  if (!cfg || !out) return fail(FSE_EINVAL, "null argument");
  if (cfg->hop_size <= 0 || cfg->hop_size % 64 != 0) return fail(FSE_EINVAL, "hop_size must be a positive multiple of 64");
  if (cfg->fft_size <= 0 || cfg->fft_size % cfg->hop_size != 0 || (cfg->fft_size / cfg->hop_size) % 2 != 0 || cfg->fft_size / cfg->hop_size > kMaxTaps)
    return fail(FSE_EINVAL, "fft_size must be an even multiple (<= %d) of hop_size", kMaxTaps);
  if (cfg->win_length != cfg->fft_size) return fail(FSE_EINVAL, "win_length must equal fft_size (the reference's setting)");
  if (cfg->num_mels <= 0 || cfg->num_mels % 16 != 0) return fail(FSE_EINVAL, "num_mels must be a positive multiple of 16");
  if (cfg->sample_rate <= 0) return fail(FSE_EINVAL, "sample_rate must be positive");
  const double fmin = cfg->fmin < 0 ? 0.0 : cfg->fmin, fmax = cfg->fmax < 0 ? cfg->sample_rate / 2.0 : cfg->fmax;    // -1 conventions of :63-64
  if (!(fmin < fmax) || fmax > cfg->sample_rate / 2.0 + 1e-6) return fail(FSE_EINVAL, "need 0 <= fmin < fmax <= sample_rate / 2");
  if (!(cfg->eps > 0)) return fail(FSE_EINVAL, "eps must be positive");
  int dev = 0;
  FSE_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  FSE_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(FSE_ECUDA, "device is sm_%d%d; this library is built for sm_100a only (no fallback)", prop.major, prop.minor);
  auto* h = new fse_mel_frontend();
  h->cfg = *cfg;
  h->ctx.mode = FSE_MODE_SIMT_F32;
  h->ctx.bf16 = false;
  h->ctx.hidden = 0;
  const int n_fft = cfg->fft_size, hop = cfg->hop_size, R = n_fft / hop, nbins = n_fft / 2 + 1;
  const int ndft = (2 * nbins + 63) / 64 * 64;
  h->taps = R; h->nbins = nbins; h->ndft = ndft;
  {   // windowed DFT as conv weights (mel_frontend_weights.h); tap r reads row t + r - R/2
    std::vector<float> w;
    melfe::build_dft_weights(n_fft, hop, ndft, w);
    int offs[kMaxTaps];
    for (int r = 0; r < R; ++r) offs[r] = r - R / 2;
    const int rc = pack_conv_raw(&h->ctx, "stft", w.data(), nullptr, ndft, hop, R, offs, h->dft);
    if (rc != FSE_OK) { h->ctx.release(); delete h; return rc; }
  }
  {   // librosa.filters.mel (htk=False, norm='slaney') over the padded magnitude row [ndft / 2]
    std::vector<float> w;
    melfe::build_mel_weights(cfg->sample_rate, n_fft, cfg->num_mels, fmin, fmax, ndft / 2, w);
    const int zero = 0;
    const int rc = pack_conv_raw(&h->ctx, "mel_basis", w.data(), nullptr, cfg->num_mels, ndft / 2, 1, &zero, h->mel);
    if (rc != FSE_OK) { h->ctx.release(); delete h; return rc; }
  }
  *out = h;
  return FSE_OK;
}

void fse_mel_frontend_destroy(fse_mel_frontend* h) {
  if (!h) return;
  h->ctx.release();
  delete h;
}

int64_t fse_mel_frontend_frames(const fse_mel_frontend* h, int64_t n_samples) {
  if (!h || n_samples < 0) return 0;
  return 1 + n_samples / h->cfg.hop_size;
}

int64_t fse_mel_frontend_workspace_bytes(const fse_mel_frontend* h, int32_t B, int64_t n_samples) {
  if (!h || B <= 0 || n_samples <= 0) return 0;
  return static_cast<int64_t>(align_up(static_cast<size_t>(B) * fse_mel_frontend_frames(h, n_samples) * (h->ndft / 2) * sizeof(float), 1024));
}

int fse_mel_frontend_forward(fse_mel_frontend* h, const float* wav, float* mel, int32_t B, int64_t n_samples, void* workspace,
                             int64_t workspace_bytes, void* stream) {
  if (!h || !wav || !mel) return fail(FSE_EINVAL, "null argument");
  const int hop = h->cfg.hop_size;
  if (B <= 0 || n_samples <= 0 || n_samples % hop != 0) return fail(FSE_EINVAL, "n_samples must be a positive multiple of hop_size (pad with zeros)");
  if (n_samples / hop > (1 << 24)) return fail(FSE_EINVAL, "utterance too long");
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 1023)) return fail(FSE_EINVAL, "workspace must be non-null and 1024-byte aligned");
  if (workspace_bytes < fse_mel_frontend_workspace_bytes(h, B, n_samples)) return fail(FSE_EINVAL, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Tc = static_cast<int>(n_samples / hop), Tf = Tc + 1;
  float* mag = static_cast<float*>(workspace);
  h->ctx.launches = 0;
  {   // windowed DFT + magnitude: rows = frames, source rows = hop-sample chunks, taps at -R/2 .. R/2-1
    ConvGemmParams p = make_params(B, Tf, Tc, h->dft.Cin, h->dft.ntaps, h->dft.offs, 0, h->dft.N, h->dft.KB);
    GemmOperands op; op.A0 = wav; op.W = h->dft.W; op.BN = h->dft.BN;
    EpiMagnitude epi{mag, h->ndft, Tf};
    FSE_TRY((run_conv_gemm<float>(FSE_MODE_SIMT_F32, p, op, epi, st, LaunchCtx{&h->ctx.launches, nullptr, 0})));
  }
  {
    ConvGemmParams p = make_params(B, Tf, Tf, h->mel.Cin, 1, h->mel.offs, 0, h->mel.N, h->mel.KB);
    GemmOperands op; op.A0 = mag; op.W = h->mel.W; op.BN = h->mel.BN;
    EpiLog10 epi{mel, h->cfg.num_mels, Tf, h->cfg.eps};
    FSE_TRY((run_conv_gemm<float>(FSE_MODE_SIMT_F32, p, op, epi, st, LaunchCtx{&h->ctx.launches, nullptr, 0})));
  }
  h->launches = h->ctx.launches;
  return FSE_OK;
}

int64_t fse_mel_frontend_last_launches(const fse_mel_frontend* h) { return h ? h->launches : 0; }

}  // extern "C"
