// Mel losses of the FluentSpeech training step behind the C ABI of include/fse_b200.h: fse_mel_loss_forward / fse_mel_loss_backward.
//
// Reference: SpeechBaseTask.add_mel_loss with the shipped `mel_losses: l1:0.5|ssim:0.5` (tasks/tts/speech_base.py:219-257):
//   w[b,t]  = (sum_c |target[b,t,c]| != 0)                                   weights_nonzero_speech, repeated over the bins
//   l1      = sum(|out - target| * w) / sum(w)
//   ssim    = sum((1 - SSIM(out + 6, target + 6)) * w) / sum(w)               utils/metrics/ssim.py:24-44: 11 x 11 Gaussian window
//                                                                             (sigma 1.5), zero padding 5, C1 = 1e-4, C2 = 9e-4
// and their gradient with respect to `out` (what torch.autograd derives from those lines).
//
// An utterance is an image of T rows x M bins.  HBM-bound CUDA-core work (2 reads + 3 writes of the image forward, 5 reads + 1 write
// backward); the 11 x 11 window is applied directly from a shared-memory tile (121 taps, the reference's own summation domain), one
// thread per pixel.  Everything is deterministic: sum(w) is an integer count, the loss sums go block partial -> one fixed-order
// final reduction, no floating-point atomics.
//
// Backward algebra.  With a = out + 6, b = target + 6, mu1 = W*a, mu2 = W*b, Eaa = W*(a a), Ebb = W*(b b), Eab = W*(a b):
//   A1 = 2 mu1 mu2 + C1, A2 = 2 (Eab - mu1 mu2) + C2, B1 = mu1^2 + mu2^2 + C1, B2 = Eaa - mu1^2 + Ebb - mu2^2 + C2, S = A1 A2 / (B1 B2)
//   Gmu = dS/dmu1 = 2 mu2 (A2 - A1) / (B1 B2) - 2 mu1 S (1/B1 - 1/B2),   Gaa = dS/dEaa = -S / B2,   Gab = dS/dEab = 2 A1 / (B1 B2)
// and, the window being symmetric,  dS_total/da(q) = (W*Gmu)(q) + 2 a(q) (W*Gaa)(q) + b(q) (W*Gab)(q).
// The forward stores the three fields pre-multiplied by -w(p) / sum(w) (the derivative of the weighted mean of 1 - S).
#include <cuda_runtime.h>

#include <cmath>

#include "fse_common.cuh"

namespace fse {
namespace {

constexpr int kWin = 11, kPad = 5, kRows = 16, kThreads = 256, kMaxBins = 128;
constexpr float kBias = 6.0f, kC1 = 1.0e-4f, kC2 = 9.0e-4f;          // float(0.01 ** 2), float(0.03 ** 2)

struct Window { float w[kWin * kWin]; };

// gaussian() / create_window() of utils/metrics/ssim.py:12-22 in the same arithmetic: the 11 exponentials are Python doubles stored
// into a float tensor, normalised in fp32, and the 2-D window is the fp32 outer product.
Window make_window() {
  float g[kWin], s = 0.f;
  for (int x = 0; x < kWin; ++x) { g[x] = static_cast<float>(std::exp(-static_cast<double>((x - kWin / 2) * (x - kWin / 2)) / (2.0 * 1.5 * 1.5))); s += g[x]; }
  for (int x = 0; x < kWin; ++x) g[x] = g[x] / s;
  Window w;
  for (int i = 0; i < kWin; ++i)
    for (int j = 0; j < kWin; ++j) w.w[i * kWin + j] = g[i] * g[j];
  return w;
}

// Workspace: [0] frame count with w = 1 (u64), [16 ..) per-block partial sums (float2: l1, ssim), then w[B*T] and the three G fields.
struct Layout {
  size_t partials, w, g, total;
  int blocks;
};
Layout layout(int B, int T, int M) {
  Layout L;
  L.blocks = B * ((T + kRows - 1) / kRows);
  L.partials = 16;
  L.w = align_up(L.partials + static_cast<size_t>(L.blocks) * 8, 256);
  L.g = align_up(L.w + static_cast<size_t>(B) * T * 4, 256);
  L.total = L.g + 3 * static_cast<size_t>(B) * T * M * 4;
  return L;
}

__global__ void __launch_bounds__(kThreads) frame_weight_kernel(const float* __restrict__ target, float* __restrict__ w, unsigned long long* __restrict__ count,
                                                                size_t frames, int M) {
  const size_t row = static_cast<size_t>(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  if (row < frames)
    for (int c = lane; c < M; c += 32) s += fabsf(target[row * M + c]);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const bool on = row < frames && s != 0.f;
  __shared__ int block_count;
  if (threadIdx.x == 0) block_count = 0;
  __syncthreads();
  if (lane == 0 && row < frames) {
    w[row] = on ? 1.f : 0.f;
    if (on) atomicAdd(&block_count, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0 && block_count) atomicAdd(count, static_cast<unsigned long long>(block_count));
}

__device__ __forceinline__ float block_sum(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x == 0)
    for (int i = 0; i < kThreads / 32; ++i) r += scratch[i];
  __syncthreads();
  return r;
}

// One block = kRows frames x M bins of one utterance.
__global__ void __launch_bounds__(kThreads) mel_loss_forward_kernel(const float* __restrict__ out, const float* __restrict__ target, const float* __restrict__ w,
                                                                    const unsigned long long* __restrict__ count, Window win, float2* __restrict__ partials,
                                                                    float* __restrict__ G, int T, int M, int want_grad) {
  extern __shared__ float smem[];
  const int tiles = (T + kRows - 1) / kRows;
  const int b = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * kRows;
  const int SW = M + 2 * kPad, SH = kRows + 2 * kPad;
  float* sa = smem;                       // [SH][SW]  out + 6 inside the image, 0 outside (conv2d zero padding)
  float* sb = smem + SH * SW;
  __shared__ float swin[kWin * kWin];
  __shared__ float scratch[kThreads / 32];
  for (int i = threadIdx.x; i < kWin * kWin; i += kThreads) swin[i] = win.w[i];
  const size_t item = static_cast<size_t>(b) * T;
  for (int i = threadIdx.x; i < SH * SW; i += kThreads) {
    const int r = i / SW, c = i % SW;
    const int t = t0 + r - kPad, m = c - kPad;
    float a = 0.f, bb = 0.f;
    if (t >= 0 && t < T && m >= 0 && m < M) {
      a = out[(item + t) * M + m] + kBias;
      bb = target[(item + t) * M + m] + kBias;
    }
    sa[i] = a;
    sb[i] = bb;
  }
  __syncthreads();
  const float inv_wsum = 1.f / (static_cast<float>(*count) * static_cast<float>(M));
  const size_t plane = static_cast<size_t>(gridDim.x / tiles) * T * M;
  float acc_l1 = 0.f, acc_ssim = 0.f;
  for (int i = threadIdx.x; i < kRows * M; i += kThreads) {
    const int r = i / M, m = i % M, t = t0 + r;
    if (t >= T) continue;
    float mu1 = 0.f, mu2 = 0.f, eaa = 0.f, ebb = 0.f, eab = 0.f;
#pragma unroll 1
    for (int dy = 0; dy < kWin; ++dy) {
      const float* ra = sa + (r + dy) * SW + m;
      const float* rb = sb + (r + dy) * SW + m;
#pragma unroll
      for (int dx = 0; dx < kWin; ++dx) {
        const float k = swin[dy * kWin + dx], a = ra[dx], bb = rb[dx];
        mu1 = fmaf(k, a, mu1);
        mu2 = fmaf(k, bb, mu2);
        eaa = fmaf(k, a * a, eaa);
        ebb = fmaf(k, bb * bb, ebb);
        eab = fmaf(k, a * bb, eab);
      }
    }
    const float A1 = 2.f * mu1 * mu2 + kC1, A2 = 2.f * (eab - mu1 * mu2) + kC2;
    const float B1 = mu1 * mu1 + mu2 * mu2 + kC1, B2 = (eaa - mu1 * mu1) + (ebb - mu2 * mu2) + kC2;
    const float inv = 1.f / (B1 * B2), S = A1 * A2 * inv;
    const float wt = w[item + t];
    const size_t o = (item + t) * M + m;
    acc_l1 += fabsf(out[o] - target[o]) * wt;
    acc_ssim += (1.f - S) * wt;
    if (want_grad) {
      const float up = -wt * inv_wsum;
      G[o] = up * (2.f * mu2 * (A2 - A1) * inv - 2.f * mu1 * S * (1.f / B1 - 1.f / B2));
      G[plane + o] = up * (-S / B2);
      G[2 * plane + o] = up * (2.f * A1 * inv);
    }
  }
  const float s1 = block_sum(acc_l1, scratch), s2 = block_sum(acc_ssim, scratch);
  if (threadIdx.x == 0) partials[blockIdx.x] = make_float2(s1, s2);
}

__global__ void __launch_bounds__(kThreads) mel_loss_final_kernel(const float2* __restrict__ partials, int n, const unsigned long long* __restrict__ count, int M,
                                                                  float lambda_l1, float lambda_ssim, float* __restrict__ losses) {
  __shared__ double s1[kThreads], s2[kThreads];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += kThreads) { a += partials[i].x; b += partials[i].y; }
  s1[threadIdx.x] = a;
  s2[threadIdx.x] = b;
  __syncthreads();
  for (int o = kThreads / 2; o; o >>= 1) {
    if (threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double wsum = static_cast<double>(*count) * M;           // 0 frames with speech: 0 / 0 = NaN, as in the reference
    losses[0] = static_cast<float>(s1[0] / wsum) * lambda_l1;
    losses[1] = static_cast<float>(s2[0] / wsum) * lambda_ssim;
  }
}

__global__ void __launch_bounds__(kThreads) mel_loss_backward_kernel(const float* __restrict__ out, const float* __restrict__ target, const float* __restrict__ w,
                                                                     const unsigned long long* __restrict__ count, Window win, const float* __restrict__ G,
                                                                     const float* __restrict__ dlosses, float lambda_l1, float lambda_ssim,
                                                                     float* __restrict__ grad, int T, int M) {
  extern __shared__ float smem[];
  const int tiles = (T + kRows - 1) / kRows;
  const int b = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * kRows;
  const int SW = M + 2 * kPad, SH = kRows + 2 * kPad;
  float* g0 = smem;
  float* g1 = smem + SH * SW;
  float* g2 = smem + 2 * SH * SW;
  __shared__ float swin[kWin * kWin];
  for (int i = threadIdx.x; i < kWin * kWin; i += kThreads) swin[i] = win.w[i];
  const size_t item = static_cast<size_t>(b) * T;
  const size_t plane = static_cast<size_t>(gridDim.x / tiles) * T * M;
  for (int i = threadIdx.x; i < SH * SW; i += kThreads) {
    const int r = i / SW, c = i % SW;
    const int t = t0 + r - kPad, m = c - kPad;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    if (t >= 0 && t < T && m >= 0 && m < M) {
      const size_t o = (item + t) * M + m;
      x0 = G[o];
      x1 = G[plane + o];
      x2 = G[2 * plane + o];
    }
    g0[i] = x0;
    g1[i] = x1;
    g2[i] = x2;
  }
  __syncthreads();
  const float up_l1 = dlosses[0] * lambda_l1 / (static_cast<float>(*count) * static_cast<float>(M));
  const float up_ssim = dlosses[1] * lambda_ssim;
  for (int i = threadIdx.x; i < kRows * M; i += kThreads) {
    const int r = i / M, m = i % M, t = t0 + r;
    if (t >= T) continue;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll 1
    for (int dy = 0; dy < kWin; ++dy) {
      const int base = (r + dy) * SW + m;
#pragma unroll
      for (int dx = 0; dx < kWin; ++dx) {
        const float k = swin[dy * kWin + dx];
        c0 = fmaf(k, g0[base + dx], c0);
        c1 = fmaf(k, g1[base + dx], c1);
        c2 = fmaf(k, g2[base + dx], c2);
      }
    }
    const size_t o = (item + t) * M + m;
    const float x = out[o], y = target[o];
    const float a = x + kBias, bb = y + kBias;
    const float diff = x - y;
    const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
    grad[o] = up_ssim * (c0 + 2.f * a * c1 + bb * c2) + up_l1 * sgn * w[item + t];
  }
}

int check_args(const void* a, const void* b, const void* c, int B, int T, int M, const void* ws, int64_t ws_bytes) {
  if (!a || !b || !c || !ws) return fail(FSE_EINVAL, "null argument");
  if (B <= 0 || T <= 0 || M <= 0 || M > kMaxBins) return fail(FSE_EINVAL, "B and T must be positive and 1 <= n_mels <= %d", kMaxBins);
  if (ws_bytes < static_cast<int64_t>(layout(B, T, M).total)) return fail(FSE_EINVAL, "workspace too small: %lld < %lld bytes",
                                                                           static_cast<long long>(ws_bytes), static_cast<long long>(layout(B, T, M).total));
  int dev = 0;
  FSE_CUDA(cudaGetDevice(&dev));
  int major = 0;
  FSE_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(FSE_ECUDA, "this library is built for sm_100a only (no fallback)");
  return FSE_OK;
}

}  // namespace
}  // namespace fse

using namespace fse;

extern "C" {

int64_t fse_mel_loss_workspace_bytes(int32_t B, int32_t T, int32_t n_mels) {
  if (B <= 0 || T <= 0 || n_mels <= 0) return 0;
  return static_cast<int64_t>(layout(B, T, n_mels).total);
}

int fse_mel_loss_forward(const float* mel_out, const float* target, float lambda_l1, float lambda_ssim, float* losses, int32_t want_grad, int32_t B, int32_t T,
                         int32_t n_mels, void* workspace, int64_t workspace_bytes, void* stream) {
  FSE_TRY(check_args(mel_out, target, losses, B, T, n_mels, workspace, workspace_bytes));
  const Layout L = layout(B, T, n_mels);
  auto st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  auto* count = reinterpret_cast<unsigned long long*>(ws);
  auto* partials = reinterpret_cast<float2*>(ws + L.partials);
  auto* w = reinterpret_cast<float*>(ws + L.w);
  auto* G = reinterpret_cast<float*>(ws + L.g);
  FSE_CUDA(cudaMemsetAsync(count, 0, 16, st));
  const size_t frames = static_cast<size_t>(B) * T;
  frame_weight_kernel<<<static_cast<unsigned>((frames + kThreads / 32 - 1) / (kThreads / 32)), kThreads, 0, st>>>(target, w, count, frames, n_mels);
  const size_t smem = 2 * static_cast<size_t>(kRows + 2 * kPad) * (n_mels + 2 * kPad) * sizeof(float);
  mel_loss_forward_kernel<<<L.blocks, kThreads, smem, st>>>(mel_out, target, w, count, make_window(), partials, G, T, n_mels, want_grad);
  mel_loss_final_kernel<<<1, kThreads, 0, st>>>(partials, L.blocks, count, n_mels, lambda_l1, lambda_ssim, losses);
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

int fse_mel_loss_backward(const float* mel_out, const float* target, const float* dlosses, float lambda_l1, float lambda_ssim, float* grad, int32_t B, int32_t T,
                          int32_t n_mels, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!dlosses) return fail(FSE_EINVAL, "null argument");
  FSE_TRY(check_args(mel_out, target, grad, B, T, n_mels, workspace, workspace_bytes));
  const Layout L = layout(B, T, n_mels);
  auto st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  const size_t smem = 3 * static_cast<size_t>(kRows + 2 * kPad) * (n_mels + 2 * kPad) * sizeof(float);
  mel_loss_backward_kernel<<<L.blocks, kThreads, smem, st>>>(mel_out, target, reinterpret_cast<const float*>(ws + L.w), reinterpret_cast<const unsigned long long*>(ws),
                                                            make_window(), reinterpret_cast<const float*>(ws + L.g), dlosses, lambda_l1, lambda_ssim, grad, T, n_mels);
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

}  // extern "C"
