// MelEncoder forward (modules/speech_editing/commons/mel_encoder.py:3-19) behind the C ABI of include/fse_b200.h:
//   out = fc_out( relu( W1 relu( W0 x + b0 ) + b1 ) ) + b2,   x = ref_mels * (1 - time_mel_masks)   [B, T, n_mels]
// optionally fused with its only call site (spec_denoiser.py:162-164):  cond = decoder_inp + out * tgt_nonpadding.
// Three launches of the conv-as-GEMM primitive (one tap): frames are GEMM rows, the two hidden activations stay bf16
// operand tiles (fp32 in FSE_MODE_SIMT_F32), the last epilogue writes fp32.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "fse_common.cuh"

namespace fse {
namespace {

// hidden layer: out = relu(acc + bias) in the operand type
template <typename TOp>
struct EpiRelu {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  const float* bias;
  TOp* out;   // [B*T, N]
  int N, T;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = fmaxf(acc[i] + __ldg(bias + n0 + i), 0.f);
    st_vec<NV>(out + (static_cast<size_t>(b) * T + t) * N + n0, v);
  }
};

// fc_out: y = acc + bias;  out = add + y * scale[row]   (add / scale optional; separate roundings as in torch)
struct EpiCond {
  static constexpr int kAux = 1;
  static constexpr bool kTransposed = true;
  const float* bias;
  const float* add;     // [B*T, N] fp32 or null
  const float* scale;   // [B*T] fp32 or null
  float* out;           // [B*T, N] fp32
  int N, T;
  template <int NV>
  __device__ __forceinline__ void load_aux(int b, int t, int n0, float* aux) const {
    if (add) {
      const size_t o = (static_cast<size_t>(b) * T + t) * N + n0;
#pragma unroll
      for (int i = 0; i < NV / 4; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(add + o) + i);
        aux[4 * i] = v.x; aux[4 * i + 1] = v.y; aux[4 * i + 2] = v.z; aux[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) aux[i] = 0.f;
    }
  }
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float* aux) const {
    const size_t row = static_cast<size_t>(b) * T + t;
    const float s = scale ? __ldg(scale + row) : 1.0f;
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float y = acc[i] + __ldg(bias + n0 + i);
      if (scale) y = __fmul_rn(y, s);
      v[i] = add ? __fadd_rn(aux[i], y) : y;
    }
    st_vec<NV>(out + row * N + n0, v);
  }
};

__global__ void __launch_bounds__(256) mel_f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n4) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n4) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

struct Layer {
  void* W = nullptr; float* bias = nullptr; CUtensorMap map{};
  int Cin = 0, N = 0, Kp = 0, BN = 0;
};

}  // namespace
}  // namespace fse

using namespace fse;

struct fse_mel_encoder {
  fse_mel_encoder_config cfg{};
  bool bf16 = true, loaded = false;
  Layer l[3];
  struct Plan { const void* ws = nullptr; const void* x = nullptr; int B = 0, T = 0; CUtensorMap mx{}, m1{}, m2{}; } plan;
  long long launches = 0;
};

namespace {

struct MWs { void* xb; void* a1; void* a2; size_t bytes; };
MWs mcarve(const fse_mel_encoder* h, void* base, int B, int T) {
  const size_t N = static_cast<size_t>(B) * T, es = h->bf16 ? 2 : 4;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += align_up(b, 1024); return o; };
  uint8_t* p = static_cast<uint8_t*>(base);
  MWs w{};
  size_t o;
  o = take(h->bf16 ? N * h->cfg.n_mels * es : 0); w.xb = p + o;
  o = take(N * h->cfg.hidden * es); w.a1 = p + o;
  o = take(N * h->cfg.hidden * es); w.a2 = p + o;
  w.bytes = off;
  return w;
}

int pack_linear(fse_mel_encoder* h, const TensorTable& tt, const std::string& name, int N, int Cin, Layer& L) {
  int rc = FSE_OK;
  const float* w = tt.get(name + ".weight", static_cast<int64_t>(N) * Cin, &rc);
  if (rc) return rc;
  const float* b = tt.get(name + ".bias", N, &rc);
  if (rc) return rc;
  const int KB = mode_kb(h->cfg.mode);
  L.Cin = Cin; L.N = N; L.Kp = (Cin + KB - 1) / KB * KB; L.BN = N;
  std::vector<float> p(static_cast<size_t>(N) * L.Kp, 0.f);
  for (int o = 0; o < N; ++o)
    for (int c = 0; c < Cin; ++c) p[static_cast<size_t>(o) * L.Kp + c] = w[static_cast<size_t>(o) * Cin + c];
  FSE_TRY(upload_operand(p, h->bf16, &L.W, h->cfg.mode == FSE_MODE_TC_TF32));
  FSE_TRY(upload_f32(std::vector<float>(b, b + N), &L.bias));
  if (mode_is_tc(h->cfg.mode)) FSE_TRY(make_map_w(&L.map, L.W, L.Kp, N, KB, L.BN, h->bf16 ? 2 : 4));
  return FSE_OK;
}

template <typename TOp>
int forward_impl(fse_mel_encoder* h, const float* x, const float* add, const float* scale, float* out, int B, int T, void* ws, cudaStream_t st) {
  MWs w = mcarve(h, ws, B, T);
  const int M = h->cfg.n_mels, H = h->cfg.hidden, mode = h->cfg.mode;
  const bool tc = mode_is_tc(mode);
  const int zero = 0, KB = mode_kb(mode), es = h->bf16 ? 2 : 4;
  const void* x_src = h->bf16 ? w.xb : static_cast<const void*>(x);       // fp32 operands: the caller's tensor is the operand
  if (tc && !(h->plan.ws == ws && h->plan.B == B && h->plan.T == T && h->plan.x == x_src)) {
    FSE_TRY(make_map_act(&h->plan.mx, x_src, M, T, B, KB, kTileM, es));
    FSE_TRY(make_map_act(&h->plan.m1, w.a1, H, T, B, KB, kTileM, es));
    FSE_TRY(make_map_act(&h->plan.m2, w.a2, H, T, B, KB, kTileM, es));
    h->plan.ws = ws; h->plan.B = B; h->plan.T = T; h->plan.x = x_src;
  }
  const void* x_op = x;
  if constexpr (std::is_same<TOp, __nv_bfloat16>::value) {
    const size_t n = static_cast<size_t>(B) * T * M;
    mel_f32_to_bf16_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, st>>>(x, static_cast<__nv_bfloat16*>(w.xb), n / 4);
    FSE_CUDA(cudaGetLastError());
    ++h->launches;
    x_op = w.xb;
  }
  {
    ConvGemmParams p = make_params(B, T, T, M, 1, &zero, 0, H, KB);
    GemmOperands op; op.A0 = x_op; op.W = h->l[0].W; op.mA0 = &h->plan.mx; op.mW = &h->l[0].map; op.BN = h->l[0].BN;
    EpiRelu<TOp> epi{h->l[0].bias, static_cast<TOp*>(w.a1), H, T};
    FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, LaunchCtx{&h->launches, nullptr, 0})));
  }
  {
    ConvGemmParams p = make_params(B, T, T, H, 1, &zero, 0, H, KB);
    GemmOperands op; op.A0 = w.a1; op.W = h->l[1].W; op.mA0 = &h->plan.m1; op.mW = &h->l[1].map; op.BN = h->l[1].BN;
    EpiRelu<TOp> epi{h->l[1].bias, static_cast<TOp*>(w.a2), H, T};
    FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, LaunchCtx{&h->launches, nullptr, 0})));
  }
  {
    ConvGemmParams p = make_params(B, T, T, H, 1, &zero, 0, H, KB);
    GemmOperands op; op.A0 = w.a2; op.W = h->l[2].W; op.mA0 = &h->plan.m2; op.mW = &h->l[2].map; op.BN = h->l[2].BN;
    EpiCond epi{h->l[2].bias, add, scale, out, H, T};
    FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, LaunchCtx{&h->launches, nullptr, 0})));
  }
  return FSE_OK;
}

}  // namespace

extern "C" {

int fse_mel_encoder_create(const fse_mel_encoder_config* cfg, fse_mel_encoder** out) {
  if (!cfg || !out) return fail(FSE_EINVAL, "null argument");
  if (cfg->mode < 0 || cfg->mode > 3) return fail(FSE_EINVAL, "unknown mode %d", cfg->mode);
  if (cfg->n_mels <= 0 || cfg->n_mels % 8 != 0) return fail(FSE_EINVAL, "n_mels must be a positive multiple of 8");
  if (cfg->hidden <= 0 || cfg->hidden % 32 != 0 || cfg->hidden > 256) return fail(FSE_EINVAL, "hidden must be a multiple of 32, <= 256");
  int dev = 0;
  FSE_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  FSE_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(FSE_ECUDA, "device is sm_%d%d; this library is built for sm_100a only (no fallback)", prop.major, prop.minor);
  auto* h = new fse_mel_encoder();
  h->cfg = *cfg;
  h->bf16 = mode_is_bf16(cfg->mode);
  *out = h;
  return FSE_OK;
}

void fse_mel_encoder_destroy(fse_mel_encoder* h) {
  if (!h) return;
  for (auto& L : h->l) { if (L.W) cudaFree(L.W); if (L.bias) cudaFree(L.bias); }
  delete h;
}

int fse_mel_encoder_load_weights(fse_mel_encoder* h, const fse_tensor* tensors, int32_t n) {
  if (!h || !tensors || n <= 0) return fail(FSE_EINVAL, "null argument");
  if (h->loaded) return fail(FSE_ESTATE, "weights already loaded");
  TensorTable tt(tensors, n);
  FSE_TRY(pack_linear(h, tt, "encoder.0", h->cfg.hidden, h->cfg.n_mels, h->l[0]));
  FSE_TRY(pack_linear(h, tt, "encoder.2", h->cfg.hidden, h->cfg.hidden, h->l[1]));
  FSE_TRY(pack_linear(h, tt, "fc_out", h->cfg.hidden, h->cfg.hidden, h->l[2]));
  h->loaded = true;
  return FSE_OK;
}

int64_t fse_mel_encoder_workspace_bytes(const fse_mel_encoder* h, int32_t B, int32_t T) {
  if (!h || B <= 0 || T <= 0) return 0;
  return static_cast<int64_t>(mcarve(h, nullptr, B, T).bytes);
}

int fse_mel_encoder_forward(fse_mel_encoder* h, const float* x, const float* add, const float* scale, float* out, int32_t B, int32_t T,
                            void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h || !x || !out) return fail(FSE_EINVAL, "null argument");
  if (!h->loaded) return fail(FSE_ESTATE, "weights not loaded");
  if (B <= 0 || T <= 0) return fail(FSE_EINVAL, "B and T must be positive");
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 1023)) return fail(FSE_EINVAL, "workspace must be non-null and 1024-byte aligned");
  if (workspace_bytes < fse_mel_encoder_workspace_bytes(h, B, T)) return fail(FSE_EINVAL, "workspace too small");
  if ((static_cast<size_t>(B) * T * h->cfg.n_mels) % 4 != 0) return fail(FSE_EINVAL, "B*T*n_mels must be a multiple of 4");
  h->launches = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return h->bf16 ? forward_impl<__nv_bfloat16>(h, x, add, scale, out, B, T, workspace, st)
                 : forward_impl<float>(h, x, add, scale, out, B, T, workspace, st);
}

int64_t fse_mel_encoder_last_launches(const fse_mel_encoder* h) { return h ? h->launches : 0; }

}  // extern "C"
