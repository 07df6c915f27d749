// Row-wise building blocks shared by the condition encoder (cond_encoder.cu) and CampNet (campnet.cu): GEMM epilogue functors
// of the LN -> conv -> GELU -> conv residual blocks, the warp-per-row LayerNorm kernel, packed conv weights and the small
// launch context (tensor-map cache, launch counter) around the conv-as-GEMM primitive of conv_gemm.cuh.
// Everything has internal linkage (anonymous namespace): each translation unit instantiates what it uses.
#pragma once
#include <cstdlib>
#include <cmath>
#include <deque>
#include <string>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "fse_common.cuh"

namespace fse {
namespace {

constexpr int kMaxPerLane = 16;   // channels per lane in the warp-per-row kernels: C <= 512, C % 32 == 0

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void store_op(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_op(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------- GEMM epilogues
// ResidualBlock: gelu((conv_k(LN(x)) + b) * k^-0.5)   (conv.py:42-48; torch.nn.GELU() is the exact erf form)
template <typename TOp>
struct EpiGeluScale {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  const float* bias;
  TOp* out;   // [B*T, N]
  int N, T;
  float scale;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float y = __fmul_rn(acc[i] + __ldg(bias + n0 + i), scale);
      v[i] = 0.5f * y * (1.0f + erff(y * 0.70710678118654752f));
    }
    st_vec<NV>(out + (static_cast<size_t>(b) * T + t) * N + n0, v);
  }
};

// residual tail of a block: x = (x + (gemm + b)) [* nonpadding], x fp32 in place
// (ResidualBlock, conv.py:59-64; EncSALayer / DecSALayer, speech_editing/commons/transformer.py:514-529,574-608)
struct EpiResidualMask {
  static constexpr int kAux = 1;
  static constexpr bool kTransposed = true;
  const float* bias;
  float* x;            // [B*T, N] fp32 residual stream (read as aux, written here)
  const float* mask;   // [B*T] or null (no masking)
  int N, T;
  template <int NV>
  __device__ __forceinline__ void load_aux(int b, int t, int n0, float* aux) const {
    const size_t o = (static_cast<size_t>(b) * T + t) * N + n0;
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      const float4 v = reinterpret_cast<const float4*>(x + o)[i];
      aux[4 * i] = v.x; aux[4 * i + 1] = v.y; aux[4 * i + 2] = v.z; aux[4 * i + 3] = v.w;
    }
  }
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float* aux) const {
    const size_t row = static_cast<size_t>(b) * T + t;
    const float m = mask ? __ldg(mask + row) : 1.f;
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = __fadd_rn(aux[i], acc[i] + __ldg(bias + n0 + i));
      if (mask) v[i] = __fmul_rn(v[i], m);
    }
    st_vec<NV>(x + row * N + n0, v);
  }
};

// predictor layers: relu(conv_k(.) + b) in fp32 (LayerNorm follows; nar_tts_modules.py:16-21, 83-88)
struct EpiReluF32 {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  const float* bias;
  float* out;   // [B*T, N]
  int N, T;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = fmaxf(acc[i] + __ldg(bias + n0 + i), 0.f);
    st_vec<NV>(out + (static_cast<size_t>(b) * T + t) * N + n0, v);
  }
};

// post_net1: (conv_k(.) + b) * nonpadding   (conv.py:113)
struct EpiBiasMaskF32 {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  const float* bias;
  const float* mask;   // [B*T]
  float* out;          // [B*T, N]
  int N, T;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    const size_t row = static_cast<size_t>(b) * T + t;
    const float m = __ldg(mask + row);
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __fmul_rn(acc[i] + __ldg(bias + n0 + i), m);
    st_vec<NV>(out + row * N + n0, v);
  }
};

// ---------------------------------------------------------------------------------------------- warp-per-row kernels
// channel LayerNorm of one row (layers.py:5-24; biased variance, eps inside the sqrt), two-pass in registers.
//   in_scale[row]  (optional) multiplies the input first       (ConvBlocks: res_blocks(x) * nonpadding, conv.py:111)
//   out_scale[row] (optional) multiplies the output            (last_norm(x) * nonpadding; predictor padding masks)
//   mask_out[row]  (optional) receives (sum_c |input| > 0)     (ResidualBlock's own nonpadding, conv.py:58)
//   out_op: operand of the next GEMM (bf16 / fp32 by mode); out_f32: fp32 copy for a CUDA-core head; either may be null
template <typename TOp>
__global__ void __launch_bounds__(256) row_layer_norm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ in_scale,
                                                              const float* __restrict__ out_scale, float* __restrict__ mask_out,
                                                              TOp* __restrict__ out_op, float* __restrict__ out_f32, int rows, int C,
                                                              float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<size_t>(row) * C;
  const int n = C >> 5;
  const float si = in_scale ? __ldg(in_scale + row) : 1.f;
  float v[kMaxPerLane];
  float s = 0.f, sa = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    v[i] = 0.f;
    if (i < n) {
      float t = xr[lane + 32 * i];
      if (in_scale) t = __fmul_rn(t, si);
      v[i] = t; s += t; sa += fabsf(t);
    }
  }
  s = warp_sum(s);
  sa = warp_sum(sa);
  const float mean = s / static_cast<float>(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i)
    if (i < n) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  q = warp_sum(q);
  const float rstd = 1.0f / sqrtf(q / static_cast<float>(C) + eps);
  if (mask_out && lane == 0) mask_out[row] = sa > 0.f ? 1.f : 0.f;
  const float so = out_scale ? __ldg(out_scale + row) : 1.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    if (i < n) {
      const int c = lane + 32 * i;
      float y = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      if (out_scale) y = __fmul_rn(y, so);
      if (out_op) store_op(out_op + static_cast<size_t>(row) * C + c, y);
      if (out_f32) out_f32[static_cast<size_t>(row) * C + c] = y;
    }
  }
}

// The same LayerNorm for C % 64 == 0, C <= 256 (hidden 192 / 256): 16 lanes per row, float4 loads (a half-warp reads 256 contiguous
// bytes per instruction), 8- or 16-byte stores — the scalar kernel above moved 4 B per lane per load and 2 B per store and ran at
// 0.39 of the HBM roofline (30 us for 65 536 x 192 rows against 11.5 us of traffic; ncu launch list of the CampNet forward).
__device__ __forceinline__ void store_op4(float* p, const float* y) { *reinterpret_cast<float4*>(p) = make_float4(y[0], y[1], y[2], y[3]); }
__device__ __forceinline__ void store_op4(__nv_bfloat16* p, const float* y) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(y[0], y[1]), hi = __floats2bfloat162_rn(y[2], y[3]);
  uint2 w;
  w.x = *reinterpret_cast<const uint32_t*>(&lo);
  w.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = w;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename TOp>
__global__ void __launch_bounds__(256) row_layer_norm_vec_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, const float* __restrict__ in_scale,
                                                                  const float* __restrict__ out_scale, float* __restrict__ mask_out,
                                                                  TOp* __restrict__ out_op, float* __restrict__ out_f32, int rows, int C,
                                                                  float eps) {
  constexpr int kMaxVec = 4;
  const int row = blockIdx.x * 16 + (threadIdx.x >> 4), sub = threadIdx.x & 15;
  const bool live = row < rows;                                   // whole half-warps go idle together; shuffles stay inside a half-warp
  const int nv = C >> 6;                                          // float4 per lane
  const float* xr = x + static_cast<size_t>(live ? row : 0) * C;
  const float si = (in_scale && live) ? __ldg(in_scale + row) : 1.f;
  float4 v[kMaxVec];
  float s = 0.f, sa = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nv && live) {
      float4 t = reinterpret_cast<const float4*>(xr)[sub + 16 * i];
      if (in_scale) { t.x = __fmul_rn(t.x, si); t.y = __fmul_rn(t.y, si); t.z = __fmul_rn(t.z, si); t.w = __fmul_rn(t.w, si); }
      v[i] = t;
      s += (t.x + t.y) + (t.z + t.w);
      sa += (fabsf(t.x) + fabsf(t.y)) + (fabsf(t.z) + fabsf(t.w));
    }
  }
  s = half_warp_sum(s);
  sa = half_warp_sum(sa);
  const float mean = s / static_cast<float>(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (i < nv) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q = fmaf(dx, dx, q); q = fmaf(dy, dy, q); q = fmaf(dz, dz, q); q = fmaf(dw, dw, q);
    }
  q = half_warp_sum(q);
  if (!live) return;
  const float rstd = 1.0f / sqrtf(q / static_cast<float>(C) + eps);
  if (mask_out && sub == 0) mask_out[row] = sa > 0.f ? 1.f : 0.f;
  const float so = out_scale ? __ldg(out_scale + row) : 1.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    if (i < nv) {
      const int c = 4 * (sub + 16 * i);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c)), bb = __ldg(reinterpret_cast<const float4*>(beta + c));
      float y[4] = {(v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y, (v[i].z - mean) * rstd * g.z + bb.z,
                    (v[i].w - mean) * rstd * g.w + bb.w};
      if (out_scale) {
#pragma unroll
        for (int k = 0; k < 4; ++k) y[k] = __fmul_rn(y[k], so);
      }
      if (out_op) store_op4(out_op + static_cast<size_t>(row) * C + c, y);
      if (out_f32) store_op4(out_f32 + static_cast<size_t>(row) * C + c, y);
    }
  }
}

// mask[row] = (sum_c |x[row, c]| > 0): the data-derived nonpadding of ConvBlocks / TransformerDecoder
// (modules/commons/conv.py:104, speech_editing/commons/transformer.py:784); optional: first[row] = (x[row, 0] != 0), the
// flag make_positions sees when it is handed x[..., 0] (transformer.py:787)
__global__ void __launch_bounds__(256) row_absmask_kernel(const float* __restrict__ x, float* __restrict__ mask, float* __restrict__ first,
                                                          int rows, int C) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float sa = 0.f;
  for (int c = lane; c < C; c += 32) sa += fabsf(x[static_cast<size_t>(row) * C + c]);
  sa = warp_sum(sa);
  if (lane == 0) {
    mask[row] = sa > 0.f ? 1.f : 0.f;
    if (first) first[row] = x[static_cast<size_t>(row) * C] != 0.f ? 1.f : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------- host side
struct ConvW {
  void* W = nullptr; float* bias = nullptr; CUtensorMap map{};
  int Cin = 0, N = 0, ntaps = 0, KB = 64, Kp = 0, BN = 0;
  int offs[kMaxTaps] = {};
};
struct LNW { float* g = nullptr; float* b = nullptr; };

// what a handle needs around the GEMM primitive: arithmetic mode, its device allocations, a cache of activation tensor maps
struct LayerCtx {
  int mode = FSE_MODE_TC_BF16;
  bool bf16 = true;
  int hidden = 0;
  struct MapEntry { const void* buf; int C, T, B, KB, rows = 128; CUtensorMap map; };
  std::deque<MapEntry> cache;
  std::vector<void*> owned;             // every device allocation of the handle
  long long launches = 0;
  void release() { for (void* p : owned) cudaFree(p); owned.clear(); }
};

// scratch rows shared by the stages of a handle: three fp32 streams, two operand buffers (the second 4x wide at most), two masks
struct RowBufs { float* x32; float* tmp32; float* y32; void* opA; void* opB; float* m0; float* m1; size_t bytes; };
inline RowBufs carve_rows(const LayerCtx& ctx, void* base, size_t rows, int wide = 2) {
  const size_t H = static_cast<size_t>(ctx.hidden), es = ctx.bf16 ? 2 : 4;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += align_up(b, 1024); return o; };
  uint8_t* p = static_cast<uint8_t*>(base);
  RowBufs w{};
  w.x32 = reinterpret_cast<float*>(p + take(rows * H * 4));
  w.tmp32 = reinterpret_cast<float*>(p + take(rows * H * 4));
  w.y32 = reinterpret_cast<float*>(p + take(rows * H * 4));
  w.opA = p + take(rows * H * es);
  w.opB = p + take(rows * wide * H * es);
  w.m0 = reinterpret_cast<float*>(p + take(rows * 4));
  w.m1 = reinterpret_cast<float*>(p + take(rows * 4));
  w.bytes = off;
  return w;
}

inline int dev_f32(LayerCtx* ctx, const float* src, size_t n, float** out) {
  FSE_TRY(upload_f32(std::vector<float>(src, src + n), out));
  ctx->owned.push_back(*out);
  return FSE_OK;
}
inline int load_vec(LayerCtx* ctx, const TensorTable& tt, const std::string& name, int64_t numel, float** out) {
  int rc = FSE_OK;
  const float* p = tt.get(name, numel, &rc);
  if (rc) return rc;
  return dev_f32(ctx, p, static_cast<size_t>(numel), out);
}
// embedding table whose row count comes from the checkpoint (dur_embed: 2000 rows, pitch_embed: 300, fs.py:67,74)
inline int load_table(LayerCtx* ctx, const TensorTable& tt, const std::string& name, int C, float** out, int* rows) {
  auto it = tt.map.find(name);
  if (it == tt.map.end()) return fail(FSE_EINVAL, "missing weight tensor '%s'", name.c_str());
  const int64_t numel = it->second->numel;
  if (numel <= 0 || numel % C != 0) return fail(FSE_EINVAL, "weight '%s' has %lld elements, not a multiple of hidden %d", name.c_str(),
                                                static_cast<long long>(numel), C);
  *rows = static_cast<int>(numel / C);
  return dev_f32(ctx, it->second->data, static_cast<size_t>(numel), out);
}
inline int load_ln(LayerCtx* ctx, const TensorTable& tt, const std::string& name, int C, LNW& ln) {
  FSE_TRY(load_vec(ctx, tt, name + ".weight", C, &ln.g));
  return load_vec(ctx, tt, name + ".bias", C, &ln.b);
}

// A conv / linear weight w[Cout, Cin, k] (host, row-major) with tap offsets offs[k] -> packed [Cout, k * nkb * 64]; bias may be
// null (a zero vector is uploaded so that epilogues need no branch).  Tile width: one of the widths the GEMM tests exercise.
inline int pack_conv_raw(LayerCtx* ctx, const std::string& name, const float* w, const float* bias, int Cout, int Cin, int k, const int* offs,
                         ConvW& cw) {
  if (k > kMaxTaps || k < 1) return fail(FSE_EINVAL, "%s: %d taps unsupported (1..%d)", name.c_str(), k, kMaxTaps);
  const int KB = mode_kb(ctx->mode);                  // 128 bytes of K per row: 64 bf16 or 32 fp32 (tf32) channels
  cw.Cin = Cin; cw.N = Cout; cw.ntaps = k; cw.KB = KB;
  const int nkb = (Cin + KB - 1) / KB;
  cw.Kp = k * nkb * KB;
  cw.BN = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : (Cout % 32 == 0 ? 32 : (Cout % 16 == 0 && Cout <= 256 ? Cout : 0))));
  // hidden 192 and its multiples (every projection / FFN of CampNet and of the condition encoder): 192-wide tiles instead of three
  // 64-wide (or 128-wide) ones — the activation tile is read once per 192 outputs and the MMA is N = 192 (an N = 64 instruction is
  // bound by its shared-memory operand reads).  FSE_BN192=0 restores the old choice.
  static const bool bn192 = !(std::getenv("FSE_BN192") && std::getenv("FSE_BN192")[0] == '0');
  if (bn192 && Cout % 192 == 0 && Cout % 256 != 0) cw.BN = 192;
  if (cw.BN == 0) return fail(FSE_EINVAL, "%s: %d output channels is not a multiple of 16", name.c_str(), Cout);
  for (int j = 0; j < k; ++j) cw.offs[j] = offs[j];
  std::vector<float> p(static_cast<size_t>(Cout) * cw.Kp, 0.f);
  for (int o = 0; o < Cout; ++o)
    for (int c = 0; c < Cin; ++c)
      for (int j = 0; j < k; ++j) p[static_cast<size_t>(o) * cw.Kp + j * nkb * KB + c] = w[(static_cast<size_t>(o) * Cin + c) * k + j];
  FSE_TRY(upload_operand(p, ctx->bf16, &cw.W, ctx->mode == FSE_MODE_TC_TF32));
  ctx->owned.push_back(cw.W);
  if (bias) {
    FSE_TRY(dev_f32(ctx, bias, Cout, &cw.bias));
  } else {
    std::vector<float> z(Cout, 0.f);
    FSE_TRY(dev_f32(ctx, z.data(), Cout, &cw.bias));
  }
  if (mode_is_tc(ctx->mode)) FSE_TRY(make_map_w(&cw.map, cw.W, cw.Kp, cw.N, cw.KB, cw.BN, ctx->bf16 ? 2 : 4));
  return FSE_OK;
}
// Conv1d(k, dilation dil) from a state_dict: "same" padding (tap j at (j - (k-1)/2) dil) or, with left = true, the causal
// 'LEFT' padding of TransformerFFNLayer (tap j at j - (k-1); speech_editing/commons/transformer.py:84-88)
inline int pack_conv(LayerCtx* ctx, const TensorTable& tt, const std::string& name, int Cout, int Cin, int k, int dil, ConvW& cw,
                     bool left = false) {
  int rc = FSE_OK;
  const float* w = tt.get(name + ".weight", static_cast<int64_t>(Cout) * Cin * k, &rc);
  if (rc) return rc;
  const float* bias = tt.get(name + ".bias", Cout, &rc);
  if (rc) return rc;
  if (k > kMaxTaps || (!left && k % 2 == 0)) return fail(FSE_EINVAL, "%s: kernel size %d unsupported (odd, <= %d)", name.c_str(), k, kMaxTaps);
  int offs[kMaxTaps];
  for (int j = 0; j < k; ++j) offs[j] = left ? (j - (k - 1)) * dil : (j - (k - 1) / 2) * dil;
  return pack_conv_raw(ctx, name, w, bias, Cout, Cin, k, offs, cw);
}

inline int get_act_map(LayerCtx* ctx, const void* buf, int C, int T, int B, int KB, const CUtensorMap** out, int rows = kTileM) {
  for (auto& e : ctx->cache)
    if (e.buf == buf && e.C == C && e.T == T && e.B == B && e.KB == KB && e.rows == rows) { *out = &e.map; return FSE_OK; }
  if (ctx->cache.size() >= 96) ctx->cache.clear();      // maps are copied into the launch, dropping them is safe
  ctx->cache.emplace_back();
  auto& e = ctx->cache.back();
  e.buf = buf; e.C = C; e.T = T; e.B = B; e.KB = KB; e.rows = rows;
  const int rc = make_map_act(&e.map, buf, C, T, B, KB, rows, ctx->bf16 ? 2 : 4);
  if (rc != FSE_OK) { ctx->cache.pop_back(); return rc; }
  *out = &e.map;
  return FSE_OK;
}

// one conv / linear layer over a channels-last activation A[B, T, cw.Cin] (operand type), epilogue functor epi
template <typename TOp, class Epi>
inline int run_conv(LayerCtx* ctx, const ConvW& cw, const void* A, int B, int T, const Epi& epi, cudaStream_t st) {
  ConvGemmParams p = make_params(B, T, T, cw.Cin, cw.ntaps, cw.offs, 0, cw.N, cw.KB);
  GemmOperands op; op.A0 = A; op.W = cw.W; op.mW = &cw.map; op.BN = cw.BN;
  if (mode_is_tc(ctx->mode)) {
    // multi-tap convs (k = 9 FFN, k = 5 encoder / predictor convs): one activation load per channel block, every tap a row-shifted
    // descriptor of that copy — the per-tap loads made the k = 9 FFN conv ingest-bound (9 x 16 KB per channel block and tile)
    static const bool shared_ok = !(std::getenv("FSE_ROW_SHARED_A") && std::getenv("FSE_ROW_SHARED_A")[0] == '0');
    int rows = kTileM;
    if (shared_ok && cw.ntaps >= 2 && enable_shared_a(p, 1, ctx->bf16 ? 2 : 4)) rows = p.Rbox;
    FSE_TRY(get_act_map(ctx, A, cw.Cin, T, B, cw.KB, &op.mA0, rows));
  }
  return run_conv_gemm<TOp>(ctx->mode, p, op, epi, st, LaunchCtx{&ctx->launches, nullptr, 0});
}

inline unsigned row_blocks(size_t rows) { return static_cast<unsigned>((rows + 7) / 8); }

template <typename TOp>
inline int layer_norm(LayerCtx* ctx, const float* x, const LNW& ln, const float* in_scale, const float* out_scale, float* mask_out,
                      void* out_op, float* out_f32, size_t rows, cudaStream_t st) {
  const int C = ctx->hidden;
  const bool aligned = (reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out_op) | reinterpret_cast<uintptr_t>(out_f32) |
                        reinterpret_cast<uintptr_t>(ln.g) | reinterpret_cast<uintptr_t>(ln.b)) % 16 == 0;
  if (C % 64 == 0 && C <= 256 && aligned)
    row_layer_norm_vec_kernel<TOp><<<static_cast<unsigned>((rows + 15) / 16), 256, 0, st>>>(x, ln.g, ln.b, in_scale, out_scale, mask_out,
                                                                                          static_cast<TOp*>(out_op), out_f32, static_cast<int>(rows), C, 1e-5f);
  else
    row_layer_norm_kernel<TOp><<<row_blocks(rows), 256, 0, st>>>(x, ln.g, ln.b, in_scale, out_scale, mask_out, static_cast<TOp*>(out_op),
                                                                out_f32, static_cast<int>(rows), C, 1e-5f);
  FSE_CUDA(cudaGetLastError());
  ++ctx->launches;
  return FSE_OK;
}

// ConvBlocks (modules/commons/conv.py:68-116, norm_type 'ln'): n_blocks ResidualBlocks of layers_in_block x
// [LN -> Conv k (dilation d) -> * k^-0.5 -> GELU -> Conv 1x1, + residual, * nonpadding], then * nonpadding, last_norm * nonpadding,
// post_net1 (its epilogue is the caller's: * nonpadding and whatever follows).
struct ConvBlocksW {
  int n_blocks = 0, layers_in_block = 0, kernel = 0;
  std::vector<LNW> ln;              // [block * layers_in_block + j]
  std::vector<ConvW> c1, c2;
  LNW last_norm;
  ConvW post;
};
inline int load_conv_blocks(LayerCtx* ctx, const TensorTable& tt, const std::string& prefix, int n_blocks, const int* dilations,
                            int layers_in_block, int kernel, int post_kernel, ConvBlocksW& w) {
  const int H = ctx->hidden, nsub = n_blocks * layers_in_block;
  w.n_blocks = n_blocks; w.layers_in_block = layers_in_block; w.kernel = kernel;
  w.ln.resize(nsub); w.c1.resize(nsub); w.c2.resize(nsub);
  for (int i = 0; i < n_blocks; ++i)
    for (int j = 0; j < layers_in_block; ++j) {
      const int q = i * layers_in_block + j;
      const std::string pre = prefix + "res_blocks." + std::to_string(i) + ".blocks." + std::to_string(j) + ".";
      FSE_TRY(load_ln(ctx, tt, pre + "0", H, w.ln[q]));
      FSE_TRY(pack_conv(ctx, tt, pre + "1", 2 * H, H, kernel, dilations ? dilations[i] : 1, w.c1[q]));
      FSE_TRY(pack_conv(ctx, tt, pre + "4", H, 2 * H, 1, 1, w.c2[q]));
    }
  FSE_TRY(load_ln(ctx, tt, prefix + "last_norm", H, w.last_norm));
  return pack_conv(ctx, tt, prefix + "post_net1", H, H, post_kernel, 1, w.post);
}
// r.x32 holds the input [B*T, H] fp32 (overwritten), r.m0 the ConvBlocks-level nonpadding; epi_post consumes post_net1
template <typename TOp, class EpiPost>
inline int conv_blocks_forward(LayerCtx* ctx, const ConvBlocksW& w, const RowBufs& r, int B, int T, const EpiPost& epi_post, cudaStream_t st) {
  const size_t rows = static_cast<size_t>(B) * T;
  const int H = ctx->hidden;
  const float kscale = static_cast<float>(std::pow(static_cast<double>(w.kernel), -0.5));
  for (int i = 0; i < w.n_blocks; ++i)
    for (int j = 0; j < w.layers_in_block; ++j) {
      const int q = i * w.layers_in_block + j;
      // the block's nonpadding comes from its own input: computed by the first LayerNorm pass over it
      FSE_TRY((layer_norm<TOp>(ctx, r.x32, w.ln[q], nullptr, nullptr, j == 0 ? r.m1 : nullptr, r.opA, nullptr, rows, st)));
      EpiGeluScale<TOp> e1{w.c1[q].bias, static_cast<TOp*>(r.opB), w.c1[q].N, T, kscale};
      FSE_TRY((run_conv<TOp>(ctx, w.c1[q], r.opA, B, T, e1, st)));
      EpiResidualMask e2{w.c2[q].bias, r.x32, r.m1, H, T};
      FSE_TRY((run_conv<TOp>(ctx, w.c2[q], r.opB, B, T, e2, st)));
    }
  FSE_TRY((layer_norm<TOp>(ctx, r.x32, w.last_norm, r.m0, r.m0, nullptr, r.opA, nullptr, rows, st)));
  return run_conv<TOp>(ctx, w.post, r.opA, B, T, epi_post, st);
}

}  // namespace
}  // namespace fse
