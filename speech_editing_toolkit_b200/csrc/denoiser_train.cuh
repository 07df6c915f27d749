// Training-mode DiffNet (BASELINE configs[4]): the forward of
// `x_0_pred = denoise_fn(x_t, t, cond)` (spec_denoiser.py:168-176, diffnet.py:110-132) with what the backward needs kept, and the
// activation-gradient chain of its backward, as conv-GEMM launches of this library (tcgen05 in the tensor-core modes) with the
// layer tails in their epilogues.  Included at the end of denoiser.cu.
//
// What is native here and what is not.  Native: every GEMM on the data path forwards (input projection, gated conv, output
// projection with residual / skip accumulation, skip and output projections) and backwards (the dgrad of each of them: a
// conv-GEMM over the transposed / tap-reversed weights), plus the elementwise pieces between them.  The weight gradients are
// GEMMs over tensors this code leaves in the workspace (dW = dY^T A with K = all frames): csrc/wgrad.cu (fse_wgrad, tcgen05 over
// MN-major operands), called per weight by speech_editing_toolkit_b200/train.py, which also owns the bias sums, the timestep-MLP
// backward (a [B, 256] problem), the optimizer and the NCCL all-reduce.
//
//   forward, per layer l:   hin = h + d_l                      (add_bcast_cast_kernel; d_l = diffusion_projection(temb), given)
//                           y   = conv_k3(hin) + W_cp cond + b (GEMM, K = 3C + H, fp32 out)
//                           u   = sigmoid(y[:, :C]) * tanh(y[:, C:])      (gate_fwd_kernel; du/dg and du/df kept)
//                           o   = W_op u + b;  h <- (h + o[:, :C]) / sqrt(2);  S += o[:, C:]     (GEMM, EpiResSkip)
//            tail:          x0  = W_out relu(W_skip (S / sqrt(L)) + b) + b
//   backward, given dx0:    dz = (W_out^T dx0) * [r > 0];  dS = W_skip^T dz / sqrt(L)
//            per layer l (downwards), with dh the gradient of the residual stream (0 above the top layer):
//                           du = W_op^T [dh / sqrt(2) | dS]                 (two-source GEMM, EpiGateBwd -> dy_l = [dg | df])
//                           dh <- dh / sqrt(2) + conv_k3^T(dy_l)            (GEMM over tap-reversed W_dc^T, EpiDh)
//            at the end:    dcond = sum_l W_cp,l^T dy_l                      (ONE GEMM over dy of all layers, K = L * 2C)
//
// Arithmetic follows the handle's mode: FSE_MODE_SIMT_F32 (exact fp32, the gradient-parity reference), FSE_MODE_TC_BF16
// (BASELINE configs[4] says bf16), FSE_MODE_TC_TF32.
#pragma once

namespace fse {

// Operand copies written for a kind::tf32 GEMM are rounded to tf32 (round-to-nearest) here: the tensor core would otherwise
// TRUNCATE the fp32 container, a one-sided error twice as large (measured on the gradient fixture: 2.1e-2 -> see the test).
template <int NV>
__device__ __forceinline__ void st_operand(float* p, const float* v, bool rnd) {
  float w[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) w[i] = rnd ? ptx::round_tf32(v[i]) : v[i];
  st_vec<NV>(p, w);
}
template <int NV>
__device__ __forceinline__ void st_operand(__nv_bfloat16* p, const float* v, bool) { st_vec<NV>(p, v); }

// NV consecutive operand-type values -> fp32 (vector loads: NV is a multiple of 4).  The loads carry the L2::256B prefetch hint: an
// epilogue chunk touches only 64 (bf16) or 128 (fp32) bytes of a frame's row, the neighbouring pieces of the row are read by later
// chunks - fetched piecewise from DRAM, every piece re-opens the DRAM page (measured: 1.35 TB/s for the whole launch).
template <int NV>
__device__ __forceinline__ void ld_operand(const float* p, float* out) {
#pragma unroll
  for (int i = 0; i < NV / 4; ++i)
    asm volatile("ld.global.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(out[4 * i]), "=f"(out[4 * i + 1]), "=f"(out[4 * i + 2]), "=f"(out[4 * i + 3]) : "l"(p + 4 * i));
}
template <int NV>
__device__ __forceinline__ void ld_operand(const __nv_bfloat16* p, float* out) {
#pragma unroll
  for (int i = 0; i < NV / 4; ++i) {
    uint2 w;
    asm volatile("ld.global.L2::256B.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(p + 4 * i));
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&w.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&w.y);
    out[4 * i] = __low2float(lo); out[4 * i + 1] = __high2float(lo); out[4 * i + 2] = __low2float(hi); out[4 * i + 3] = __high2float(hi);
  }
}

// ------------------------------------------------------------------ epilogues
template <typename TOp>
struct EpiStoreF32 {                       // out = acc + bias (fp32 rows)
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  const float* bias0;   // [N] or null
  const float* bias1;   // [N] or null (second bias, e.g. dilated_conv.bias + conditioner_projection.bias)
  float* out;           // [B*T, N]
  int N, T;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = acc[i] + (bias0 ? __ldg(bias0 + n0 + i) : 0.f) + (bias1 ? __ldg(bias1 + n0 + i) : 0.f);
    st_vec<NV>(out + (static_cast<size_t>(b) * T + t) * N + n0, v);
  }
};

// o = W_op u + b: columns [0, C) update the residual stream, columns [C, 2C) accumulate the skip sum (diffnet.py:79-81, :126)
template <typename TOp, bool Fast>
struct EpiResSkip {
  static constexpr int kAux = 1;
  static constexpr bool kTransposed = true;
  const float* bias;    // [2C]
  float* h;             // [B*T, C]
  float* S;             // [B*T, C]
  int C, T;
  template <int NV>
  __device__ __forceinline__ void load_aux(int b, int t, int n0, float* aux) const {
    const float* src = n0 < C ? h + (static_cast<size_t>(b) * T + t) * C + n0 : S + (static_cast<size_t>(b) * T + t) * C + (n0 - C);
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      const float4 v = reinterpret_cast<const float4*>(src)[i];
      aux[4 * i] = v.x; aux[4 * i + 1] = v.y; aux[4 * i + 2] = v.z; aux[4 * i + 3] = v.w;
    }
  }
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float* aux) const {
    float v[NV];
    if (n0 < C) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float sum = aux[i] + (acc[i] + __ldg(bias + n0 + i));
        v[i] = Fast ? sum * 0.70710678118654752440f : __fdiv_rn(sum, 1.41421356237309504880f);
      }
      st_vec<NV>(h + (static_cast<size_t>(b) * T + t) * C + n0, v);
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = aux[i] + (acc[i] + __ldg(bias + n0 + i));
      st_vec<NV>(S + (static_cast<size_t>(b) * T + t) * C + (n0 - C), v);
    }
  }
};

// dz = acc * [r > 0]   (backward of relu(skip_projection(.)), r kept from the forward)
template <typename TOp>
struct EpiReluBwd {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  const TOp* r;         // [B*T, N]
  TOp* out;             // [B*T, N]
  int N, T;
  bool rnd;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    const size_t o = (static_cast<size_t>(b) * T + t) * N + n0;
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = to_f32(r[o + i]) > 0.f ? acc[i] : 0.f;
    st_operand<NV>(out + o, v, rnd);
  }
};

template <typename TOp>
struct EpiScaleCast {                      // out = op(acc * scale)
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  TOp* out;
  int N, T;
  float scale;
  bool rnd;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = acc[i] * scale;
    st_operand<NV>(out + (static_cast<size_t>(b) * T + t) * N + n0, v, rnd);
  }
};

// du -> dy = [dg | df]:  dg = du * (tanh(f) sg (1 - sg)),  df = du * (sg (1 - tanh(f)^2))   (backward of u = sigmoid(g) tanh(f);
// the two factors were formed in fp32 by gate_fwd_kernel)
template <typename TOp>
struct EpiGateBwd {
  // The two derivative factors are "late" operands (kLate = 2 per column): the epilogue issues the loads of a whole chunk (8 x 4 frames)
  // together, as 8-byte vectors with the L2::256B prefetch hint, before the chunk's first store.  ncu (profiles/
  // r02l_ncu_train_backward_kernels.md): 61 % of the stall samples are the epilogue warps waiting for these loads and the whole launch
  // moves 1.35 TB/s - every chunk touches 64 bytes of each frame's 512-byte row, piecewise DRAM fetches of rows written 20 layers ago.
  // Vector loads alone changed nothing (72 us per launch), prefetching a chunk ahead (kAux = 2) was slower (registers), the 256-byte L2
  // fetch took it to 65 us; a full-row mapping of the epilogue is what would fix it.
  static constexpr int kAux = 0;
  static constexpr int kLate = 2;
  static constexpr bool kTransposed = true;
  const TOp* sg;        // [B*T, C] du/dg of this layer
  const TOp* tf;        // [B*T, C] du/df
  TOp* dy;              // [B*T, ld] rows; this layer's 2C columns start at col0
  int C, T, ld, col0;
  bool rnd;
  template <int NV>
  __device__ __forceinline__ void load_late(int b, int t, int n0, float* dst) const {
    const size_t o = (static_cast<size_t>(b) * T + t) * C + n0;
    ld_operand<NV>(sg + o, dst);
    ld_operand<NV>(tf + o, dst + NV);
  }
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float* aux) const {
    const size_t row = static_cast<size_t>(b) * T + t;
    float dg[NV], df[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      dg[i] = acc[i] * aux[i];
      df[i] = acc[i] * aux[NV + i];
    }
    st_operand<NV>(dy + row * ld + col0 + n0, dg, rnd);
    st_operand<NV>(dy + row * ld + col0 + C + n0, df, rnd);
  }
};

// dh <- dh / sqrt(2) + acc;  dres_below = op(dh / sqrt(2)) for the layer below (its output-projection dgrad / wgrad operand)
template <typename TOp>
struct EpiDh {
  static constexpr int kAux = 1;
  static constexpr bool kTransposed = true;
  float* dh;            // [B*T, C]
  TOp* dres_below;      // [B*T, C] or null (layer 0)
  int C, T;
  bool rnd;
  template <int NV>
  __device__ __forceinline__ void load_aux(int b, int t, int n0, float* aux) const {
    const float4* p = reinterpret_cast<const float4*>(dh + (static_cast<size_t>(b) * T + t) * C + n0);
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      const float4 v = p[i];
      aux[4 * i] = v.x; aux[4 * i + 1] = v.y; aux[4 * i + 2] = v.z; aux[4 * i + 3] = v.w;
    }
  }
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float* aux) const {
    const size_t o = (static_cast<size_t>(b) * T + t) * C + n0;
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = aux[i] * 0.70710678118654752440f + acc[i];
    st_vec<NV>(dh + o, v);
    if (dres_below) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] *= 0.70710678118654752440f;
      st_operand<NV>(dres_below + o, v, rnd);
    }
  }
};

// ------------------------------------------------------------------ elementwise kernels
template <typename TOp>
__global__ void __launch_bounds__(256) add_bcast_cast_kernel(const float* __restrict__ h, const float* __restrict__ d, TOp* __restrict__ out,
                                                             int T, int C, size_t n4, bool rnd) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const size_t e = i * 4, row = e / C;
  const int c = static_cast<int>(e - row * C), b = static_cast<int>(row / T);
  const float4 hv = reinterpret_cast<const float4*>(h)[i];
  const float4 dv = *reinterpret_cast<const float4*>(d + static_cast<size_t>(b) * C + c);
  const float v[4] = {hv.x + dv.x, hv.y + dv.y, hv.z + dv.z, hv.w + dv.w};
  st_operand<4>(out + e, v, rnd);
}

template <typename TOp, bool Fast>
__global__ void __launch_bounds__(256) gate_fwd_kernel(const float* __restrict__ y, TOp* __restrict__ sg, TOp* __restrict__ tf, TOp* __restrict__ u,
                                                       int C, size_t n4, bool rnd) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const size_t e = i * 4, row = e / C;
  const int c = static_cast<int>(e - row * C);
  const float4 g = *reinterpret_cast<const float4*>(y + row * 2 * C + c);
  const float4 f = *reinterpret_cast<const float4*>(y + row * 2 * C + C + c);
  const float gs[4] = {sigmoid_f<Fast>(g.x), sigmoid_f<Fast>(g.y), sigmoid_f<Fast>(g.z), sigmoid_f<Fast>(g.w)};
  const float ts[4] = {tanh_f<Fast>(f.x), tanh_f<Fast>(f.y), tanh_f<Fast>(f.z), tanh_f<Fast>(f.w)};
  float us[4], da[4], db[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    us[q] = gs[q] * ts[q];
    // kept for the backward: the two derivative factors themselves, formed in fp32 BEFORE rounding to the operand type
    // (1 - tanh^2 of a bf16-rounded tanh near +-1 would lose every significant bit)
    da[q] = ts[q] * gs[q] * (1.f - gs[q]);              // du/dg
    db[q] = gs[q] * (1.f - ts[q] * ts[q]);              // du/df
  }
  st_vec<4>(sg + e, da);
  st_vec<4>(tf + e, db);
  st_operand<4>(u + e, us, rnd);
}

template <typename TOp>
__global__ void __launch_bounds__(256) scale_cast_kernel(const float* __restrict__ src, TOp* __restrict__ dst, float scale, size_t n4, bool rnd) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 s = reinterpret_cast<const float4*>(src)[i];
  const float v[4] = {s.x * scale, s.y * scale, s.z * scale, s.w * scale};
  st_operand<4>(dst + 4 * i, v, rnd);
}

// dst[r * ldd + col0 + j * cblk + c] = op(src[r * sr + c * sc + j * sj] (+ add[...]))   (weight repacking on the device, once per optimizer step)
// All repacking jobs of one fse_train_load_weights_device call in ONE launch (125 jobs for 20 layers: as separate launches they cost
// 0.4 ms of device time and as much host time per optimizer step).  Block b belongs to the job whose block range contains it.
struct PackJob {
  void* dst; const float* src; const float* add;       // add: optional second source with the same indexing (summed biases)
  long ldd, col0, sr, sc, sj, cblk;
  int R, Cn, J, block0;                 // block0: first block of this job
};
template <typename TOp>
__global__ void __launch_bounds__(256) pack_jobs_kernel(const PackJob* __restrict__ jobs, int njobs, bool tf32) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {                     // last job with block0 <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block0 <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const PackJob jb = jobs[lo];
  const size_t i = static_cast<size_t>(blockIdx.x - jb.block0) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(jb.R) * jb.Cn * jb.J;
  if (i >= total) return;
  const int c = static_cast<int>(i % jb.Cn);
  const int j = static_cast<int>((i / jb.Cn) % jb.J);
  const int r = static_cast<int>(i / (static_cast<size_t>(jb.Cn) * jb.J));
  float v = jb.src[r * jb.sr + c * jb.sc + j * jb.sj];
  if (jb.add) v += jb.add[r * jb.sr + c * jb.sc + j * jb.sj];
  if (tf32) v = ptx::round_tf32(v);
  TOp* p = static_cast<TOp*>(jb.dst) + r * jb.ldd + jb.col0 + j * jb.cblk + c;
  if constexpr (std::is_same<TOp, float>::value) *p = v;
  else *p = __float2bfloat16_rn(v);
}

}  // namespace fse

using namespace fse;

// ------------------------------------------------------------------ handle
struct fse_trainer {
  fse_denoiser_config cfg{};
  std::vector<PackJob> pack_host, pack_fp32_host;     // repacking jobs of the last load (re-uploaded only when a pointer changed)
  PackJob* pack_dev = nullptr; PackJob* pack_fp32_dev = nullptr;
  int pack_blocks = 0, pack_fp32_blocks = 0;
  bool bf16 = true, tc = true;
  int KB = 64;
  bool loaded = false;
  // packed operands (device; rewritten by fse_train_load_weights_device)
  void* W_in = nullptr; int Kp_in = 0;        // [C, Kp_in]
  void* Wy = nullptr;                          // [L][2C][3C + H]   rows: gate 0..C-1, filter C..2C-1 (diffnet.py:76 chunk order)
  void* Wo = nullptr;                          // [L][2C][C]
  void* W_skip = nullptr; void* W_out = nullptr;   // [C][C], [M][C]
  void* WoutT = nullptr; void* WskipT = nullptr;   // [C][Kp_in], [C][C]
  void* WoT = nullptr;                         // [L][C][2C]
  void* WyT = nullptr;                         // [L][C][3 * 2C]   tap j holds W_dc[:, :, j]^T; run with offsets -off_j
  void* WcpT = nullptr;                        // [H][L * 2C]
  float* b_y = nullptr;                        // [L][2C] = dilated_conv.bias + conditioner_projection.bias
  const float *b_in = nullptr, *b_skip = nullptr, *b_out = nullptr;   // the caller's parameter tensors (device), valid until the next load
  std::vector<const float*> b_op;              // [L] -> [2C]
  CUtensorMap mW_in{}, mW_skip{}, mW_out{}, mWoutT{}, mWskipT{}, mWcpT{};
  std::vector<CUtensorMap> mWy, mWo, mWoT, mWyT;
  struct Plan {
    const void* ws = nullptr; const void* cond = nullptr; int B = 0, T = 0;
    CUtensorMap m_x{}, m_cond{}, m_s{}, m_r{}, m_dx{}, m_dz{}, m_dS{}, m_dy{};
    std::vector<CUtensorMap> m_hin, m_u, m_dres;
  } plan;
  long long launches = 0;
};

namespace {

struct TrainWs {
  void* x_rows; void* h0_op; float* h; float* S; float* y; void* hin; void* sg; void* tf; void* u; void* s_op; void* r_op;
  void* condb; void* dx_rows; void* dz; void* dS; float* dh; void* dres; void* dy;
  size_t bytes;
  size_t off[20];
};

TrainWs tcarve(const fse_trainer* h, void* base, int B, int T) {
  const size_t N = static_cast<size_t>(B) * T;
  const int C = h->cfg.channels, H = h->cfg.hidden, M = h->cfg.n_mels, L = h->cfg.layers;
  const size_t es = h->bf16 ? 2 : 4;
  size_t off = 0;
  int k = 0;
  TrainWs w{};
  uint8_t* p = static_cast<uint8_t*>(base);
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); w.off[k++] = o; return p + o; };
  w.x_rows = take(N * M * es);                                   // 0  x_t as rows [B*T, M]                 (operand type)
  w.h0_op = take(N * C * es);                                    // 1  h_0 = relu(input_projection(x))      (operand type)
  w.h = reinterpret_cast<float*>(take(N * C * 4));               // 2  residual stream (fp32; final value after the forward)
  w.S = reinterpret_cast<float*>(take(N * C * 4));               // 3  skip sum (fp32)
  w.y = reinterpret_cast<float*>(take(N * 2 * C * 4));           // 4  scratch: pre-activation of the current layer
  w.hin = take(N * C * es * L);                                  // 5  [L][B*T, C] h + d_l                   (operand type)
  w.sg = take(N * C * es * L);                                   // 6  [L][B*T, C] du/dg = tanh(f) sg (1 - sg)
  w.tf = take(N * C * es * L);                                   // 7  [L][B*T, C] du/df = sg (1 - tanh(f)^2)
  w.u = take(N * C * es * L);                                    // 8  [L][B*T, C] gate output
  w.s_op = take(N * C * es);                                     // 9  S / sqrt(L)
  w.r_op = take(N * C * es);                                     // 10 relu(skip_projection(.))
  w.condb = take(h->bf16 ? N * H * 2 : 0);                       // 11 bf16 copy of cond (bf16 mode)
  w.dx_rows = take(N * M * es);                                  // 12 dx0 as rows
  w.dz = take(N * C * es);                                       // 13
  w.dS = take(N * C * es);                                       // 14 gradient of every layer's skip output
  w.dh = reinterpret_cast<float*>(take(N * C * 4));              // 15 gradient of the residual stream (after backward: of h_0)
  w.dres = take(N * C * es * L);                                 // 16 [L][B*T, C] gradient of each layer's residual output
  w.dy = take(N * 2 * C * es * L);                               // 17 [B*T, L * 2C] gradient of every layer's pre-activation
  w.bytes = off;
  return w;
}

int train_plan(fse_trainer* h, const TrainWs& w, const void* ws, const void* cond, int B, int T) {
  if (!h->tc) return FSE_OK;
  auto& pl = h->plan;
  if (pl.ws == ws && pl.B == B && pl.T == T && pl.cond == cond) return FSE_OK;
  const int C = h->cfg.channels, H = h->cfg.hidden, M = h->cfg.n_mels, L = h->cfg.layers, KB = h->KB, es = h->bf16 ? 2 : 4;
  const size_t N = static_cast<size_t>(B) * T;
  FSE_TRY(make_map_act(&pl.m_x, w.x_rows, M, T, B, KB, kTileM, es));
  FSE_TRY(make_map_act(&pl.m_cond, h->bf16 ? w.condb : cond, H, T, B, KB, kTileM, es));
  FSE_TRY(make_map_act(&pl.m_s, w.s_op, C, T, B, KB, kTileM, es));
  FSE_TRY(make_map_act(&pl.m_r, w.r_op, C, T, B, KB, kTileM, es));
  FSE_TRY(make_map_act(&pl.m_dx, w.dx_rows, M, T, B, KB, kTileM, es));
  FSE_TRY(make_map_act(&pl.m_dz, w.dz, C, T, B, KB, kTileM, es));
  FSE_TRY(make_map_act(&pl.m_dS, w.dS, C, T, B, KB, kTileM, es));
  FSE_TRY(make_map_act(&pl.m_dy, w.dy, L * 2 * C, T, B, KB, kTileM, es));
  pl.m_hin.resize(L); pl.m_u.resize(L); pl.m_dres.resize(L);
  for (int l = 0; l < L; ++l) {
    const size_t o = static_cast<size_t>(l) * N * C * es;
    FSE_TRY(make_map_act(&pl.m_hin[l], static_cast<uint8_t*>(w.hin) + o, C, T, B, KB, kTileM, es));
    FSE_TRY(make_map_act(&pl.m_u[l], static_cast<uint8_t*>(w.u) + o, C, T, B, KB, kTileM, es));
    FSE_TRY(make_map_act(&pl.m_dres[l], static_cast<uint8_t*>(w.dres) + o, C, T, B, KB, kTileM, es));
  }
  pl.ws = ws; pl.B = B; pl.T = T; pl.cond = cond;
  return FSE_OK;
}

template <typename TOp>
int train_forward_impl(fse_trainer* h, const float* x_t, const float* cond, const float* d, float* x0, int B, int T, void* ws, cudaStream_t st) {
  TrainWs w = tcarve(h, ws, B, T);
  FSE_TRY(train_plan(h, w, ws, cond, B, T));
  const int C = h->cfg.channels, H = h->cfg.hidden, M = h->cfg.n_mels, L = h->cfg.layers, KB = h->KB, mode = h->cfg.mode;
  const size_t N = static_cast<size_t>(B) * T, es = sizeof(TOp);
  const bool fast = mode == FSE_MODE_TC_BF16;
  const bool rnd = mode == FSE_MODE_TC_TF32;
  const int zero = 0;
  LaunchCtx ctx{&h->launches, nullptr, 0};
  const void* cond_op = cond;
  if constexpr (std::is_same<TOp, __nv_bfloat16>::value) {
    const size_t n = N * H;
    f32_to_bf16_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, st>>>(cond, static_cast<__nv_bfloat16*>(w.condb), n / 4);
    FSE_CUDA(cudaGetLastError());
    cond_op = w.condb;
  }
  x_to_rows_kernel<TOp><<<dim3((T + 255) / 256, B), 256, 0, st>>>(x_t, static_cast<TOp*>(w.x_rows), M, T);
  FSE_CUDA(cudaGetLastError());
  FSE_CUDA(cudaMemsetAsync(w.S, 0, N * C * 4, st));
  {  // h_0 = relu(input_projection(x))   (diffnet.py:117-120)
    ConvGemmParams p = make_params(B, T, T, M, 1, &zero, 0, C, KB);
    GemmOperands op; op.A0 = w.x_rows; op.W = h->W_in; op.mA0 = &h->plan.m_x; op.mW = &h->mW_in; op.BN = 256;
    EpiIn<TOp> epi{h->b_in, w.h, static_cast<TOp*>(w.h0_op), C, T};
    FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, ctx)));
  }
  const unsigned eb = static_cast<unsigned>((N * C / 4 + 255) / 256);
  for (int l = 0; l < L; ++l) {
    const size_t lo = static_cast<size_t>(l) * N * C * es;
    TOp* hin = reinterpret_cast<TOp*>(static_cast<uint8_t*>(w.hin) + lo);
    TOp* sg = reinterpret_cast<TOp*>(static_cast<uint8_t*>(w.sg) + lo);
    TOp* tf = reinterpret_cast<TOp*>(static_cast<uint8_t*>(w.tf) + lo);
    TOp* u = reinterpret_cast<TOp*>(static_cast<uint8_t*>(w.u) + lo);
    add_bcast_cast_kernel<TOp><<<eb, 256, 0, st>>>(w.h, d + static_cast<size_t>(l) * B * C, hin, T, C, N * C / 4, rnd);
    FSE_CUDA(cudaGetLastError());
    const int dil = 1 << (l % h->cfg.dilation_cycle_length);
    const int offs[3] = {-dil, 0, dil};
    {
      ConvGemmParams p = make_params(B, T, T, C, 3, offs, H, 2 * C, KB);
      GemmOperands op; op.A0 = hin; op.A1 = cond_op; op.W = static_cast<uint8_t*>(h->Wy) + static_cast<size_t>(l) * 2 * C * (3 * C + H) * es;
      op.mA0 = &h->plan.m_hin[l]; op.mA1 = &h->plan.m_cond; op.mW = h->tc ? &h->mWy[l] : nullptr; op.BN = 256;
      EpiStoreF32<TOp> epi{h->b_y + static_cast<size_t>(l) * 2 * C, nullptr, w.y, 2 * C, T};
      FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, ctx)));
    }
    if (fast) gate_fwd_kernel<TOp, true><<<eb, 256, 0, st>>>(w.y, sg, tf, u, C, N * C / 4, rnd);
    else gate_fwd_kernel<TOp, false><<<eb, 256, 0, st>>>(w.y, sg, tf, u, C, N * C / 4, rnd);
    FSE_CUDA(cudaGetLastError());
    {
      ConvGemmParams p = make_params(B, T, T, C, 1, &zero, 0, 2 * C, KB);
      GemmOperands op; op.A0 = u; op.W = static_cast<uint8_t*>(h->Wo) + static_cast<size_t>(l) * 2 * C * C * es;
      op.mA0 = &h->plan.m_u[l]; op.mW = h->tc ? &h->mWo[l] : nullptr; op.BN = 256;
      if (fast) {
        EpiResSkip<TOp, true> epi{h->b_op[l], w.h, w.S, C, T};
        FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, ctx)));
      } else {
        EpiResSkip<TOp, false> epi{h->b_op[l], w.h, w.S, C, T};
        FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, ctx)));
      }
    }
    h->launches += 2;
  }
  scale_cast_kernel<TOp><<<eb, 256, 0, st>>>(w.S, static_cast<TOp*>(w.s_op), static_cast<float>(1.0 / std::sqrt(static_cast<double>(L))), N * C / 4, rnd);
  FSE_CUDA(cudaGetLastError());
  {  // r = relu(skip_projection(S / sqrt(L)))   (diffnet.py:128-130)
    ConvGemmParams p = make_params(B, T, T, C, 1, &zero, 0, C, KB);
    GemmOperands op; op.A0 = w.s_op; op.W = h->W_skip; op.mA0 = &h->plan.m_s; op.mW = &h->mW_skip; op.BN = 256;
    EpiSkip<TOp> epi{h->b_skip, static_cast<TOp*>(w.r_op), C, T};
    FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, ctx)));
  }
  {  // x0 = output_projection(r)   (diffnet.py:131) -> [B, M, T]
    ConvGemmParams p = make_params(B, T, T, C, 1, &zero, 0, M, KB);
    GemmOperands op; op.A0 = w.r_op; op.W = h->W_out; op.mA0 = &h->plan.m_r; op.mW = &h->mW_out; op.BN = M;
    EpiOut<TOp> epi{h->b_out, M, T, 0, nullptr, x0, nullptr, nullptr, nullptr, 0u, 0.f, 0.f, 0.f, nullptr, nullptr, nullptr};
    FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, ctx)));
  }
  return FSE_OK;
}

template <typename TOp>
int train_backward_impl(fse_trainer* h, const float* dx0, float* dcond, int B, int T, void* ws, cudaStream_t st) {
  TrainWs w = tcarve(h, ws, B, T);
  if (h->tc && !(h->plan.ws == ws && h->plan.B == B && h->plan.T == T)) return fail(FSE_ESTATE, "fse_train_backward: no forward ran on this workspace / shape");
  const int C = h->cfg.channels, H = h->cfg.hidden, M = h->cfg.n_mels, L = h->cfg.layers, KB = h->KB, mode = h->cfg.mode;
  const size_t N = static_cast<size_t>(B) * T, es = sizeof(TOp);
  const int zero = 0;
  const bool rnd = mode == FSE_MODE_TC_TF32;
  LaunchCtx ctx{&h->launches, nullptr, 0};
  x_to_rows_kernel<TOp><<<dim3((T + 255) / 256, B), 256, 0, st>>>(dx0, static_cast<TOp*>(w.dx_rows), M, T);
  FSE_CUDA(cudaGetLastError());
  {  // dz = (W_out^T dx0) * [r > 0]
    ConvGemmParams p = make_params(B, T, T, M, 1, &zero, 0, C, KB);
    GemmOperands op; op.A0 = w.dx_rows; op.W = h->WoutT; op.mA0 = &h->plan.m_dx; op.mW = &h->mWoutT; op.BN = 256;
    EpiReluBwd<TOp> epi{static_cast<const TOp*>(w.r_op), static_cast<TOp*>(w.dz), C, T, rnd};
    FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, ctx)));
  }
  {  // dS = W_skip^T dz / sqrt(L): the gradient of EVERY layer's skip output
    ConvGemmParams p = make_params(B, T, T, C, 1, &zero, 0, C, KB);
    GemmOperands op; op.A0 = w.dz; op.W = h->WskipT; op.mA0 = &h->plan.m_dz; op.mW = &h->mWskipT; op.BN = 256;
    EpiScaleCast<TOp> epi{static_cast<TOp*>(w.dS), C, T, static_cast<float>(1.0 / std::sqrt(static_cast<double>(L))), rnd};
    FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, ctx)));
  }
  FSE_CUDA(cudaMemsetAsync(w.dh, 0, N * C * 4, st));                                          // nothing reads h_L
  FSE_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(w.dres) + static_cast<size_t>(L - 1) * N * C * es, 0, N * C * es, st));
  for (int l = L - 1; l >= 0; --l) {
    const size_t lo = static_cast<size_t>(l) * N * C * es;
    const TOp* sg = reinterpret_cast<const TOp*>(static_cast<uint8_t*>(w.sg) + lo);
    const TOp* tf = reinterpret_cast<const TOp*>(static_cast<uint8_t*>(w.tf) + lo);
    {  // du = W_op^T [dres_l | dS]  ->  dy_l = [dg | df]
      ConvGemmParams p = make_params(B, T, T, C, 1, &zero, C, C, KB);
      GemmOperands op; op.A0 = static_cast<uint8_t*>(w.dres) + lo; op.A1 = w.dS;
      op.W = static_cast<uint8_t*>(h->WoT) + static_cast<size_t>(l) * C * 2 * C * es;
      op.mA0 = &h->plan.m_dres[l]; op.mA1 = &h->plan.m_dS; op.mW = h->tc ? &h->mWoT[l] : nullptr; op.BN = 256;
      EpiGateBwd<TOp> epi{sg, tf, static_cast<TOp*>(w.dy), C, T, L * 2 * C, l * 2 * C, rnd};
      FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, ctx)));
    }
    {  // dh <- dh / sqrt(2) + conv_k3^T(dy_l); the layer below gets dres = dh / sqrt(2)
      const int dil = 1 << (l % h->cfg.dilation_cycle_length);
      const int offs[3] = {dil, 0, -dil};                       // tap j of the forward read hin[t + off_j]: its transpose reads dy[t - off_j]
      ConvGemmParams p = make_params(B, T, T, 2 * C, 3, offs, 0, C, KB);
      p.c_off0 = l * 2 * C; p.ld0 = L * 2 * C;
      GemmOperands op; op.A0 = w.dy; op.W = static_cast<uint8_t*>(h->WyT) + static_cast<size_t>(l) * C * 6 * C * es;
      op.mA0 = &h->plan.m_dy; op.mW = h->tc ? &h->mWyT[l] : nullptr; op.BN = 256;
      EpiDh<TOp> epi{w.dh, l > 0 ? reinterpret_cast<TOp*>(static_cast<uint8_t*>(w.dres) + lo - N * C * es) : nullptr, C, T, rnd};
      FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, ctx)));
    }
  }
  {  // dcond = sum_l W_cp,l^T dy_l: one GEMM over the dy of all layers (K = L * 2C)
    ConvGemmParams p = make_params(B, T, T, L * 2 * C, 1, &zero, 0, H, KB);
    GemmOperands op; op.A0 = w.dy; op.W = h->WcpT; op.mA0 = &h->plan.m_dy; op.mW = &h->mWcpT; op.BN = H;
    EpiStoreF32<TOp> epi{nullptr, nullptr, dcond, H, T};
    FSE_TRY((run_conv_gemm<TOp>(mode, p, op, epi, st, ctx)));
  }
  return FSE_OK;
}

template <typename TOp>
int train_pack(fse_trainer* h, const TensorTable& tt, cudaStream_t st) {
  const int C = h->cfg.channels, H = h->cfg.hidden, M = h->cfg.n_mels, L = h->cfg.layers;
  const bool tf32 = h->cfg.mode == FSE_MODE_TC_TF32;
  int rc = FSE_OK;
  auto G = [&](const std::string& name, int64_t numel) { return rc == FSE_OK ? tt.get(name, numel, &rc) : nullptr; };
  // the repacking jobs are collected and run as ONE launch (plus one for the fp32 bias sums) at the end of this function
  std::vector<PackJob> jobs, jobs32;
  int blocks = 0, blocks32 = 0;
  auto add_job = [](std::vector<PackJob>& v, int& nb, void* dst, long ldd, long col0, const float* src, int R, int Cn, int J, long sr, long sc, long sj, long cblk,
                    const float* add = nullptr) {
    const size_t total = static_cast<size_t>(R) * Cn * J;
    PackJob jb{};                                           // zero-initialised incl. padding: the tables are compared with memcmp
    jb.dst = dst; jb.src = src; jb.add = add; jb.ldd = ldd; jb.col0 = col0; jb.sr = sr; jb.sc = sc; jb.sj = sj; jb.cblk = cblk;
    jb.R = R; jb.Cn = Cn; jb.J = J; jb.block0 = nb;
    v.push_back(jb);
    nb += static_cast<int>((total + 255) / 256);
  };
  auto pack = [&](void* dst, long ldd, long col0, const float* src, int R, int Cn, int J, long sr, long sc, long sj, long cblk) {
    add_job(jobs, blocks, dst, ldd, col0, src, R, Cn, J, sr, sc, sj, cblk);
  };
  const float* w_in = G("input_projection.weight", (int64_t)C * M);
  h->b_in = G("input_projection.bias", C);
  const float* w_skip = G("skip_projection.weight", (int64_t)C * C);
  h->b_skip = G("skip_projection.bias", C);
  const float* w_out = G("output_projection.weight", (int64_t)M * C);
  h->b_out = G("output_projection.bias", M);
  if (rc) return rc;
  pack(h->W_in, h->Kp_in, 0, w_in, C, M, 1, M, 1, 0, 0);                 // [C, M] -> [C, Kp_in]
  pack(h->WoutT, h->Kp_in, 0, w_out, C, M, 1, 1, C, 0, 0);               // W_out^T: dst[c, m] = w_out[m, c]
  pack(h->W_skip, C, 0, w_skip, C, C, 1, C, 1, 0, 0);
  pack(h->WskipT, C, 0, w_skip, C, C, 1, 1, C, 0, 0);
  pack(h->W_out, C, 0, w_out, M, C, 1, C, 1, 0, 0);
  const long Ky = 3 * C + H;
  h->b_op.resize(L);
  for (int l = 0; l < L; ++l) {
    const std::string pre = "residual_layers." + std::to_string(l) + ".";
    const float* wdc = G(pre + "dilated_conv.weight", (int64_t)2 * C * C * 3);
    const float* bdc = G(pre + "dilated_conv.bias", 2 * C);
    const float* wcp = G(pre + "conditioner_projection.weight", (int64_t)2 * C * H);
    const float* bcp = G(pre + "conditioner_projection.bias", 2 * C);
    const float* wop = G(pre + "output_projection.weight", (int64_t)2 * C * C);
    h->b_op[l] = G(pre + "output_projection.bias", 2 * C);
    if (rc) return rc;
    TOp* Wy = static_cast<TOp*>(h->Wy) + static_cast<size_t>(l) * 2 * C * Ky;
    pack(Wy, Ky, 0, wdc, 2 * C, C, 3, (long)C * 3, 3, 1, C);            // dst[n, j*C + c] = wdc[n, c, j]
    pack(Wy, Ky, 3 * C, wcp, 2 * C, H, 1, H, 1, 0, 0);                   // dst[n, 3C + h] = wcp[n, h]
    pack(static_cast<TOp*>(h->Wo) + static_cast<size_t>(l) * 2 * C * C, C, 0, wop, 2 * C, C, 1, C, 1, 0, 0);
    pack(static_cast<TOp*>(h->WoT) + static_cast<size_t>(l) * C * 2 * C, 2 * C, 0, wop, C, 2 * C, 1, 1, C, 0, 0);     // dst[c, n] = wop[n, c]
    pack(static_cast<TOp*>(h->WyT) + static_cast<size_t>(l) * C * 6 * C, 6 * C, 0, wdc, C, 2 * C, 3, 3, (long)C * 3, 1, 2 * C);   // dst[c, j*2C + n] = wdc[n, c, j]
    pack(h->WcpT, (long)L * 2 * C, (long)l * 2 * C, wcp, H, 2 * C, 1, 1, H, 0, 0);                                       // dst[h, l*2C + n] = wcp[n, h]
    // b_y = dilated_conv.bias + conditioner_projection.bias (one job with two sources)
    add_job(jobs32, blocks32, h->b_y + static_cast<size_t>(l) * 2 * C, 0, 0, bdc, 1, 2 * C, 1, 0, 1, 0, 0, bcp);
  }
  auto same = [](const std::vector<PackJob>& a, const std::vector<PackJob>& b) {
    return a.size() == b.size() && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * sizeof(PackJob)) == 0);
  };
  auto upload = [&](const std::vector<PackJob>& v, std::vector<PackJob>& host, PackJob*& dev) -> int {
    if (dev && same(v, host)) return FSE_OK;              // same parameter tensors as last step: the table on the device is current
    if (dev && host.size() != v.size()) { FSE_CUDA(cudaFree(dev)); dev = nullptr; }
    if (!dev) FSE_CUDA(cudaMalloc(reinterpret_cast<void**>(&dev), v.size() * sizeof(PackJob)));
    FSE_CUDA(cudaStreamSynchronize(st));                    // a previous launch may still read the old table / the host copy is reused
    host = v;
    FSE_CUDA(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(PackJob), cudaMemcpyHostToDevice, st));
    return FSE_OK;
  };
  FSE_TRY(upload(jobs, h->pack_host, h->pack_dev));
  FSE_TRY(upload(jobs32, h->pack_fp32_host, h->pack_fp32_dev));
  h->pack_blocks = blocks; h->pack_fp32_blocks = blocks32;
  pack_jobs_kernel<TOp><<<blocks, 256, 0, st>>>(h->pack_dev, static_cast<int>(jobs.size()), tf32);
  pack_jobs_kernel<float><<<blocks32, 256, 0, st>>>(h->pack_fp32_dev, static_cast<int>(jobs32.size()), false);
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

}  // namespace

extern "C" {

int fse_train_create(const fse_denoiser_config* cfg, fse_trainer** out) {
  if (!cfg || !out) return fail(FSE_EINVAL, "null argument");
  if (cfg->n_mels <= 0 || cfg->n_mels % 16 != 0) return fail(FSE_EINVAL, "n_mels must be a positive multiple of 16");
  if (cfg->channels != 256) return fail(FSE_EINVAL, "the training path is built for 256 residual channels (got %d)", cfg->channels);
  if (cfg->hidden <= 0 || cfg->hidden % 64 != 0 || cfg->hidden > 256) return fail(FSE_EINVAL, "hidden must be a multiple of 64, <= 256");
  if (cfg->layers <= 0 || cfg->dilation_cycle_length <= 0) return fail(FSE_EINVAL, "layers / dilation_cycle_length must be positive");
  if (cfg->mode != FSE_MODE_TC_BF16 && cfg->mode != FSE_MODE_SIMT_F32 && cfg->mode != FSE_MODE_TC_TF32) return fail(FSE_EINVAL, "mode must be tc_bf16, tc_tf32 or simt_f32");
  FSE_TRY(check_device());
  auto* h = new fse_trainer();
  h->cfg = *cfg;
  h->bf16 = mode_is_bf16(cfg->mode); h->tc = mode_is_tc(cfg->mode); h->KB = mode_kb(cfg->mode);
  const int C = cfg->channels, H = cfg->hidden, M = cfg->n_mels, L = cfg->layers, KB = h->KB;
  const size_t es = h->bf16 ? 2 : 4;
  h->Kp_in = (M + KB - 1) / KB * KB;
  struct Alloc { void** p; size_t n; };
  const Alloc allocs[] = {{&h->W_in, (size_t)C * h->Kp_in}, {&h->WoutT, (size_t)C * h->Kp_in}, {&h->W_skip, (size_t)C * C}, {&h->WskipT, (size_t)C * C},
                          {&h->W_out, (size_t)M * C}, {&h->Wy, (size_t)L * 2 * C * (3 * C + H)}, {&h->Wo, (size_t)L * 2 * C * C},
                          {&h->WoT, (size_t)L * C * 2 * C}, {&h->WyT, (size_t)L * C * 6 * C}, {&h->WcpT, (size_t)H * L * 2 * C}};
  for (const auto& a : allocs) {
    if (cudaMalloc(a.p, a.n * es) != cudaSuccess || cudaMemset(*a.p, 0, a.n * es) != cudaSuccess) { fse_train_destroy(h); return fail(FSE_ECUDA, "cudaMalloc failed (trainer operands)"); }
  }
  if (cudaMalloc(reinterpret_cast<void**>(&h->b_y), (size_t)L * 2 * C * 4) != cudaSuccess) { fse_train_destroy(h); return fail(FSE_ECUDA, "cudaMalloc failed"); }
  if (h->tc) {
    const int ies = static_cast<int>(es);
    int rc = FSE_OK;
    auto M_ = [&](CUtensorMap* m, void* p, int Kp, int N, int BN) { if (rc == FSE_OK) rc = make_map_w(m, p, Kp, N, KB, BN, ies); };
    M_(&h->mW_in, h->W_in, h->Kp_in, C, 256); M_(&h->mWoutT, h->WoutT, h->Kp_in, C, 256);
    M_(&h->mW_skip, h->W_skip, C, C, 256); M_(&h->mWskipT, h->WskipT, C, C, 256);
    M_(&h->mW_out, h->W_out, C, M, M); M_(&h->mWcpT, h->WcpT, L * 2 * C, H, H);
    h->mWy.resize(L); h->mWo.resize(L); h->mWoT.resize(L); h->mWyT.resize(L);
    for (int l = 0; l < L; ++l) {
      M_(&h->mWy[l], static_cast<uint8_t*>(h->Wy) + (size_t)l * 2 * C * (3 * C + H) * es, 3 * C + H, 2 * C, 256);
      M_(&h->mWo[l], static_cast<uint8_t*>(h->Wo) + (size_t)l * 2 * C * C * es, C, 2 * C, 256);
      M_(&h->mWoT[l], static_cast<uint8_t*>(h->WoT) + (size_t)l * C * 2 * C * es, 2 * C, C, 256);
      M_(&h->mWyT[l], static_cast<uint8_t*>(h->WyT) + (size_t)l * C * 6 * C * es, 6 * C, C, 256);
    }
    if (rc != FSE_OK) { fse_train_destroy(h); return rc; }
  }
  *out = h;
  return FSE_OK;
}

void fse_train_destroy(fse_trainer* h) {
  if (!h) return;
  void* ptrs[] = {h->W_in, h->WoutT, h->W_skip, h->WskipT, h->W_out, h->Wy, h->Wo, h->WoT, h->WyT, h->WcpT, h->b_y};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (h->pack_dev) cudaFree(h->pack_dev);
  if (h->pack_fp32_dev) cudaFree(h->pack_fp32_dev);
  delete h;
}

int fse_train_load_weights_device(fse_trainer* h, const fse_tensor* tensors, int32_t n, void* stream) {
  if (!h || !tensors || n <= 0) return fail(FSE_EINVAL, "null argument");
  TensorTable tt(tensors, n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rc = h->bf16 ? train_pack<__nv_bfloat16>(h, tt, st) : train_pack<float>(h, tt, st);
  if (rc != FSE_OK) return rc;
  h->loaded = true;
  return FSE_OK;
}

int64_t fse_train_workspace_bytes(const fse_trainer* h, int32_t B, int32_t T) {
  if (!h || B <= 0 || T <= 0) return 0;
  return static_cast<int64_t>(tcarve(h, nullptr, B, T).bytes);
}

int fse_train_layout(const fse_trainer* h, int32_t B, int32_t T, int64_t* offsets, int32_t n) {
  if (!h || !offsets || B <= 0 || T <= 0 || n < 18) return fail(FSE_EINVAL, "bad argument (need room for 18 offsets)");
  TrainWs w = tcarve(h, nullptr, B, T);
  for (int i = 0; i < 18; ++i) offsets[i] = static_cast<int64_t>(w.off[i]);
  return FSE_OK;
}

static int train_validate(const fse_trainer* h, int B, int T, const void* ws, int64_t ws_bytes) {
  if (!h) return fail(FSE_EINVAL, "null handle");
  if (!h->loaded) return fail(FSE_ESTATE, "weights not loaded");
  if (B <= 0 || T <= 0) return fail(FSE_EINVAL, "B and T must be positive");
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return fail(FSE_EINVAL, "workspace must be non-null and 1024-byte aligned");
  if (ws_bytes < fse_train_workspace_bytes(h, B, T)) return fail(FSE_EINVAL, "workspace too small");
  if ((static_cast<size_t>(B) * T * h->cfg.hidden) % 4 != 0) return fail(FSE_EINVAL, "B*T*hidden must be a multiple of 4");
  return FSE_OK;
}

int fse_train_forward(fse_trainer* h, const float* x_t, const float* cond, const float* d, float* x0, int32_t B, int32_t T, void* workspace,
                      int64_t workspace_bytes, void* stream) {
  FSE_TRY(train_validate(h, B, T, workspace, workspace_bytes));
  if (!x_t || !cond || !d || !x0) return fail(FSE_EINVAL, "null tensor argument");
  h->launches = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return h->bf16 ? train_forward_impl<__nv_bfloat16>(h, x_t, cond, d, x0, B, T, workspace, st)
                 : train_forward_impl<float>(h, x_t, cond, d, x0, B, T, workspace, st);
}

int fse_train_backward(fse_trainer* h, const float* dx0, float* dcond, int32_t B, int32_t T, void* workspace, int64_t workspace_bytes, void* stream) {
  FSE_TRY(train_validate(h, B, T, workspace, workspace_bytes));
  if (!dx0 || !dcond) return fail(FSE_EINVAL, "null tensor argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return h->bf16 ? train_backward_impl<__nv_bfloat16>(h, dx0, dcond, B, T, workspace, st)
                 : train_backward_impl<float>(h, dx0, dcond, B, T, workspace, st);
}

int64_t fse_train_last_launches(const fse_trainer* h) { return h ? h->launches : 0; }

}  // extern "C"
