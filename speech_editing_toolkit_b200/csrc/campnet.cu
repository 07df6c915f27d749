// CampNet mask-predict forward (BASELINE.json configs[3]) behind the C ABI of include/fse_b200.h.
// Reference (file:line relative to the reference tree):
//   modules/speech_editing/campnet/campnet.py:40-69              CampNet.forward / run_text_encoder / run_decoder
//   modules/speech_editing/commons/transformer.py:14-71          SinusoidalPositionalEmbedding, utils/nn/seq_utils.py:6-18 make_positions
//   modules/speech_editing/commons/transformer.py:74-110         TransformerFFNLayer (conv k9 'SAME' | 'LEFT' -> * k^-0.5 -> GELU -> Linear)
//   modules/speech_editing/commons/transformer.py:138-419        MultiheadAttention (2 heads x 96, bias=False, fp32 softmax)
//   modules/speech_editing/commons/transformer.py:489-608        EncSALayer / DecSALayer
//   modules/speech_editing/commons/transformer.py:639-812        FFTBlocks / TransformerEncoder / TransformerDecoder
//   modules/commons/conv.py:68-116 (decoder_fine), modules/speech_editing/commons/mel_encoder.py:3-19 (MelEncoder)
//
// Layout: channels-last rows (a token / frame is a GEMM row).  Every projection, FFN convolution (9 taps = 9 row-shifted
// K-slices), ConvBlocks layer and the two 192 -> 80 output projections is a launch of the tensor-core conv-GEMM primitive
// (conv_gemm.cuh) with the layer's pointwise tail in its epilogue; LayerNorm is the warp-per-row kernel of rowwise.cuh.
// Attention (softmax(QK^T)V per item and head) is a flash-style tiled kernel with an online fp32 softmax.  In this first
// version its two contractions run on CUDA cores over the stored (bf16 / fp32) q, k, v: it is the kernel to move onto
// tcgen05 next (S tile in TMEM, P through shared memory as the A operand of the PV MMA); everything around it already is.
#include <cstdlib>

#include "rowwise.cuh"
#include "attention_tc.cuh"

namespace fse {
namespace {

constexpr int kAttTile = 64;       // queries per block and keys per tile of the CUDA-core attention kernel

// q/k/v projection without bias: columns < nscaled (the query part) are multiplied by head_dim^-0.5 (transformer.py:296).
// With vt != null the value columns (n >= vstart) are ALSO stored transposed, vt[(b*heads + head)*D + c][t] (row length Tkp):
// the K-major B operand of the tensor-core P.V contraction (attention_tc.cuh).
template <typename TOp>
struct EpiScaleCols {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  TOp* out;   // [B*T, N]
  int N, T, nscaled;
  float scale;
  TOp* vt;    // or null
  int vstart, Tkp, heads;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = n0 + i < nscaled ? __fmul_rn(acc[i], scale) : acc[i];
    st_vec<NV>(out + (static_cast<size_t>(b) * T + t) * N + n0, v);
    if (vt && n0 >= vstart) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = n0 + i - vstart;
        store_op(vt + (static_cast<size_t>(b) * heads * kAttD + c) * Tkp + t, v[i]);      // c = head * D + channel
      }
    }
  }
};

// post_net1 of decoder_fine: (conv + b) * nonpadding as the operand of the output projection
template <typename TOp>
struct EpiBiasMaskOp {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  const float* bias;
  const float* mask;   // [B*T]
  TOp* out;            // [B*T, N]
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

// the two 192 -> 80 projections (no bias) with the compositing that follows them (campnet.py:58-68):
//   y = gemm * nonpad;  raw (optional) = y;  comp = base * (one_minus ? 1 - m : 1) + y * m
//   coarse: base = mels, one_minus: mel_coarse = mels (1 - m) + mel_out_coarse m;  fine: base = mel_coarse: mel_out_fine = mel_coarse + y m
struct EpiMelOut {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  const float* nonpad;   // [B*T]
  const float* m;        // [B*T] time_mel_masks
  const float* base;     // [B*T, N]
  float* raw;            // [B*T, N] or null
  float* comp;           // [B*T, N]
  int N, T, one_minus;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    const size_t row = static_cast<size_t>(b) * T + t;
    const float np = __ldg(nonpad + row), mm = __ldg(m + row);
    const float bs = one_minus ? __fsub_rn(1.f, mm) : 1.f;
    float y[NV], c[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      y[i] = __fmul_rn(acc[i], np);
      c[i] = __fadd_rn(__fmul_rn(__ldg(base + row * N + n0 + i), bs), __fmul_rn(y[i], mm));
    }
    if (raw) st_vec<NV>(raw + row * N + n0, y);
    st_vec<NV>(comp + row * N + n0, c);
  }
};

// make_positions (utils/nn/seq_utils.py:6-18): non-padding symbols are numbered 1, 2, ... per item, padding keeps 0.
// flags: txt (int64, != 0) or a float 0/1 vector; one thread per item (runs twice per forward).
__global__ void camp_positions_kernel(const int64_t* __restrict__ txt, const float* __restrict__ flag, int* __restrict__ pos, int B, int T) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int n = 0;
  for (int t = 0; t < T; ++t) {
    const size_t i = static_cast<size_t>(b) * T + t;
    const bool on = txt ? txt[i] != 0 : flag[i] != 0.f;
    if (on) ++n;
    pos[i] = on ? n : 0;
  }
}

// encoder input: x = (embed_scale * embed_tokens(txt) + positions) * (txt != 0)   (transformer.py:749-756, :680); keep = (txt != 0)
__global__ void __launch_bounds__(256) camp_embed_kernel(const int64_t* __restrict__ txt, const int* __restrict__ pos, const float* __restrict__ table,
                                                         const float* __restrict__ sinus, float* __restrict__ x, float* __restrict__ keep,
                                                         int rows, int C, int vocab, int max_pos, float scale) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  long long tok = txt[row];
  const float k = tok != 0 ? 1.f : 0.f;
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
  int p = pos[row];
  p = p >= max_pos ? max_pos - 1 : p;
  for (int c = lane; c < C; c += 32)
    x[static_cast<size_t>(row) * C + c] = __fmul_rn(__fadd_rn(__fmul_rn(scale, __ldg(table + static_cast<size_t>(tok) * C + c)),
                                                              __ldg(sinus + static_cast<size_t>(p) * C + c)), k);
  if (lane == 0) keep[row] = k;
}

// decoder input: x = (x + alpha * positions) * keep   (transformer.py:786-791)
__global__ void __launch_bounds__(256) camp_add_pos_kernel(float* __restrict__ x, const int* __restrict__ pos, const float* __restrict__ sinus,
                                                           const float* __restrict__ keep, int rows, int C, int max_pos, float alpha) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  int p = pos[row];
  p = p >= max_pos ? max_pos - 1 : p;
  const float k = keep[row];
  for (int c = lane; c < C; c += 32) {
    const size_t i = static_cast<size_t>(row) * C + c;
    x[i] = __fmul_rn(__fadd_rn(x[i], __fmul_rn(alpha, __ldg(sinus + static_cast<size_t>(p) * C + c))), k);
  }
}

// mel_input_coarse = mels (1 - m) + mask_emb m  and  mel_nonpadding = (sum |mels| > 0)   (campnet.py:55-57)
__global__ void __launch_bounds__(256) camp_mel_input_kernel(const float* __restrict__ mels, const float* __restrict__ m,
                                                             const float* __restrict__ mask_emb, float* __restrict__ xin,
                                                             float* __restrict__ nonpad, int rows, int M) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float mm = m[row], om = __fsub_rn(1.f, mm);
  float sa = 0.f;
  for (int c = lane; c < M; c += 32) {
    const float v = mels[static_cast<size_t>(row) * M + c];
    sa += fabsf(v);
    xin[static_cast<size_t>(row) * M + c] = __fadd_rn(__fmul_rn(v, om), __fmul_rn(__ldg(mask_emb + c), mm));
  }
  sa = warp_sum(sa);
  if (lane == 0) nonpad[row] = sa > 0.f ? 1.f : 0.f;
}

// ---------------------------------------------------------------------------------------------- attention
// O[b, q, h*D:(h+1)*D] = softmax_k(Q_h[b,q] . K_h[b,k]) V_h[b,k], q already carries the head_dim^-0.5 scale (transformer.py:296,
// 365-407).  One block = 64 queries of one (item, head); keys / values stream through shared memory in tiles of 64 with an
// online fp32 softmax.  256 threads as a 16 x 16 grid: thread (ty, tx) owns score rows 4ty..4ty+3 x columns tx + 16j and
// output rows 4ty..4ty+3 x columns 6tx..6tx+5 (bank-conflict-free strides).  Padded keys (key_keep == 0) get -1e8 as in the
// reference (exp underflows to exactly 0); keys past Tk are excluded.
template <typename TOp>
__global__ void __launch_bounds__(256) camp_attention_kernel(const TOp* __restrict__ Q, int ldq, int qoff, const TOp* __restrict__ K,
                                                             const TOp* __restrict__ V, int ldkv, int koff, int voff,
                                                             const float* __restrict__ key_keep, TOp* __restrict__ O, int ldo, int Tq, int Tk,
                                                             float* __restrict__ probs, float probs_scale) {
  constexpr int D = kAttD, TQ = kAttTile, TK = kAttTile, QS = D + 1, PS = TK + 1, NC = D / 16;
  extern __shared__ float smem[];
  float* Qs = smem;                 // [TQ][QS]
  float* Ks = Qs + TQ * QS;         // [TK][QS]
  float* Vs = Ks + TK * QS;         // [TK][D]
  float* Ps = Vs + TK * D;          // [TQ][PS]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  for (int i = tid; i < TQ * D; i += 256) {
    const int r = i / D, c = i % D, q = q0 + r;
    Qs[r * QS + c] = q < Tq ? to_f32(Q[(static_cast<size_t>(b) * Tq + q) * ldq + qoff + h * D + c]) : 0.f;
  }
  float o[4][NC], mrow[4], lrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mrow[i] = -INFINITY; lrow[i] = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) o[i][c] = 0.f;
  }
  for (int k0 = 0; k0 < Tk; k0 += TK) {
    __syncthreads();                 // previous tile fully consumed (and Qs visible on the first pass)
    for (int i = tid; i < TK * D; i += 256) {
      const int r = i / D, c = i % D, k = k0 + r;
      const size_t g = (static_cast<size_t>(b) * Tk + (k < Tk ? k : 0)) * ldkv + h * D + c;
      Ks[r * QS + c] = k < Tk ? to_f32(K[g + koff]) : 0.f;
      Vs[r * D + c] = k < Tk ? to_f32(V[g + voff]) : 0.f;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int c = 0; c < D; ++c) {
      float a[4], kk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Qs[(4 * ty + i) * QS + c];
#pragma unroll
      for (int j = 0; j < 4; ++j) kk[j] = Ks[(tx + 16 * j) * QS + c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(a[i], kk[j], s[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx + 16 * j;
      const bool oob = k >= Tk;
      const bool padded = !oob && key_keep && key_keep[static_cast<size_t>(b) * Tk + k] == 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i][j] = oob ? -INFINITY : (padded ? -1e8f : s[i][j]);
    }
    float alpha[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = fmaxf(fmaxf(s[i][0], s[i][1]), fmaxf(s[i][2], s[i][3]));
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mnew = fmaxf(mrow[i], mx);                       // finite: every tile holds at least one key < Tk
      alpha[i] = expf(mrow[i] - mnew);                             // exp(-inf) = 0 on the first tile
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = expf(s[i][j] - mnew);
        Ps[(4 * ty + i) * PS + tx + 16 * j] = p;
        sum += p;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      lrow[i] = lrow[i] * alpha[i] + sum;
      mrow[i] = mnew;
#pragma unroll
      for (int c = 0; c < NC; ++c) o[i][c] *= alpha[i];
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < TK; ++k) {
      float p[4], v[NC];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = Ps[(4 * ty + i) * PS + k];
#pragma unroll
      for (int c = 0; c < NC; ++c) v[c] = Vs[k * D + NC * tx + c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c) o[i][c] = fmaf(p[i], v[c], o[i][c]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + 4 * ty + i;
    if (q >= Tq) continue;
    const float inv = 1.0f / lrow[i];
    TOp* dst = O + (static_cast<size_t>(b) * Tq + q) * ldo + h * D + NC * tx;
#pragma unroll
    for (int c = 0; c < NC; ++c) store_op(dst + c, o[i][c] * inv);
  }
  if (probs == nullptr) return;
  // second pass (the `attn` entry of CampNet's output dict: head-averaged probabilities of decoder layer 0, transformer.py:
  // 410-416, :803): with the final (max, sum) of every row known, recompute the score tiles and accumulate
  // p / heads into probs[B, Tq, Tk] (zeroed by the caller; one atomicAdd per head and element, order-independent for 2 heads)
  for (int k0 = 0; k0 < Tk; k0 += TK) {
    __syncthreads();
    for (int i = tid; i < TK * D; i += 256) {
      const int r = i / D, c = i % D, k = k0 + r;
      Ks[r * QS + c] = k < Tk ? to_f32(K[(static_cast<size_t>(b) * Tk + k) * ldkv + h * D + c + koff]) : 0.f;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int c = 0; c < D; ++c) {
      float a[4], kk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Qs[(4 * ty + i) * QS + c];
#pragma unroll
      for (int j = 0; j < 4; ++j) kk[j] = Ks[(tx + 16 * j) * QS + c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(a[i], kk[j], s[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx + 16 * j;
      if (k >= Tk) continue;
      const bool padded = key_keep && key_keep[static_cast<size_t>(b) * Tk + k] == 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int q = q0 + 4 * ty + i;
        if (q >= Tq) continue;
        const float pr = expf((padded ? -1e8f : s[i][j]) - mrow[i]) / lrow[i];
        atomicAdd(probs + (static_cast<size_t>(b) * Tq + q) * Tk + k, pr * probs_scale);
      }
    }
  }
}

struct AttnW { ConvW qkv, q, kv, out; };      // self-attention uses qkv + out; encoder-decoder attention uses q + kv + out
struct EncLayerW { LNW ln1, ln2; AttnW self; ConvW ffn1, ffn2; };
struct DecLayerW { LNW ln1, ln2, ln3; AttnW self, cross; ConvW ffn1, ffn2; };

}  // namespace
}  // namespace fse

using namespace fse;

struct fse_campnet {
  fse_campnet_config cfg{};
  LayerCtx ctx;
  bool loaded = false;
  float *embed_tokens = nullptr, *sinus = nullptr, *mask_emb = nullptr;
  int max_pos = 0;
  float alpha = 1.f;
  std::vector<EncLayerW> enc;
  LNW enc_norm;
  std::vector<DecLayerW> dec;
  LNW dec_norm;
  ConvBlocksW fine;
  ConvW out_coarse, out_fine;
  fse_mel_encoder* mel = nullptr;
  bool attn_tc = false;                 // tcgen05 attention (FSE_MODE_TC_BF16; FSE_CAMP_ATTN=simt selects the CUDA-core kernel)
  bool attn_tc2 = false;                // ... its two-query-tile schedule (the default; FSE_CAMP_ATTN=tc selects one tile per CTA)
  bool probs_simt = false;              // FSE_CAMP_PROBS=simt: the probabilities call of layer 0 on the CUDA-core kernel (cross-check)
  struct VtMap { const void* buf = nullptr; int Tkp = 0, B = 0; CUtensorMap map{}; } vtmap;
};

namespace {

struct KWs {
  RowBufs r;            // x32 / tmp32 / y32 / opA / opB (4H wide) / m0 / m1 over R = B * max(T, Tt) rows
  void* qkv;            // [R, 3H] operand
  void* kvx;            // [B*Tt, 2H] operand
  void* vt;             // [B*H, Tp] operand: V^T of the current attention (tensor-core attention only), Tp = max(T, Tt) rounded up to 8
  size_t vt_bytes;
  float* enc32;         // [B*Tt, H]
  void* encop;          // [B*Tt, H] operand
  float* enc_keep;      // [B*Tt]
  float* xin;           // [B*T, M]
  float* melc;          // [B*T, M]
  float *nonpad, *keep, *first;   // [B*T]
  int* pos;             // [R]
  void* melws;          // MelEncoder workspace
  size_t melws_bytes, bytes;
};
KWs kcarve(const fse_campnet* h, void* base, int B, int Tt, int T) {
  const size_t H = h->cfg.hidden, M = h->cfg.n_mels, es = h->ctx.bf16 ? 2 : 4;
  const size_t Rt = static_cast<size_t>(B) * Tt, Rf = static_cast<size_t>(B) * T, R = Rt > Rf ? Rt : Rf;
  KWs w{};
  w.r = carve_rows(h->ctx, base, R, 4);
  size_t off = w.r.bytes;
  auto take = [&](size_t b) { size_t o = off; off += align_up(b, 1024); return o; };
  uint8_t* p = static_cast<uint8_t*>(base);
  w.qkv = p + take(R * 3 * H * es);
  w.kvx = p + take(Rt * 2 * H * es);
  const size_t Tp = static_cast<size_t>(((T > Tt ? T : Tt) + 7) / 8 * 8);
  w.vt_bytes = h->attn_tc ? static_cast<size_t>(B) * H * Tp * 2 : 0;
  w.vt = p + take(w.vt_bytes);
  w.enc32 = reinterpret_cast<float*>(p + take(Rt * H * 4));
  w.encop = p + take(Rt * H * es);
  w.enc_keep = reinterpret_cast<float*>(p + take(Rt * 4));
  w.xin = reinterpret_cast<float*>(p + take(Rf * M * 4));
  w.melc = reinterpret_cast<float*>(p + take(Rf * M * 4));
  w.nonpad = reinterpret_cast<float*>(p + take(Rf * 4));
  w.keep = reinterpret_cast<float*>(p + take(Rf * 4));
  w.first = reinterpret_cast<float*>(p + take(Rf * 4));
  w.pos = reinterpret_cast<int*>(p + take(R * 4));
  w.melws_bytes = h->mel ? static_cast<size_t>(fse_mel_encoder_workspace_bytes(h->mel, B, T)) : 0;
  w.melws = p + take(w.melws_bytes);
  w.bytes = off;
  return w;
}

// rows [r0, r0 + n) of an in_proj_weight [3C, C] as a bias-free 1-tap GEMM weight
int pack_rows(fse_campnet* h, const TensorTable& tt, const std::string& name, int r0, int n, ConvW& cw) {
  const int C = h->cfg.hidden;
  int rc = FSE_OK;
  const float* w = tt.get(name, static_cast<int64_t>(3) * C * C, &rc);
  if (rc) return rc;
  const int zero = 0;
  return pack_conv_raw(&h->ctx, name, w + static_cast<size_t>(r0) * C, nullptr, n, C, 1, &zero, cw);
}
int pack_linear(fse_campnet* h, const TensorTable& tt, const std::string& name, int Cout, int Cin, bool bias, ConvW& cw) {
  int rc = FSE_OK;
  const float* w = tt.get(name + ".weight", static_cast<int64_t>(Cout) * Cin, &rc);
  if (rc) return rc;
  const float* b = nullptr;
  if (bias) { b = tt.get(name + ".bias", Cout, &rc); if (rc) return rc; }
  const int zero = 0;
  return pack_conv_raw(&h->ctx, name, w, b, Cout, Cin, 1, &zero, cw);
}
int load_attn(fse_campnet* h, const TensorTable& tt, const std::string& pre, bool cross, AttnW& a) {
  const int C = h->cfg.hidden;
  if (cross) {
    FSE_TRY(pack_rows(h, tt, pre + ".in_proj_weight", 0, C, a.q));
    FSE_TRY(pack_rows(h, tt, pre + ".in_proj_weight", C, 2 * C, a.kv));
  } else {
    FSE_TRY(pack_rows(h, tt, pre + ".in_proj_weight", 0, 3 * C, a.qkv));
  }
  return pack_linear(h, tt, pre + ".out_proj", C, C, false, a.out);
}

template <typename TOp>
int attention(fse_campnet* h, const void* Q, int ldq, int qoff, const void* K, const void* V, int ldkv, int koff, int voff,
              const float* key_keep, void* O, int B, int Tq, int Tk, float* probs, const void* vt, int Tkp, cudaStream_t st) {
  if constexpr (std::is_same<TOp, __nv_bfloat16>::value) {
    // probabilities (decoder layer 0) come out of the two-tile tensor-core kernel when the keys fit one tile (Tk <= 128: the
    // cross-attention over the phone sequence); otherwise that call runs the CUDA-core kernel below
    const bool probs_tc = probs != nullptr && h->attn_tc2 && Tk <= kAtcN && !h->probs_simt;
    if (h->attn_tc && vt != nullptr && (probs == nullptr || probs_tc)) {
      // tensor cores: S = Q K^T and O = P V as tcgen05.mma, V^T from the projection's epilogue (attention_tc.cuh)
      // function attributes are per context: flags keyed by device ordinal, not process-global
      int dev = 0;
      FSE_CUDA(cudaGetDevice(&dev));
      if (dev < 0 || dev >= kMaxDevices) return fail(FSE_ECUDA, "device ordinal %d out of range", dev);
      static bool tc_attr[kMaxDevices] = {};
      if (!tc_attr[dev]) {
        FSE_CUDA(cudaFuncSetAttribute(camp_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtcSmemBytes));
        tc_attr[dev] = true;
      }
      const CUtensorMap *mq = nullptr, *mk = nullptr;
      FSE_TRY(get_act_map(&h->ctx, Q, ldq, Tq, B, 64, &mq));
      const CUtensorMap q_map = *mq;                     // by value: the second lookup may recycle the cache
      FSE_TRY(get_act_map(&h->ctx, K, ldkv, Tk, B, 64, &mk));
      if (!(h->vtmap.buf == vt && h->vtmap.Tkp == Tkp && h->vtmap.B == B)) {
        FSE_TRY(make_map_w(&h->vtmap.map, vt, Tkp, B * h->cfg.hidden, 64, kAttD));
        h->vtmap.buf = vt; h->vtmap.Tkp = Tkp; h->vtmap.B = B;
      }
      AttnTcParams ap{Tq, Tk, h->cfg.heads, qoff, koff, h->cfg.hidden, key_keep, static_cast<__nv_bfloat16*>(O)};
      if (probs_tc) {
        FSE_CUDA(cudaMemsetAsync(probs, 0, static_cast<size_t>(B) * Tq * Tk * sizeof(float), st));
        ap.probs = probs;
      }
      if (h->attn_tc2) {       // two query tiles per CTA, two softmax warp groups (second schedule of attention_tc.cuh)
        static bool tc2_attr[kMaxDevices] = {};
        if (!tc2_attr[dev]) {
          FSE_CUDA(cudaFuncSetAttribute(camp_attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtc2SmemBytes));
          tc2_attr[dev] = true;
        }
        dim3 grid2((Tq + 2 * kAtcM - 1) / (2 * kAtcM), h->cfg.heads, B);
        camp_attention_tc2_kernel<<<grid2, kAtc2Threads, kAtc2SmemBytes, st>>>(q_map, *mk, h->vtmap.map, ap);
        FSE_CUDA(cudaGetLastError());
        ++h->ctx.launches;
        return FSE_OK;
      }
      dim3 grid((Tq + kAtcM - 1) / kAtcM, h->cfg.heads, B);
      camp_attention_tc_kernel<<<grid, kAtcThreads, kAtcSmemBytes, st>>>(q_map, *mk, h->vtmap.map, ap);
      FSE_CUDA(cudaGetLastError());
      ++h->ctx.launches;
      return FSE_OK;
    }
  }
  constexpr size_t smem = (2 * kAttTile * (kAttD + 1) + kAttTile * kAttD + kAttTile * (kAttTile + 1)) * sizeof(float);
  int dev_simt = 0;
  FSE_CUDA(cudaGetDevice(&dev_simt));
  if (dev_simt < 0 || dev_simt >= kMaxDevices) return fail(FSE_ECUDA, "device ordinal %d out of range", dev_simt);
  static bool attr_set[kMaxDevices] = {};
  auto kern = camp_attention_kernel<TOp>;
  if (!attr_set[dev_simt]) {
    FSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set[dev_simt] = true;
  }
  if (probs) FSE_CUDA(cudaMemsetAsync(probs, 0, static_cast<size_t>(B) * Tq * Tk * sizeof(float), st));
  dim3 grid((Tq + kAttTile - 1) / kAttTile, h->cfg.heads, B);
  kern<<<grid, 256, smem, st>>>(static_cast<const TOp*>(Q), ldq, qoff, static_cast<const TOp*>(K), static_cast<const TOp*>(V), ldkv, koff,
                                voff, key_keep, static_cast<TOp*>(O), h->cfg.hidden, Tq, Tk, probs, 1.0f / static_cast<float>(h->cfg.heads));
  FSE_CUDA(cudaGetLastError());
  ++h->ctx.launches;
  return FSE_OK;
}

template <typename TOp>
int forward_impl(fse_campnet* h, const int64_t* txt, const float* mels, const float* mask, float* out_coarse, float* out_fine, float* attn_out,
                 float* enc_out, int B, int Tt, int T, void* ws, cudaStream_t st) {
  const KWs w = kcarve(h, ws, B, Tt, T);
  LayerCtx* ctx = &h->ctx;
  const int H = h->cfg.hidden, M = h->cfg.n_mels, k = h->cfg.ffn_kernel;
  const size_t Rt = static_cast<size_t>(B) * Tt, Rf = static_cast<size_t>(B) * T;
  const float qscale = 1.0f / sqrtf(static_cast<float>(kAttD));
  const float fscale = static_cast<float>(std::pow(static_cast<double>(k), -0.5));
  auto launched = [&]() -> int { FSE_CUDA(cudaGetLastError()); ++ctx->launches; return FSE_OK; };
  // V^T rows are padded to a multiple of 8 keys (16-byte TMA strides); the pad columns must be finite zeros (P = 0 there)
  TOp* vt = h->attn_tc ? static_cast<TOp*>(w.vt) : nullptr;
  const int TpT = (Tt + 7) / 8 * 8, TpF = (T + 7) / 8 * 8;
  if (vt) FSE_CUDA(cudaMemsetAsync(w.vt, 0, w.vt_bytes, st));

  // ---- text encoder (TransformerEncoder.forward): keep = m0 = (txt != 0)
  camp_positions_kernel<<<(B + 63) / 64, 64, 0, st>>>(txt, nullptr, w.pos, B, Tt);
  FSE_TRY(launched());
  camp_embed_kernel<<<row_blocks(Rt), 256, 0, st>>>(txt, w.pos, h->embed_tokens, h->sinus, w.r.x32, w.r.m0, static_cast<int>(Rt), H,
                                                   h->cfg.vocab, h->max_pos, sqrtf(static_cast<float>(H)));
  FSE_TRY(launched());
  for (const EncLayerW& L : h->enc) {
    FSE_TRY((layer_norm<TOp>(ctx, w.r.x32, L.ln1, nullptr, nullptr, nullptr, w.r.opA, nullptr, Rt, st)));
    EpiScaleCols<TOp> eq{static_cast<TOp*>(w.qkv), 3 * H, Tt, H, qscale, vt, 2 * H, TpT, h->cfg.heads};
    FSE_TRY((run_conv<TOp>(ctx, L.self.qkv, w.r.opA, B, Tt, eq, st)));
    FSE_TRY((attention<TOp>(h, w.qkv, 3 * H, 0, w.qkv, w.qkv, 3 * H, H, 2 * H, w.r.m0, w.r.opA, B, Tt, Tt, nullptr, vt, TpT, st)));
    EpiResidualMask eo{L.self.out.bias, w.r.x32, w.r.m0, H, Tt};
    FSE_TRY((run_conv<TOp>(ctx, L.self.out, w.r.opA, B, Tt, eo, st)));
    FSE_TRY((layer_norm<TOp>(ctx, w.r.x32, L.ln2, nullptr, nullptr, nullptr, w.r.opA, nullptr, Rt, st)));
    EpiGeluScale<TOp> e1{L.ffn1.bias, static_cast<TOp*>(w.r.opB), 4 * H, Tt, fscale};
    FSE_TRY((run_conv<TOp>(ctx, L.ffn1, w.r.opA, B, Tt, e1, st)));
    EpiResidualMask e2{L.ffn2.bias, w.r.x32, w.r.m0, H, Tt};
    FSE_TRY((run_conv<TOp>(ctx, L.ffn2, w.r.opB, B, Tt, e2, st)));
  }
  FSE_TRY((layer_norm<TOp>(ctx, w.r.x32, h->enc_norm, nullptr, w.r.m0, nullptr, w.encop, w.enc32, Rt, st)));
  row_absmask_kernel<<<row_blocks(Rt), 256, 0, st>>>(w.enc32, w.enc_keep, nullptr, static_cast<int>(Rt), H);   // encoder_padding_mask (:783)
  FSE_TRY(launched());
  if (enc_out) FSE_CUDA(cudaMemcpyAsync(enc_out, w.enc32, Rt * H * sizeof(float), cudaMemcpyDeviceToDevice, st));

  // ---- coarse decoder input: MelEncoder(mels (1-m) + mask_emb m) * nonpad, + alpha * positions
  camp_mel_input_kernel<<<row_blocks(Rf), 256, 0, st>>>(mels, mask, h->mask_emb, w.xin, w.nonpad, static_cast<int>(Rf), M);
  FSE_TRY(launched());
  FSE_TRY(fse_mel_encoder_forward(h->mel, w.xin, nullptr, w.nonpad, w.r.x32, B, T, w.melws, static_cast<int64_t>(w.melws_bytes), st));
  ctx->launches += fse_mel_encoder_last_launches(h->mel);
  row_absmask_kernel<<<row_blocks(Rf), 256, 0, st>>>(w.r.x32, w.keep, w.first, static_cast<int>(Rf), H);
  FSE_TRY(launched());
  camp_positions_kernel<<<(B + 63) / 64, 64, 0, st>>>(nullptr, w.first, w.pos, B, T);
  FSE_TRY(launched());
  camp_add_pos_kernel<<<row_blocks(Rf), 256, 0, st>>>(w.r.x32, w.pos, h->sinus, w.keep, static_cast<int>(Rf), H, h->max_pos, h->alpha);
  FSE_TRY(launched());

  // ---- decoder_coarse (TransformerDecoder.forward): 6 x [self-attention, encoder-decoder attention, causal-padded conv FFN]
  bool first_layer = true;
  for (const DecLayerW& L : h->dec) {
    FSE_TRY((layer_norm<TOp>(ctx, w.r.x32, L.ln1, nullptr, nullptr, nullptr, w.r.opA, nullptr, Rf, st)));
    EpiScaleCols<TOp> eq{static_cast<TOp*>(w.qkv), 3 * H, T, H, qscale, vt, 2 * H, TpF, h->cfg.heads};
    FSE_TRY((run_conv<TOp>(ctx, L.self.qkv, w.r.opA, B, T, eq, st)));
    FSE_TRY((attention<TOp>(h, w.qkv, 3 * H, 0, w.qkv, w.qkv, 3 * H, H, 2 * H, nullptr, w.r.opA, B, T, T, nullptr, vt, TpF, st)));
    EpiResidualMask eo{L.self.out.bias, w.r.x32, nullptr, H, T};
    FSE_TRY((run_conv<TOp>(ctx, L.self.out, w.r.opA, B, T, eo, st)));

    FSE_TRY((layer_norm<TOp>(ctx, w.r.x32, L.ln2, nullptr, nullptr, nullptr, w.r.opA, nullptr, Rf, st)));
    EpiScaleCols<TOp> ecq{static_cast<TOp*>(w.qkv), H, T, H, qscale, nullptr, 0, 0, 0};          // q of the cross attention: [B*T, H]
    FSE_TRY((run_conv<TOp>(ctx, L.cross.q, w.r.opA, B, T, ecq, st)));
    EpiScaleCols<TOp> ekv{static_cast<TOp*>(w.kvx), 2 * H, Tt, 0, 1.f, vt, H, TpT, h->cfg.heads};
    FSE_TRY((run_conv<TOp>(ctx, L.cross.kv, w.encop, B, Tt, ekv, st)));
    // layer 0 also reports its head-averaged probabilities (the CUDA-core kernel's second pass)
    float* probs = first_layer ? attn_out : nullptr;
    first_layer = false;
    FSE_TRY((attention<TOp>(h, w.qkv, H, 0, w.kvx, w.kvx, 2 * H, 0, H, w.enc_keep, w.r.opA, B, T, Tt, probs, vt, TpT, st)));
    EpiResidualMask ex{L.cross.out.bias, w.r.x32, nullptr, H, T};
    FSE_TRY((run_conv<TOp>(ctx, L.cross.out, w.r.opA, B, T, ex, st)));

    FSE_TRY((layer_norm<TOp>(ctx, w.r.x32, L.ln3, nullptr, nullptr, nullptr, w.r.opA, nullptr, Rf, st)));
    EpiGeluScale<TOp> e1{L.ffn1.bias, static_cast<TOp*>(w.r.opB), 4 * H, T, fscale};
    FSE_TRY((run_conv<TOp>(ctx, L.ffn1, w.r.opA, B, T, e1, st)));
    EpiResidualMask e2{L.ffn2.bias, w.r.x32, w.keep, H, T};
    FSE_TRY((run_conv<TOp>(ctx, L.ffn2, w.r.opB, B, T, e2, st)));
  }
  // layer_norm(x) * keep; the following `* mel_nonpadding` (campnet.py:60) is implied: keep <= nonpad row by row
  FSE_TRY((layer_norm<TOp>(ctx, w.r.x32, h->dec_norm, nullptr, w.keep, nullptr, w.r.opA, nullptr, Rf, st)));
  EpiMelOut ec{w.nonpad, mask, mels, out_coarse, w.melc, M, T, 1};
  FSE_TRY((run_conv<TOp>(ctx, h->out_coarse, w.r.opA, B, T, ec, st)));

  // ---- fine decoder: MelEncoder(mel_coarse) * nonpad -> ConvBlocks -> 192 -> 80, mel_out_fine = mel_coarse + y m
  FSE_TRY(fse_mel_encoder_forward(h->mel, w.melc, nullptr, w.nonpad, w.r.x32, B, T, w.melws, static_cast<int64_t>(w.melws_bytes), st));
  ctx->launches += fse_mel_encoder_last_launches(h->mel);
  row_absmask_kernel<<<row_blocks(Rf), 256, 0, st>>>(w.r.x32, w.r.m0, nullptr, static_cast<int>(Rf), H);
  FSE_TRY(launched());
  // post_net1 reads opA (with its k = 3 halo) while its epilogue stores: the result goes to the (now free) qkv buffer as [B*T, H]
  EpiBiasMaskOp<TOp> ep{h->fine.post.bias, w.r.m0, static_cast<TOp*>(w.qkv), H, T};
  FSE_TRY((conv_blocks_forward<TOp>(ctx, h->fine, w.r, B, T, ep, st)));
  EpiMelOut ef{w.nonpad, mask, w.melc, nullptr, out_fine, M, T, 0};
  return run_conv<TOp>(ctx, h->out_fine, w.qkv, B, T, ef, st);
}

}  // namespace

extern "C" {

int fse_campnet_create(const fse_campnet_config* cfg, fse_campnet** out) {
  if (!cfg || !out) return fail(FSE_EINVAL, "null argument");
  if (cfg->mode < 0 || cfg->mode > 3) return fail(FSE_EINVAL, "unknown mode %d", cfg->mode);
  if (cfg->heads < 1 || cfg->hidden != cfg->heads * kAttD)
    return fail(FSE_EINVAL, "hidden must equal heads * %d (head_dim of the attention kernel); got hidden %d, heads %d", kAttD, cfg->hidden, cfg->heads);
  if (cfg->hidden % 64 != 0 || cfg->hidden > 512) return fail(FSE_EINVAL, "hidden must be a multiple of 64, <= 512");
  if (cfg->vocab <= 0) return fail(FSE_EINVAL, "vocab must be positive");
  if (cfg->n_mels <= 0 || cfg->n_mels % 16 != 0) return fail(FSE_EINVAL, "n_mels must be a positive multiple of 16");
  if (cfg->enc_layers < 0 || cfg->dec_layers < 1 || cfg->fine_blocks < 1) return fail(FSE_EINVAL, "layer counts out of range");
  if (cfg->ffn_kernel < 1 || cfg->ffn_kernel > kMaxTaps || cfg->ffn_kernel % 2 == 0) return fail(FSE_EINVAL, "ffn_kernel must be odd, <= %d", kMaxTaps);
  if (cfg->fine_kernel < 1 || cfg->fine_kernel > kMaxTaps || cfg->fine_kernel % 2 == 0) return fail(FSE_EINVAL, "fine_kernel must be odd, <= %d", kMaxTaps);
  int dev = 0;
  FSE_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  FSE_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(FSE_ECUDA, "device is sm_%d%d; this library is built for sm_100a only (no fallback)", prop.major, prop.minor);
  auto* h = new fse_campnet();
  h->cfg = *cfg;
  h->ctx.mode = cfg->mode;
  h->ctx.bf16 = mode_is_bf16(cfg->mode);       // FSE_MODE_TC_TF32: tf32 tensor-core GEMMs over fp32 rows, fp32 CUDA-core attention
  h->ctx.hidden = cfg->hidden;
  // FSE_CAMP_ATTN = "tc2" (default in FSE_MODE_TC_BF16: tcgen05, two query tiles per CTA) | "tc" (one tile per CTA) | "simt"
  const char* sel = std::getenv("FSE_CAMP_ATTN");
  h->attn_tc = cfg->mode == FSE_MODE_TC_BF16 && !(sel && std::string(sel) == "simt");
  h->attn_tc2 = h->attn_tc && !(sel && std::string(sel) == "tc");
  const char* psel = std::getenv("FSE_CAMP_PROBS");
  h->probs_simt = psel && std::string(psel) == "simt";
  *out = h;
  return FSE_OK;
}

void fse_campnet_destroy(fse_campnet* h) {
  if (!h) return;
  if (h->mel) fse_mel_encoder_destroy(h->mel);
  h->ctx.release();
  delete h;
}

int fse_campnet_load_weights(fse_campnet* h, const fse_tensor* tensors, int32_t n) {
  if (!h || !tensors || n <= 0) return fail(FSE_EINVAL, "null argument");
  if (h->loaded) return fail(FSE_ESTATE, "weights already loaded");
  TensorTable tt(tensors, n);
  const auto& c = h->cfg;
  const int H = c.hidden, k = c.ffn_kernel;
  LayerCtx* ctx = &h->ctx;
  FSE_TRY(load_vec(ctx, tt, "mask_emb", c.n_mels, &h->mask_emb));
  FSE_TRY(load_vec(ctx, tt, "encoder.embed_tokens.weight", static_cast<int64_t>(c.vocab) * H, &h->embed_tokens));
  {   // SinusoidalPositionalEmbedding.get_embedding (transformer.py:33-49) in fp32 like the reference; row 0 = padding = zeros
    h->max_pos = 4096;
    const int half = H / 2;
    const float step = static_cast<float>(std::log(10000.0) / (half - 1));
    std::vector<float> tab(static_cast<size_t>(h->max_pos) * H, 0.f);
    for (int p = 1; p < h->max_pos; ++p)
      for (int j = 0; j < half; ++j) {
        const float f = std::exp(static_cast<float>(j) * -step);
        const float a = static_cast<float>(p) * f;
        tab[static_cast<size_t>(p) * H + j] = std::sin(a);
        tab[static_cast<size_t>(p) * H + half + j] = std::cos(a);
      }
    FSE_TRY(dev_f32(ctx, tab.data(), tab.size(), &h->sinus));
  }
  h->enc.resize(c.enc_layers);
  for (int i = 0; i < c.enc_layers; ++i) {
    const std::string pre = "encoder.layers." + std::to_string(i) + ".op.";
    EncLayerW& L = h->enc[i];
    FSE_TRY(load_ln(ctx, tt, pre + "layer_norm1", H, L.ln1));
    FSE_TRY(load_attn(h, tt, pre + "self_attn", false, L.self));
    FSE_TRY(load_ln(ctx, tt, pre + "layer_norm2", H, L.ln2));
    FSE_TRY(pack_conv(ctx, tt, pre + "ffn.ffn_1", 4 * H, H, k, 1, L.ffn1));                    // padding 'SAME'
    FSE_TRY(pack_linear(h, tt, pre + "ffn.ffn_2", H, 4 * H, true, L.ffn2));
  }
  FSE_TRY(load_ln(ctx, tt, "encoder.layer_norm", H, h->enc_norm));
  {
    int rc = FSE_OK;
    const float* a = tt.get("decoder_coarse.pos_embed_alpha", 1, &rc);
    if (rc) return rc;
    h->alpha = a[0];
  }
  h->dec.resize(c.dec_layers);
  for (int i = 0; i < c.dec_layers; ++i) {
    const std::string pre = "decoder_coarse.layers." + std::to_string(i) + ".op.";
    DecLayerW& L = h->dec[i];
    FSE_TRY(load_ln(ctx, tt, pre + "layer_norm1", H, L.ln1));
    FSE_TRY(load_attn(h, tt, pre + "self_attn", false, L.self));
    FSE_TRY(load_ln(ctx, tt, pre + "layer_norm2", H, L.ln2));
    FSE_TRY(load_attn(h, tt, pre + "encoder_attn", true, L.cross));
    FSE_TRY(load_ln(ctx, tt, pre + "layer_norm3", H, L.ln3));
    FSE_TRY(pack_conv(ctx, tt, pre + "ffn.ffn_1.1", 4 * H, H, k, 1, L.ffn1, /*left=*/true));   // padding 'LEFT' (:84-88)
    FSE_TRY(pack_linear(h, tt, pre + "ffn.ffn_2", H, 4 * H, true, L.ffn2));
  }
  FSE_TRY(load_ln(ctx, tt, "decoder_coarse.layer_norm", H, h->dec_norm));
  FSE_TRY(load_conv_blocks(ctx, tt, "decoder_fine.", c.fine_blocks, nullptr, 2, c.fine_kernel, 3, h->fine));
  FSE_TRY(pack_linear(h, tt, "mel_out_coarse", c.n_mels, H, false, h->out_coarse));
  FSE_TRY(pack_linear(h, tt, "mel_out_fine", c.n_mels, H, false, h->out_fine));
  {   // the MelEncoder is the library's own handle, fed the `mel_encoder.*` entries without the prefix
    fse_mel_encoder_config mc{c.n_mels, H, c.mode};
    FSE_TRY(fse_mel_encoder_create(&mc, &h->mel));
    static const char* names[] = {"encoder.0.weight", "encoder.0.bias", "encoder.2.weight", "encoder.2.bias", "fc_out.weight", "fc_out.bias"};
    std::vector<fse_tensor> sub;
    for (const char* nm : names) {
      auto it = tt.map.find(std::string("mel_encoder.") + nm);
      if (it == tt.map.end()) return fail(FSE_EINVAL, "missing weight tensor 'mel_encoder.%s'", nm);
      sub.push_back(fse_tensor{nm, it->second->data, it->second->numel});
    }
    FSE_TRY(fse_mel_encoder_load_weights(h->mel, sub.data(), static_cast<int32_t>(sub.size())));
  }
  h->loaded = true;
  return FSE_OK;
}

int64_t fse_campnet_workspace_bytes(const fse_campnet* h, int32_t B, int32_t Tt, int32_t T) {
  if (!h || !h->loaded || B <= 0 || Tt <= 0 || T <= 0) return 0;
  return static_cast<int64_t>(kcarve(h, nullptr, B, Tt, T).bytes);
}

int64_t fse_campnet_last_launches(const fse_campnet* h) { return h ? h->ctx.launches : 0; }

int fse_campnet_forward(fse_campnet* h, const int64_t* txt, const float* mels, const float* time_mel_masks, float* mel_out_coarse,
                        float* mel_out_fine, float* attn, float* encoder_out, int32_t B, int32_t Tt, int32_t T, void* workspace,
                        int64_t workspace_bytes, void* stream) {
  if (!h) return fail(FSE_EINVAL, "null handle");
  if (!h->loaded) return fail(FSE_ESTATE, "weights not loaded");
  if (!txt || !mels || !time_mel_masks || !mel_out_coarse || !mel_out_fine) return fail(FSE_EINVAL, "null argument");
  if (B <= 0 || Tt <= 0 || T <= 0) return fail(FSE_EINVAL, "B, Tt and T must be positive");
  if (Tt + 1 >= h->max_pos || T + 1 >= h->max_pos) return fail(FSE_EINVAL, "sequence longer than the positional table (%d)", h->max_pos);
  if ((static_cast<size_t>(B) * T * h->cfg.n_mels) % 4 != 0) return fail(FSE_EINVAL, "B*T*n_mels must be a multiple of 4");
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 1023)) return fail(FSE_EINVAL, "workspace must be non-null and 1024-byte aligned");
  if (workspace_bytes < fse_campnet_workspace_bytes(h, B, Tt, T)) return fail(FSE_EINVAL, "workspace too small");
  h->ctx.launches = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return h->ctx.bf16 ? forward_impl<__nv_bfloat16>(h, txt, mels, time_mel_masks, mel_out_coarse, mel_out_fine, attn, encoder_out, B, Tt, T, workspace, st)
                     : forward_impl<float>(h, txt, mels, time_mel_masks, mel_out_coarse, mel_out_fine, attn, encoder_out, B, Tt, T, workspace, st);
}

}  // extern "C"
