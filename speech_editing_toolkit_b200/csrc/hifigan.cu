// HiFi-GAN generator forward on sm_100a (modules/vocoder/hifigan/hifigan.py:27-64, 101-142 of the reference).
// Everything is channels-last [B, T_stage, C]; every conv / transposed conv is a conv_gemm launch:
//   Conv1d(k, dilation d, "same" padding)   -> k taps at offsets (j - (k-1)/2) d
//   ConvTranspose1d(k, stride u, pad p)      -> ceil(k/u)-tap GEMM with N' = u*C_out columns (one column block
//                                               per output phase r), rows q in [0, T_in], element (q, r, co)
//                                               lands at sample q*u + r - p
// weight-norm (w = g v / ||v||) is folded once at load time.
#include <cmath>
#include <cstdlib>
#include <deque>

#include "epilogues.cuh"
#include "fse_common.cuh"
#include "resblock_fused.cuh"

namespace fse {

__device__ __forceinline__ float lrelu(float x, float slope) { return x >= 0.f ? x : x * slope; }

// out = lrelu(acc + bias)  — conv_pre (followed by the stage-0 leaky_relu, hifigan.py:127-129) and ResBlock1 convs1 (:53-55)
template <typename TOp>
struct EpiAct {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  const float* bias;
  TOp* out;   // [B*T, N]
  int N, T;
  float slope;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = lrelu(acc[i] + __ldg(bias + n0 + i), slope);
    st_vec<NV>(out + (static_cast<size_t>(b) * T + t) * N + n0, v);
  }
};

// transposed-conv phase scatter: x = acc + bias (fp32 stage input) and xa = lrelu(x, 0.1) (hifigan.py:130, :53)
template <typename TOp>
struct EpiUp {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  const float* bias;   // [Cout]
  float* x;            // [B, Tout, Cout]
  TOp* xa;             // [B, Tout, Cout]
  int Cout, Tout, u, pad;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int q, int n0, const float* acc, const float*) const {
    const int r = n0 / Cout, co = n0 % Cout;
    const int tau = q * u + r - pad;
    if (tau < 0 || tau >= Tout) return;
    const size_t o = (static_cast<size_t>(b) * Tout + tau) * Cout + co;
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = acc[i] + __ldg(bias + co + i);
    st_vec<NV>(x + o, v);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = lrelu(v[i], 0.1f);
    st_vec<NV>(xa + o, v);
  }
};

// ResBlock1 convs2 + residual (hifigan.py:56-57); the last pair of each block also folds the
// mean over the parallel blocks (hifigan.py:131-137) and the activation feeding the next stage.
template <typename TOp, bool kSum>     // kSum: kinds 2 and 3 (needs the running sum); kinds 0 and 1 use the leaner kSum = false
struct EpiResAdd {
  static constexpr int kAux = 1;       // per column: the residual input (prefetched one chunk ahead)
  static constexpr int kLate = kSum ? 1 : 0;   // ... and the running sum over the parallel resblocks (read-modify-write)
  static constexpr bool kTransposed = true;
  const float* bias;
  const float* res;    // [B*T, N] fp32 residual input
  float* y;            // [B*T, N] fp32 (kind 0)
  TOp* ya;             // [B*T, N] lrelu(y, 0.1) (kind 0)
  float* xs;           // [B*T, N] running sum over resblocks (kind 1..3)
  TOp* next_a;         // [B*T, N] lrelu(mean, slope_next) operand for the next stage (kind 3)
  float* final_f32;    // [B*T, N] lrelu(mean, slope_next) fp32 for conv_post (kind 3, last stage) or null
  int N, T;
  int kind;            // 0: inside a block; 1: first block end (xs = v); 2: middle (xs += v); 3: last (mean)
  float num_kernels, slope_next;
  template <int NV>
  __device__ __forceinline__ void load_aux(int b, int t, int n0, float* aux) const {
    const size_t o = (static_cast<size_t>(b) * T + t) * N + n0;
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      const float4 v = reinterpret_cast<const float4*>(res + o)[i];
      aux[4 * i] = v.x; aux[4 * i + 1] = v.y; aux[4 * i + 2] = v.z; aux[4 * i + 3] = v.w;
    }
    // the running sum is read right before it is updated (load_late), a chunk after this call: pull its lines into L2 now so that
    // read is not a DRAM round trip per chunk (measured: the end-of-block pairs took 1.5-1.9x the time of the others)
    if constexpr (kSum) asm volatile("prefetch.global.L2 [%0];" ::"l"(xs + o));
  }
  template <int NV>
  __device__ __forceinline__ void load_late(int b, int t, int n0, float* dst) const {
    const size_t o = (static_cast<size_t>(b) * T + t) * N + n0;
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      const float4 v = reinterpret_cast<const float4*>(xs + o)[i];
      dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
    }
  }
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float* aux) const {
    const size_t o = (static_cast<size_t>(b) * T + t) * N + n0;
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = (acc[i] + __ldg(bias + n0 + i)) + aux[i];
    if (kind == 0) {
      st_vec<NV>(y + o, v);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = lrelu(v[i], 0.1f);
      st_vec<NV>(ya + o, v);
    } else if (kind == 1) {
      st_vec<NV>(xs + o, v);
    } else {
      if constexpr (kSum) {               // (kSum = false reaches here only with a single resblock per stage: the mean is v itself)
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = aux[NV + i] + v[i];
      }
      if (kind == 2) {
        st_vec<NV>(xs + o, v);
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = lrelu(__fdiv_rn(v[i], num_kernels), slope_next);
        if (final_f32) st_vec<NV>(final_f32 + o, v);
        else st_vec<NV>(next_a + o, v);
      }
    }
  }
};

// conv_post (C -> 1, k taps) + tanh on the already-activated fp32 stage output (hifigan.py:138-140)
// One block = 256 consecutive samples of one item.  The (256 + k - 1) x C input rows are staged in shared memory with
// coalesced loads (row stride C + 1: lane-per-row reads are bank-conflict free); the per-sample accumulation order (taps
// outer, channels inner, sequential fma) is that of a plain loop, so results do not depend on the staging.
__global__ void __launch_bounds__(256) conv_post_kernel(const float* __restrict__ xin, const float* __restrict__ w,
                                                        float bias, float* __restrict__ wav, int C, int T, int k) {
  extern __shared__ float sw[];   // [k][C] weights, then [256 + k - 1][C + 1] input rows
  float* tile = sw + k * C;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * blockDim.x;
  const int half = (k - 1) / 2;
  const int R = blockDim.x + k - 1, C4 = C / 4, ld = C + 1;
  for (int i = threadIdx.x; i < k * C; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < R * C4; i += blockDim.x) {
    const int row = i / C4, c4 = i - row * C4;
    const int tt = t0 - half + row;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);     // zero padding (hifigan.py:138-139, padding = 3)
    if (tt >= 0 && tt < T) v = __ldg(reinterpret_cast<const float4*>(xin + (static_cast<size_t>(b) * T + tt) * C) + c4);
    float* d = tile + row * ld + 4 * c4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= T) return;
  float acc = 0.f;
  for (int j = 0; j < k; ++j) {
    const int tt = t + j - half;
    if (tt < 0 || tt >= T) continue;
    const float* row = tile + (threadIdx.x + j) * ld;
    const float* wj = sw + j * C;
#pragma unroll 8
    for (int c = 0; c < C; ++c) acc = fmaf(row[c], wj[c], acc);
  }
  wav[static_cast<size_t>(b) * T + t] = tanhf(acc + bias);
}

__global__ void __launch_bounds__(256) voc_f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n4) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n4) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    const float f[4] = {v.x, v.y, v.z, v.w};
    st_vec<4>(dst + 4 * i, f);
  }
}

struct ConvW {
  void* W = nullptr; float* bias = nullptr;
  int Cin = 0, N = 0, ntaps = 0, KB = 64, Kp = 0, BN = 0;
  int offs[kMaxTaps] = {0};
  CUtensorMap map{};
};

}  // namespace fse

using namespace fse;

struct fse_vocoder {
  fse_vocoder_config cfg{};
  bool bf16 = true, tc = true, loaded = false;     // operand type / tensor-core back end (fse_common.cuh: mode_is_bf16, mode_is_tc)
  int hop = 1;
  ConvW pre;
  std::vector<ConvW> ups;
  std::vector<ConvW> c1, c2;      // [stage][block][m]
  float* post_w = nullptr; float post_b = 0.f; int post_k = 7;
  struct MapEntry { const void* buf; int C, T, KB, rows; CUtensorMap map; };
  struct Plan { const void* ws = nullptr; int B = 0, T = 0; std::deque<MapEntry> cache; } plan;
  bool rb2 = false;       // ResBlock2 generator: each block is `x = conv_m(lrelu(x)) + x` for two dilated convs (hifigan.py:80-85)
  int nconv = 3;          // convs per block on the residual path: 3 conv pairs (ResBlock1) or 2 single convs (ResBlock2)
  bool multi_tile = true; // FSE_VOC_MT=0 disables multi-sub-tile jobs for narrow layers; FSE_VOC_MT=2: larger jobs
  int multi_tile_level = 2;
  bool shared_a = true;   // shared-activation schedule: a job's rows (128*MT + tap halo) are loaded once per channel block and the
                          // k (tap, sub-tile) operands are row-shifted descriptors of it (FSE_VOC_SHARED_A=0: one load per tap).
                          // Measured (B=32 x T=1024): vocoder 54 -> 46 ms once MMA issue and the epilogue stores were fixed.
  bool fuse_all = false;
  bool fuse_pair = true;  // ResBlock1 conv pairs as one kernel (resblock_fused.cuh) where the job fits shared memory / TMEM
                          // (FSE_VOC_FUSE=0: always two conv_gemm launches per pair)
  long long launches = 0;
  Profiler prof;
  void* host_ws = nullptr; size_t host_ws_bytes = 0;
};

namespace {

int stage_channels(const fse_vocoder* h, int i) { return h->cfg.upsample_initial_channel >> (i + 1); }

struct VWs {
  void* melb; void* ua; float* x; void* xa; float* y; void* ya; void* tmp; float* xs; float* y2; void* ya2;
  size_t bytes;
};
VWs vcarve(const fse_vocoder* h, void* base, int B, int T) {
  const size_t es = h->bf16 ? 2 : 4;
  size_t maxel = static_cast<size_t>(T) * h->cfg.upsample_initial_channel;   // conv_pre output
  size_t Ti = T;
  for (int i = 0; i < h->cfg.num_upsamples; ++i) {
    Ti *= h->cfg.upsample_rates[i];
    maxel = std::max(maxel, Ti * stage_channels(h, i));
  }
  maxel *= B;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  uint8_t* p = static_cast<uint8_t*>(base);
  VWs w{};
  size_t o;
  o = take(static_cast<size_t>(B) * T * h->cfg.n_mels * es); w.melb = p + o;
  o = take(maxel * es); w.ua = p + o;
  o = take(maxel * 4);  w.x = reinterpret_cast<float*>(p + o);
  o = take(maxel * es); w.xa = p + o;
  o = take(maxel * 4);  w.y = reinterpret_cast<float*>(p + o);
  o = take(maxel * es); w.ya = p + o;
  o = take(maxel * es); w.tmp = p + o;
  o = take(maxel * 4);  w.xs = reinterpret_cast<float*>(p + o);
  // second (y, ya) pair: inside a resblock the convs ping-pong (x, xa) -> (y, ya) -> (y2, ya2) -> sum.  The fused pair kernel reads
  // its input rows (with the conv halo) while other jobs of the SAME launch already write their output rows, so a conv pair must
  // never update its input in place.
  o = take(maxel * 4);  w.y2 = reinterpret_cast<float*>(p + o);
  o = take(maxel * es); w.ya2 = p + o;
  w.bytes = off;
  return w;
}

// fold weight norm: w = g * v / ||v|| with the norm over all dims but 0 (torch weight_norm dim=0)
std::vector<float> fold_wn(const TensorTable& tt, const std::string& name, int64_t d0, int64_t rest, int* rc) {
  std::vector<float> w(static_cast<size_t>(d0 * rest));
  if (tt.has(name + ".weight")) {
    const float* p = tt.get(name + ".weight", d0 * rest, rc);
    if (*rc == FSE_OK) w.assign(p, p + d0 * rest);
    return w;
  }
  const float* v = tt.get(name + ".weight_v", d0 * rest, rc);
  if (*rc != FSE_OK) return w;
  const float* g = tt.get(name + ".weight_g", d0, rc);
  if (*rc != FSE_OK) return w;
  for (int64_t i = 0; i < d0; ++i) {
    double ss = 0.0;
    for (int64_t j = 0; j < rest; ++j) ss += static_cast<double>(v[i * rest + j]) * v[i * rest + j];
    const float scale = static_cast<float>(g[i] / std::sqrt(ss));
    for (int64_t j = 0; j < rest; ++j) w[i * rest + j] = v[i * rest + j] * scale;
  }
  return w;
}

int finish_convw(fse_vocoder* h, ConvW& cw, const std::vector<float>& packed, const float* bias, int nbias) {
  FSE_TRY(upload_operand(packed, h->bf16, &cw.W, h->cfg.mode == FSE_MODE_TC_TF32));
  FSE_TRY(upload_f32(std::vector<float>(bias, bias + nbias), &cw.bias));
  if (h->tc) FSE_TRY(make_map_w(&cw.map, cw.W, cw.Kp, cw.N, cw.KB, cw.BN, h->bf16 ? 2 : 4));
  return FSE_OK;
}

// Conv1d weight [Cout, Cin, k] -> packed [Cout, ntaps * nkb * KB]
int pack_conv(fse_vocoder* h, const TensorTable& tt, const std::string& name, int Cout, int Cin, int k, int dil, ConvW& cw) {
  int rc = FSE_OK;
  std::vector<float> w = fold_wn(tt, name, Cout, static_cast<int64_t>(Cin) * k, &rc);
  if (rc) return rc;
  const float* bias = tt.get(name + ".bias", Cout, &rc);
  if (rc) return rc;
  if (k > kMaxTaps || k % 2 == 0) return fail(FSE_EINVAL, "%s: kernel size %d unsupported", name.c_str(), k);
  const int kbw = mode_kb(h->cfg.mode);           // widest k-block of the operand type (128-byte rows), else half of it
  cw.Cin = Cin; cw.N = Cout; cw.ntaps = k; cw.KB = Cin % kbw == 0 || Cin > kbw ? kbw : kbw / 2;
  const int nkb = (Cin + cw.KB - 1) / cw.KB;
  cw.Kp = k * nkb * cw.KB;
  cw.BN = Cout <= 256 ? Cout : 256;
  for (int j = 0; j < k; ++j) cw.offs[j] = (j - (k - 1) / 2) * dil;
  std::vector<float> p(static_cast<size_t>(Cout) * cw.Kp, 0.f);
  for (int o = 0; o < Cout; ++o)
    for (int c = 0; c < Cin; ++c)
      for (int j = 0; j < k; ++j) p[static_cast<size_t>(o) * cw.Kp + j * nkb * cw.KB + c] = w[(static_cast<size_t>(o) * Cin + c) * k + j];
  return finish_convw(h, cw, p, bias, Cout);
}

// ConvTranspose1d weight [Cin, Cout, k], stride u -> packed [u*Cout, ntaps*nkb*KB], tap m <-> kernel index r + m*u, offset -m
int pack_up(fse_vocoder* h, const TensorTable& tt, const std::string& name, int Cin, int Cout, int k, int u, ConvW& cw) {
  int rc = FSE_OK;
  std::vector<float> w = fold_wn(tt, name, Cin, static_cast<int64_t>(Cout) * k, &rc);   // dim 0 = Cin for ConvTranspose1d
  if (rc) return rc;
  const float* bias = tt.get(name + ".bias", Cout, &rc);
  if (rc) return rc;
  const int ntaps = (k + u - 1) / u;
  if (ntaps > kMaxTaps) return fail(FSE_EINVAL, "%s: too many taps", name.c_str());
  const int KB = mode_kb(h->cfg.mode);
  cw.Cin = Cin; cw.N = u * Cout; cw.ntaps = ntaps; cw.KB = KB;
  const int nkb = (Cin + KB - 1) / KB;
  cw.Kp = ntaps * nkb * KB;
  cw.BN = cw.N % 256 == 0 ? 256 : (cw.N % 128 == 0 ? 128 : (cw.N % 64 == 0 ? 64 : 32));
  for (int m = 0; m < ntaps; ++m) cw.offs[m] = -m;
  std::vector<float> p(static_cast<size_t>(cw.N) * cw.Kp, 0.f);
  for (int r = 0; r < u; ++r)
    for (int co = 0; co < Cout; ++co)
      for (int m = 0; m < ntaps; ++m) {
        const int j = r + m * u;
        if (j >= k) continue;
        for (int ci = 0; ci < Cin; ++ci)
          p[(static_cast<size_t>(r) * Cout + co) * cw.Kp + m * nkb * KB + ci] = w[(static_cast<size_t>(ci) * Cout + co) * k + j];
      }
  return finish_convw(h, cw, p, bias, Cout);
}

// Tensor map of an activation buffer with a given box height, cached in the plan (cleared when the workspace changes).
int get_act_map(fse_vocoder* h, const void* buf, int C, int T, int B, int KB, int rows, const CUtensorMap** out) {
  for (auto& e : h->plan.cache)
    if (e.buf == buf && e.C == C && e.T == T && e.KB == KB && e.rows == rows) { *out = &e.map; return FSE_OK; }
  if (h->plan.cache.size() > 256) h->plan.cache.clear();     // caller-owned operands (fp32 mel) may move between calls
  h->plan.cache.emplace_back();
  auto& e = h->plan.cache.back();
  e.buf = buf; e.C = C; e.T = T; e.KB = KB; e.rows = rows;
  FSE_TRY(make_map_act(&e.map, buf, C, T, B, KB, rows, h->bf16 ? 2 : 4));
  *out = &e.map;
  return FSE_OK;
}

template <typename TOp, class Epi>
int run_conv(fse_vocoder* h, const ConvW& cw, const void* A, int B, int Trows, int Tsrc, const Epi& epi, cudaStream_t st, int kind) {
  ConvGemmParams p = make_params(B, Trows, Tsrc, cw.Cin, cw.ntaps, cw.offs, 0, cw.N, cw.KB);
  GemmOperands op; op.A0 = A; op.W = cw.W; op.mW = &cw.map; op.BN = cw.BN;
  if (h->tc) {
    // every tap of a conv reads the same activation tile shifted by whole frames: load it once per channel block
    // (with the tap halo) and feed the taps from row-shifted descriptors -> activation ingest / ntaps
    // narrow layers (C_out <= 128, one n-tile): one job = several 128-frame sub-tiles against the same weight tiles
    int rows = kTileM, mt = 1;
    if (h->multi_tile && cw.BN == cw.N) {
      if (h->multi_tile_level >= 3) mt = cw.BN <= 32 ? 8 : (cw.BN <= 64 ? 4 : (cw.BN <= 128 ? 4 : 2));   // wide layers: one accumulator
      else if (h->multi_tile_level >= 2) mt = cw.BN <= 32 ? 8 : (cw.BN <= 64 ? 4 : (cw.BN <= 128 ? 2 : 1));
      else mt = cw.BN <= 32 ? 4 : (cw.BN <= 128 ? 2 : 1);
    }
    if (h->shared_a && cw.ntaps >= 3 && enable_shared_a(p, mt, h->bf16 ? 2 : 4)) rows = p.Rbox;
    else p.MT = mt;
    FSE_TRY(get_act_map(h, A, cw.Cin, Tsrc, B, cw.KB, rows, &op.mA0));
  }
  return run_conv_gemm<TOp>(h->cfg.mode, p, op, epi, st, LaunchCtx{&h->launches, &h->prof, kind});
}

// Shape of a fused conv-pair job (resblock_fused.cuh) for TWO resident CTAs per SM: the largest MT whose two accumulators fit 256
// TMEM columns and whose tiles (input slot(s) with the conv1 halo, the intermediate U which doubles as the E2 scratch, >= 3 weight
// stages) fit 112 KB of shared memory.  No such shape (C = 256; C = 128 with fp32 operands) -> the pair runs as two conv_gemm launches.
constexpr int kPairNE = 4;
bool plan_pair(int C, int k, int dil, int KB, int es, PairParams* out) {
  const int RB = KB * es;
  if (C % 32 != 0 || C > 256 || k < 2 || k > kMaxTaps || (RB != 128 && RB != 64)) return false;
  const int nkb = (C + KB - 1) / KB;
  const int h2 = (k - 1) / 2, h1 = h2 * dil;
  const int budget = 112 * 1024 - 1024 - 256;
  for (int MT = 4; MT >= 1; MT >>= 1) {
    if (2 * MT * C > 256) continue;
    const int need = kTileM * MT + 2 * h1;
    int nload = (need + 255) / 256, box = 0;
    for (;; ++nload) {
      box = ((need + nload - 1) / nload + 7) / 8 * 8;
      if (box <= 256) break;
    }
    const int a_slot = static_cast<int>(align_up(static_cast<size_t>(box) * nload * RB, 1024));
    const int u_rows = (kTileM * MT + k - 1 + 7) / 8 * 8;
    const int u_kb = static_cast<int>(align_up(static_cast<size_t>(u_rows) * RB, 1024));
    if (nkb * u_kb < kPairNE * 4096) continue;                 // U doubles as the transposition scratch of E2
    const int w_stage = tc_b_stage_bytes(C, RB);
    const int a_slots = nkb >= 2 ? 2 : 1;                      // one channel block per conv: the next job's load follows C1 directly
    const int left = budget - a_slots * a_slot - nkb * u_kb;
    if (left < 3 * w_stage) continue;
    PairParams p{};
    p.C = C; p.k = k; p.dil = dil; p.nkb = nkb; p.MT = MT; p.Rout = kTileM * MT - (k - 1);
    p.Rbox = box; p.nload = nload; p.a_slots = a_slots; p.a_slot_bytes = a_slot;
    p.stages = left / w_stage > 6 ? 6 : left / w_stage; p.w_stage_bytes = w_stage; p.u_kb_bytes = u_kb;
    p.slope1 = 0.1f;
    *out = p;
    return true;
  }
  return false;
}

template <typename TOp, int KB, class Epi>
int launch_pair(fse_vocoder* h, const PairParams& p, const CUtensorMap* mA, const ConvW& c1, const ConvW& c2, const Epi& epi, cudaStream_t st) {
  static bool attr_set[kMaxDevices] = {};
  int dev = 0;
  FSE_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(FSE_EINVAL, "device ordinal %d out of range", dev);
  auto kern = resblock_pair_kernel<TOp, KB, kPairNE, Epi>;
  if (!attr_set[dev]) {
    FSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    FSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set[dev] = true;
  }
  const size_t smem = 1024 + static_cast<size_t>(p.a_slots) * p.a_slot_bytes + static_cast<size_t>(p.stages) * p.w_stage_bytes +
                      static_cast<size_t>(p.nkb) * p.u_kb_bytes + 256;
  const int jobs = p.B * ((p.T + p.Rout - 1) / p.Rout);
  const int num_sms = device_sm_count(dev);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(jobs < 2 * num_sms ? jobs : 2 * num_sms);     // persistent: two CTAs per SM
  cfg.blockDim = dim3(64 + 32 * kPairNE);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FSE_CUDA(cudaLaunchKernelEx(&cfg, kern, *mA, c1.map, c2.map, p, static_cast<const float*>(c1.bias), epi));
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

// y = x + conv2(lrelu(conv1(xa))) as one launch when the job fits; *fused tells the caller whether it ran
template <typename TOp, class Epi>
int run_pair(fse_vocoder* h, const ConvW& c1, const ConvW& c2, const void* A, int B, int T, const Epi& epi, cudaStream_t st, bool* fused) {
  *fused = false;
  if (!h->tc || !h->fuse_pair || c1.KB != c2.KB || c1.ntaps != c2.ntaps || c1.ntaps < 2 || c1.N != c1.Cin || c2.N != c1.N) return FSE_OK;
  const int es = h->bf16 ? 2 : 4;
  // Measured per stage on B200 (profiles/r02_vocoder_launches_*.csv): the fused kernel wins where the two-kernel form is bound by
  // HBM bytes or by its single MMA-issue chain on narrow tiles (C = 32 in both kinds: 2.0-2.4x with fp32 operands; C = 64 with
  // bf16), and is ~10 % slower for C = 64 with fp32 operands, where only MT = 1 fits next to a second resident CTA.
  // FSE_VOC_FUSE=2 fuses every pair that has a shape.
  if (!h->fuse_all && c1.N * es > 128) return FSE_OK;
  const int dil = c1.ntaps > 1 ? c1.offs[1] - c1.offs[0] : 1;
  PairParams p{};
  if (!plan_pair(c1.N, c1.ntaps, dil, c1.KB, es, &p)) return FSE_OK;
  p.B = B; p.T = T;
  const CUtensorMap* mA = nullptr;
  FSE_TRY(get_act_map(h, A, c1.Cin, T, B, c1.KB, p.Rbox, &mA));
  ++h->launches;
  h->prof.begin(3, st);
  int rc = FSE_OK;
  if constexpr (std::is_same<TOp, __nv_bfloat16>::value) {
    rc = c1.KB == 64 ? launch_pair<TOp, 64, Epi>(h, p, mA, c1, c2, epi, st) : launch_pair<TOp, 32, Epi>(h, p, mA, c1, c2, epi, st);
  } else {
    if (c1.KB != 32) { h->prof.end(st); --h->launches; return FSE_OK; }
    rc = launch_pair<TOp, 32, Epi>(h, p, mA, c1, c2, epi, st);
  }
  h->prof.end(st);
  *fused = rc == FSE_OK;
  return rc;
}

template <typename TOp>
int forward_impl(fse_vocoder* h, const float* mel, float* wav, int B, int T, void* ws, cudaStream_t st) {
  VWs w = vcarve(h, ws, B, T);
  const auto& cfg = h->cfg;
  const bool tc = h->tc;
  const int nu = cfg.num_upsamples, nk = cfg.num_kernels;
  if (tc && !(h->plan.ws == ws && h->plan.B == B && h->plan.T == T)) {
    h->plan.cache.clear();
    h->plan.ws = ws; h->plan.B = B; h->plan.T = T;
  }

  const void* mel_op = mel;
  if constexpr (std::is_same<TOp, __nv_bfloat16>::value) {
    const size_t n = static_cast<size_t>(B) * T * cfg.n_mels;
    voc_f32_to_bf16_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, st>>>(mel, static_cast<__nv_bfloat16*>(w.melb), n / 4);
    FSE_CUDA(cudaGetLastError());
    ++h->launches;
    mel_op = w.melb;
  }
  {  // conv_pre + leaky_relu(0.1) of stage 0 (hifigan.py:127-129)
    EpiAct<TOp> epi{h->pre.bias, static_cast<TOp*>(w.ua), h->pre.N, T, 0.1f};
    FSE_TRY((run_conv<TOp>(h, h->pre, mel_op, B, T, T, epi, st, 0)));
  }
  int Tin = T;
  for (int i = 0; i < nu; ++i) {
    const int u = cfg.upsample_rates[i], k = cfg.upsample_kernel_sizes[i], pad = (k - u) / 2;
    const int Cout = stage_channels(h, i), Tout = Tin * u;
    {
      EpiUp<TOp> epi{h->ups[i].bias, w.x, static_cast<TOp*>(w.xa), Cout, Tout, u, pad};
      FSE_TRY((run_conv<TOp>(h, h->ups[i], w.ua, B, Tin + 1, Tin, epi, st, 1)));
    }
    const bool last_stage = i == nu - 1;
    for (int j = 0; j < nk; ++j) {
      for (int m = 0; m < h->nconv; ++m) {
        const ConvW& a = h->c1[(i * nk + j) * 3 + m];
        const ConvW& c = h->rb2 ? a : h->c2[(i * nk + j) * 3 + m];
        const void* src1 = m == 0 ? w.xa : (m == 1 ? w.ya : w.ya2);
        const float* res_in = m == 0 ? w.x : (m == 1 ? w.y : w.y2);
        float* y_out = m == 0 ? w.y : w.y2;
        void* ya_out = m == 0 ? w.ya : w.ya2;
        const void* src2 = h->rb2 ? src1 : w.tmp;      // ResBlock2: the dilated conv itself carries the residual add
        {
          const int kind = m < h->nconv - 1 ? 0 : (nk == 1 ? 3 : (j == 0 ? 1 : (j == nk - 1 ? 3 : 2)));
          auto fill = [&](auto& epi) {
            epi.bias = c.bias; epi.res = res_in; epi.y = y_out; epi.ya = static_cast<TOp*>(ya_out);
            epi.xs = w.xs; epi.next_a = static_cast<TOp*>(w.ua); epi.final_f32 = last_stage ? w.xs : nullptr;
            epi.N = Cout; epi.T = Tout;
            epi.kind = kind;
            epi.num_kernels = static_cast<float>(nk);
            epi.slope_next = last_stage ? 0.01f : 0.1f;   // F.leaky_relu default slope before conv_post (hifigan.py:138)
          };
          // ResBlock1: conv1 -> lrelu -> conv2 -> + residual as ONE kernel when the job fits (resblock_fused.cuh), else conv1 to
          // `tmp` (EpiAct) followed by conv2 with the residual epilogue
          auto conv1_unfused = [&]() -> int {
            if (h->rb2) return FSE_OK;
            EpiAct<TOp> e1{a.bias, static_cast<TOp*>(w.tmp), Cout, Tout, 0.1f};
            return run_conv<TOp>(h, a, src1, B, Tout, Tout, e1, st, 2);
          };
          bool fused = false;
          if (kind >= 2 && !(kind == 3 && nk == 1)) {
            EpiResAdd<TOp, true> epi{};
            fill(epi);
            if (!h->rb2) FSE_TRY((run_pair<TOp>(h, a, c, src1, B, Tout, epi, st, &fused)));
            if (!fused) {
              FSE_TRY(conv1_unfused());
              FSE_TRY((run_conv<TOp>(h, c, src2, B, Tout, Tout, epi, st, 3)));
            }
          } else {
            EpiResAdd<TOp, false> epi{};
            fill(epi);
            if (!h->rb2) FSE_TRY((run_pair<TOp>(h, a, c, src1, B, Tout, epi, st, &fused)));
            if (!fused) {
              FSE_TRY(conv1_unfused());
              FSE_TRY((run_conv<TOp>(h, c, src2, B, Tout, Tout, epi, st, 3)));
            }
          }
        }
      }
    }
    Tin = Tout;
  }
  const int Cl = stage_channels(h, nu - 1);
  h->prof.begin(4, st);
  const size_t post_smem = (static_cast<size_t>(h->post_k) * Cl + (256 + h->post_k - 1) * (Cl + 1)) * sizeof(float);
  if (post_smem > 48 * 1024) {
    if (post_smem > 200 * 1024) return fail(FSE_EINVAL, "conv_post: %d channels x %d taps do not fit shared memory", Cl, h->post_k);
    FSE_CUDA(cudaFuncSetAttribute(conv_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(post_smem)));
  }
  conv_post_kernel<<<dim3((Tin + 255) / 256, B), 256, post_smem, st>>>(w.xs, h->post_w, h->post_b, wav, Cl, Tin, h->post_k);
  h->prof.end(st);
  FSE_CUDA(cudaGetLastError());
  ++h->launches;
  return FSE_OK;
}

}  // namespace

extern "C" {

int fse_vocoder_create(const fse_vocoder_config* cfg, fse_vocoder** out) {
  if (!cfg || !out) return fail(FSE_EINVAL, "null argument");
  if (cfg->num_upsamples <= 0 || cfg->num_upsamples > 8 || cfg->num_kernels <= 0 || cfg->num_kernels > 4)
    return fail(FSE_EINVAL, "num_upsamples / num_kernels out of range");
  if (cfg->n_mels % 8 != 0) return fail(FSE_EINVAL, "n_mels must be a multiple of 8");
  if (cfg->resblock < 0 || cfg->resblock > 2) return fail(FSE_EINVAL, "resblock must be 1 or 2 (got %d)", cfg->resblock);
  if (cfg->mode < 0 || cfg->mode > 3) return fail(FSE_EINVAL, "unknown mode %d", cfg->mode);
  int hop = 1;
  for (int i = 0; i < cfg->num_upsamples; ++i) {
    const int u = cfg->upsample_rates[i], k = cfg->upsample_kernel_sizes[i];
    if (u <= 0 || k < u || (k - u) % 2 != 0) return fail(FSE_EINVAL, "upsample stage %d: need k >= u and (k-u) even", i);
    const int c = cfg->upsample_initial_channel >> (i + 1);
    if (c < 32 || c % 32 != 0) return fail(FSE_EINVAL, "stage %d has %d channels; need a multiple of 32", i, c);
    hop *= u;
  }
  int dev = 0;
  FSE_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  FSE_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(FSE_ECUDA, "device is sm_%d%d; this library is built for sm_100a only (no fallback)", prop.major, prop.minor);
  auto* h = new fse_vocoder();
  h->cfg = *cfg;
  h->bf16 = mode_is_bf16(cfg->mode);
  h->tc = mode_is_tc(cfg->mode);
  h->hop = hop;
  h->rb2 = cfg->resblock == 2;
  h->nconv = h->rb2 ? 2 : 3;
  if (const char* e = getenv("FSE_VOC_SHARED_A")) h->shared_a = atoi(e) != 0;
  if (const char* e = getenv("FSE_VOC_FUSE")) { h->fuse_pair = atoi(e) != 0; h->fuse_all = atoi(e) == 2; }
  if (const char* e = getenv("FSE_VOC_MT")) { h->multi_tile = atoi(e) != 0; h->multi_tile_level = atoi(e); }
  *out = h;
  return FSE_OK;
}

void fse_vocoder_destroy(fse_vocoder* h) {
  if (!h) return;
  auto freew = [](ConvW& c) { if (c.W) cudaFree(c.W); if (c.bias) cudaFree(c.bias); };
  freew(h->pre);
  for (auto& c : h->ups) freew(c);
  for (auto& c : h->c1) freew(c);
  for (auto& c : h->c2) freew(c);
  if (h->post_w) cudaFree(h->post_w);
  if (h->host_ws) cudaFree(h->host_ws);
  delete h;
}

int fse_vocoder_load_weights(fse_vocoder* h, const fse_tensor* tensors, int32_t n) {
  if (!h || !tensors || n <= 0) return fail(FSE_EINVAL, "null argument");
  if (h->loaded) return fail(FSE_ESTATE, "weights already loaded");
  const auto& cfg = h->cfg;
  TensorTable tt(tensors, n);
  const int C0 = cfg.upsample_initial_channel, nu = cfg.num_upsamples, nk = cfg.num_kernels;
  FSE_TRY(pack_conv(h, tt, "conv_pre", C0, cfg.n_mels, 7, 1, h->pre));
  h->ups.resize(nu); h->c1.resize(nu * nk * 3); h->c2.resize(nu * nk * 3);
  int Cin = C0;
  for (int i = 0; i < nu; ++i) {
    const int Cout = stage_channels(h, i);
    FSE_TRY(pack_up(h, tt, "ups." + std::to_string(i), Cin, Cout, cfg.upsample_kernel_sizes[i], cfg.upsample_rates[i], h->ups[i]));
    for (int j = 0; j < nk; ++j)
      for (int m = 0; m < h->nconv; ++m) {
        const std::string pre = "resblocks." + std::to_string(i * nk + j) + ".";
        if (h->rb2) {        // ResBlock2.convs (hifigan.py:71-76)
          FSE_TRY(pack_conv(h, tt, pre + "convs." + std::to_string(m), Cout, Cout, cfg.resblock_kernel_sizes[j],
                            cfg.resblock_dilations[j][m], h->c1[(i * nk + j) * 3 + m]));
          continue;
        }
        FSE_TRY(pack_conv(h, tt, pre + "convs1." + std::to_string(m), Cout, Cout, cfg.resblock_kernel_sizes[j],
                          cfg.resblock_dilations[j][m], h->c1[(i * nk + j) * 3 + m]));
        FSE_TRY(pack_conv(h, tt, pre + "convs2." + std::to_string(m), Cout, Cout, cfg.resblock_kernel_sizes[j], 1,
                          h->c2[(i * nk + j) * 3 + m]));
      }
    Cin = Cout;
  }
  {  // conv_post [1, Cl, 7] -> [7][Cl] fp32
    int rc = FSE_OK;
    std::vector<float> w = fold_wn(tt, "conv_post", 1, static_cast<int64_t>(Cin) * 7, &rc);
    if (rc) return rc;
    const float* b = tt.get("conv_post.bias", 1, &rc);
    if (rc) return rc;
    std::vector<float> p(static_cast<size_t>(7) * Cin);
    for (int c = 0; c < Cin; ++c) for (int j = 0; j < 7; ++j) p[static_cast<size_t>(j) * Cin + c] = w[static_cast<size_t>(c) * 7 + j];
    FSE_TRY(upload_f32(p, &h->post_w));
    h->post_b = b[0];
    h->post_k = 7;
  }
  h->loaded = true;
  return FSE_OK;
}

int64_t fse_vocoder_workspace_bytes(const fse_vocoder* h, int32_t B, int32_t T) {
  if (!h || B <= 0 || T <= 0) return 0;
  return static_cast<int64_t>(vcarve(h, nullptr, B, T).bytes);
}

int fse_vocoder_forward(fse_vocoder* h, const float* mel, float* wav, int32_t B, int32_t T, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h || !mel || !wav) return fail(FSE_EINVAL, "null argument");
  if (!h->loaded) return fail(FSE_ESTATE, "weights not loaded");
  if (B <= 0 || T <= 0) return fail(FSE_EINVAL, "B and T must be positive");
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 1023)) return fail(FSE_EINVAL, "workspace must be non-null and 1024-byte aligned");
  if (workspace_bytes < fse_vocoder_workspace_bytes(h, B, T)) return fail(FSE_EINVAL, "workspace too small");
  if ((static_cast<size_t>(B) * T * h->cfg.n_mels) % 4 != 0) return fail(FSE_EINVAL, "B*T*n_mels must be a multiple of 4");
  h->launches = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return h->bf16 ? forward_impl<__nv_bfloat16>(h, mel, wav, B, T, workspace, st) : forward_impl<float>(h, mel, wav, B, T, workspace, st);
}

int fse_vocoder_forward_host(fse_vocoder* h, const float* mel, float* wav, int32_t B, int32_t T) {
  if (!h || !mel || !wav) return fail(FSE_EINVAL, "null argument");
  if (!h->loaded) return fail(FSE_ESTATE, "weights not loaded");
  if (B <= 0 || T <= 0) return fail(FSE_EINVAL, "B and T must be positive");
  const size_t ws_bytes = static_cast<size_t>(fse_vocoder_workspace_bytes(h, B, T));
  const size_t mel_b = align_up(static_cast<size_t>(B) * T * h->cfg.n_mels * 4, 1024);
  const size_t wav_b = align_up(static_cast<size_t>(B) * T * h->hop * 4, 1024);
  const size_t total = ws_bytes + mel_b + wav_b;
  if (h->host_ws_bytes < total) {
    if (h->host_ws) cudaFree(h->host_ws);
    h->host_ws = nullptr; h->host_ws_bytes = 0;
    FSE_CUDA(cudaMalloc(&h->host_ws, total));
    h->host_ws_bytes = total;
  }
  uint8_t* p = static_cast<uint8_t*>(h->host_ws);
  float* d_mel = reinterpret_cast<float*>(p + ws_bytes);
  float* d_wav = reinterpret_cast<float*>(p + ws_bytes + mel_b);
  cudaStream_t st = nullptr;
  FSE_CUDA(cudaMemcpyAsync(d_mel, mel, static_cast<size_t>(B) * T * h->cfg.n_mels * 4, cudaMemcpyHostToDevice, st));
  FSE_TRY(fse_vocoder_forward(h, d_mel, d_wav, B, T, p, static_cast<int64_t>(ws_bytes), st));
  FSE_CUDA(cudaMemcpyAsync(wav, d_wav, static_cast<size_t>(B) * T * h->hop * 4, cudaMemcpyDeviceToHost, st));
  FSE_CUDA(cudaStreamSynchronize(st));
  return FSE_OK;
}

int64_t fse_vocoder_last_launches(const fse_vocoder* h) { return h ? h->launches : 0; }

int fse_vocoder_profile(fse_vocoder* h, int32_t enable) {
  if (!h) return fail(FSE_EINVAL, "null handle");
  h->prof.enable(enable != 0);
  return FSE_OK;
}
int fse_vocoder_profile_read(fse_vocoder* h, double* ms_by_kind, int64_t* launches_by_kind) {
  if (!h || !ms_by_kind || !launches_by_kind) return fail(FSE_EINVAL, "null argument");
  return h->prof.read(ms_by_kind, launches_by_kind);
}

}  // extern "C"
