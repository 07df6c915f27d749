// Host-side plumbing shared by denoiser.cu and hifigan.cu: error reporting, device buffers,
// TMA tensor-map encoding and the conv_gemm launcher that picks the back end.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "../../include/fse_b200.h"
#include "conv_gemm.cuh"

namespace fse {

// ------------------------------------------------------------------ errors
inline std::string& last_error_ref() {
  static thread_local std::string msg;
  return msg;
}
inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}
#define FSE_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::fse::fail(FSE_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define FSE_TRY(expr)          \
  do {                         \
    int _rc = (expr);          \
    if (_rc != FSE_OK) return _rc; \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------ weights
struct TensorTable {
  std::unordered_map<std::string, const fse_tensor*> map;
  TensorTable(const fse_tensor* t, int n) {
    for (int i = 0; i < n; ++i) map[t[i].name] = &t[i];
  }
  const float* get(const std::string& name, int64_t numel, int* rc) const {
    auto it = map.find(name);
    if (it == map.end()) {
      *rc = fail(FSE_EINVAL, "missing weight tensor '%s'", name.c_str());
      return nullptr;
    }
    if (it->second->numel != numel) {
      *rc = fail(FSE_EINVAL, "weight '%s' has %lld elements, expected %lld", name.c_str(),
                 static_cast<long long>(it->second->numel), static_cast<long long>(numel));
      return nullptr;
    }
    return it->second->data;
  }
  bool has(const std::string& name) const { return map.count(name) != 0; }
};

inline uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

// fp32 -> tf32 (10 explicit mantissa bits), round to nearest even, kept in an fp32 container: the tensor core ignores the low
// 13 bits, so rounding the (constant) weights once at load time halves their representation error versus truncation.
inline float f32_to_tf32_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return f;
  u += 0xFFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  memcpy(&f, &u, 4);
  return f;
}

// Upload a host fp32 matrix as the operand type of the mode (bf16 or fp32; tf32 = fp32 rounded to tf32).
inline int upload_operand(const std::vector<float>& host, bool bf16, void** dptr, bool tf32 = false) {
  if (tf32) {
    std::vector<float> tmp(host.size());
    for (size_t i = 0; i < host.size(); ++i) tmp[i] = f32_to_tf32_rne(host[i]);
    FSE_CUDA(cudaMalloc(dptr, tmp.size() * 4));
    FSE_CUDA(cudaMemcpy(*dptr, tmp.data(), tmp.size() * 4, cudaMemcpyHostToDevice));
    return FSE_OK;
  }
  if (bf16) {
    std::vector<uint16_t> tmp(host.size());
    for (size_t i = 0; i < host.size(); ++i) tmp[i] = f32_to_bf16_rne(host[i]);
    FSE_CUDA(cudaMalloc(dptr, tmp.size() * 2));
    FSE_CUDA(cudaMemcpy(*dptr, tmp.data(), tmp.size() * 2, cudaMemcpyHostToDevice));
  } else {
    FSE_CUDA(cudaMalloc(dptr, host.size() * 4));
    FSE_CUDA(cudaMemcpy(*dptr, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
  }
  return FSE_OK;
}
inline int upload_f32(const std::vector<float>& host, float** dptr) {
  FSE_CUDA(cudaMalloc(reinterpret_cast<void**>(dptr), host.size() * 4));
  FSE_CUDA(cudaMemcpy(*dptr, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
  return FSE_OK;
}

// ------------------------------------------------------------------ TMA tensor maps
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline int get_encode_fn(PFN_tmapEncodeTiled* out) {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    FSE_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) return fail(FSE_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  *out = fn;
  return FSE_OK;
}
inline bool mode_is_tc(int mode) { return mode == FSE_MODE_TC_BF16 || mode == FSE_MODE_TC_TF32; }
inline bool mode_is_bf16(int mode) { return mode == FSE_MODE_TC_BF16 || mode == FSE_MODE_SIMT_BF16; }
// widest k-block (in channels) of an operand type: 128 bytes of K per row
inline int mode_kb(int mode) { return mode_is_bf16(mode) ? 64 : 32; }

// activation [B, T, C] channels-last (bf16, es = 2; fp32 for the tf32 kind, es = 4): dims (C, T, B), box (KB, rows, 1);
// out-of-range frames/channels read 0.
inline int make_map_act(CUtensorMap* m, const void* ptr, int C, int T, int B, int KB, int rows = kTileM, int es = 2) {
  PFN_tmapEncodeTiled enc;
  FSE_TRY(get_encode_fn(&enc));
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(C) * es, static_cast<cuuint64_t>(T) * C * es};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(KB), static_cast<cuuint32_t>(rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, KB * es == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FSE_ECUDA, "cuTensorMapEncodeTiled(act C=%d T=%d B=%d) failed: %d", C, T, B, (int)r);
  return FSE_OK;
}
// packed weight [N, Kp] row-major (K contiguous; bf16 or fp32): dims (Kp, N), box (KB, BN).
inline int make_map_w(CUtensorMap* m, const void* ptr, int Kp, int N, int KB, int BN, int es = 2) {
  PFN_tmapEncodeTiled enc;
  FSE_TRY(get_encode_fn(&enc));
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(Kp), static_cast<cuuint64_t>(N)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(Kp) * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(KB), static_cast<cuuint32_t>(BN)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, KB * es == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FSE_ECUDA, "cuTensorMapEncodeTiled(w Kp=%d N=%d) failed: %d", Kp, N, (int)r);
  return FSE_OK;
}

// ------------------------------------------------------------------ per-kind kernel timing (opt-in)
// CUDA events recorded on the launch stream around each kernel, summed per kernel kind; used by
// bench.py for the roofline object (B200_PROFILING.md: time on the launching stream).
constexpr int kProfKinds = 8;
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> kind;
  size_t used = 0;
  void enable(bool e) {
    on = e;
    used = 0;
    kind.clear();
  }
  void begin(int k, cudaStream_t st) {
    if (!on) return;
    if (ev.size() < 2 * (used + 1)) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { on = false; return; }
      ev.push_back(a); ev.push_back(b);
    }
    kind.push_back(k);
    cudaEventRecord(ev[2 * used], st);
  }
  void end(cudaStream_t st) {
    if (!on) return;
    cudaEventRecord(ev[2 * used + 1], st);
    ++used;
  }
  int read(double* ms, int64_t* counts) {
    for (int i = 0; i < kProfKinds; ++i) { ms[i] = 0.0; counts[i] = 0; }
    for (size_t i = 0; i < used; ++i) {
      FSE_CUDA(cudaEventSynchronize(ev[2 * i + 1]));
      float t = 0.f;
      FSE_CUDA(cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]));
      const int k = kind[i] < kProfKinds ? kind[i] : kProfKinds - 1;
      ms[k] += t; counts[k] += 1;
    }
    return FSE_OK;
  }
  ~Profiler() { for (auto e : ev) cudaEventDestroy(e); }
};
struct LaunchCtx {
  long long* launches = nullptr;
  Profiler* prof = nullptr;
  int kind = 0;
};

constexpr int kMaxDevices = 64;
inline int device_sm_count(int dev) {
  static int sms[kMaxDevices] = {};
  if (dev < 0 || dev >= kMaxDevices) return 0;
  if (sms[dev] == 0 && cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms[dev] = 0;
  return sms[dev];
}

// ------------------------------------------------------------------ conv_gemm launcher
struct GemmOperands {
  const void* A0 = nullptr;   // [B, Tsrc, C0] operand type
  const void* A1 = nullptr;   // [B, Tsrc, C1]
  const void* W = nullptr;    // [N, Kp]
  const CUtensorMap* mA0 = nullptr;
  const CUtensorMap* mA1 = nullptr;
  const CUtensorMap* mW = nullptr;
  int BN = 256;               // tensor-core tile width (N % BN == 0)
};

inline ConvGemmParams make_params(int B, int Trows, int Tsrc, int C0, int ntaps, const int* offs, int C1, int N, int KB) {
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.Trows = Trows; p.Tsrc = Tsrc; p.C0 = C0; p.C1 = C1; p.ntaps = ntaps; p.ld0 = C0; p.MT = 1;
  for (int i = 0; i < ntaps; ++i) p.tap_off[i] = offs[i];
  p.KB = KB;
  p.nkb0 = (C0 + KB - 1) / KB;
  p.nkb1 = (C1 + KB - 1) / KB;
  p.N = N;
  p.Kp = (ntaps * p.nkb0 + p.nkb1) * KB;
  return p;
}

// Switch a parameter block to the shared-A schedule: each channel block of source 0 is loaded once per job (128*MT rows
// plus the tap halo, as nload boxes of Rbox rows) and every (tap, sub-tile) reads a row-shifted descriptor of that copy.
// The A0 tensor map must be made with rows = p.Rbox.  Measured on B200: the swizzle is a function of the absolute smem
// address, so a row-shifted start needs NO descriptor base-offset (tools/shared_a_check.py).
inline bool enable_shared_a(ConvGemmParams& p, int MT = 1, int es = 2) {      // es = bytes per operand element (2: bf16, 4: fp32 / tf32)
  int lo = p.tap_off[0], hi = p.tap_off[0];
  for (int i = 1; i < p.ntaps; ++i) { lo = p.tap_off[i] < lo ? p.tap_off[i] : lo; hi = p.tap_off[i] > hi ? p.tap_off[i] : hi; }
  if (p.ntaps < 2 || MT < 1 || (MT > 1 && p.nkb1 > 0)) return false;
  const int need = kTileM * MT + hi - lo;
  int nload = (need + 255) / 256, box = 0;
  for (;; ++nload) {
    box = ((need + nload - 1) / nload + 7) / 8 * 8;       // boxes start on a swizzle-atom boundary (8 rows)
    if (box <= 256) break;
  }
  if (static_cast<size_t>(box) * nload * p.KB * es > 100 * 1024) return false;     // two slots must fit next to the weight ring
  p.shared_a = 1; p.MT = MT; p.off_min = lo; p.Rrows = need; p.Rbox = box; p.nload = nload; p.bo_mode = 0;
  return true;
}

template <typename TOp, int KB, int CH, class Epi>
inline int launch_tc(const ConvGemmParams& p, const GemmOperands& op, const Epi& epi, cudaStream_t st) {
  static bool attr_set[kMaxDevices] = {};       // per device ordinal: function attributes are per context
  int dev = 0;
  FSE_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(FSE_EINVAL, "device ordinal %d out of range", dev);
  constexpr int RB = KB * static_cast<int>(sizeof(TOp));
  const int nkb = p.ntaps * p.nkb0 + p.nkb1;
  int stages, a_slots, a_slot_bytes;
  constexpr int scratch = (CH == 32 && epi_transposed<Epi>::value) ? kEpiScratchBytes : 0;
  const int budget = 225 * 1024 - scratch;
  if (p.shared_a) {
    const int rows = p.Rbox * p.nload > kTileM ? p.Rbox * p.nload : kTileM;
    a_slot_bytes = static_cast<int>(align_up(static_cast<size_t>(rows) * RB, 1024));
    a_slots = p.nkb0 + p.nkb1 < 2 ? 2 : (p.nkb0 + p.nkb1 < 3 ? p.nkb0 + p.nkb1 : 3);   // >= 2: the next job's load overlaps this job's MMAs
    while (a_slots > 2 && budget - a_slots * a_slot_bytes < 3 * tc_b_stage_bytes(op.BN, RB)) --a_slots;
    stages = (budget - a_slots * a_slot_bytes) / tc_b_stage_bytes(op.BN, RB);
    if (stages > 8) stages = 8;
  } else {
    a_slot_bytes = tc_a_stage_bytes(RB) * (p.MT > 0 ? p.MT : 1);
    stages = budget / (a_slot_bytes + tc_b_stage_bytes(op.BN, RB));
    if (stages > 6) stages = 6;
    a_slots = 0;   // = stages, set below
  }
  if (stages > nkb) stages = nkb;
  if (stages < 1) return fail(FSE_EINVAL, "conv_gemm: tile does not fit shared memory");
  if (!p.shared_a) a_slots = stages;
  auto kern = conv_gemm_tc_kernel<TOp, KB, CH, Epi>;
  if (!attr_set[dev]) {
    FSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev] = true;
  }
  const size_t smem = tc_smem_bytes(op.BN, RB, stages, a_slots, a_slot_bytes, scratch);
  const int num_sms = device_sm_count(dev);
  if (num_sms <= 0) return fail(FSE_ECUDA, "cannot query the SM count of device %d", dev);
  const int tile_rows = kTileM * (p.MT > 0 ? p.MT : 1);
  const int total_tiles = p.B * ((p.Trows + tile_rows - 1) / tile_rows) * (p.N / op.BN);
  dim3 grid(total_tiles < num_sms ? total_tiles : num_sms);   // persistent: one CTA per SM
  const CUtensorMap* mA1 = op.mA1 ? op.mA1 : op.mA0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: prologue overlaps the previous kernel's tail
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FSE_CUDA(cudaLaunchKernelEx(&cfg, kern, *op.mA0, *mA1, *op.mW, p, op.BN, stages, a_slots, a_slot_bytes, scratch, epi));
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

// mode: FSE_MODE_*.  TOp = operand element type of this handle.
template <typename TOp, class Epi>
inline int run_conv_gemm_impl(int mode, const ConvGemmParams& p, const GemmOperands& op, const Epi& epi, cudaStream_t st) {
  if constexpr (std::is_same<TOp, __nv_bfloat16>::value) {
    if (mode == FSE_MODE_TC_BF16) {
      if (p.N % op.BN != 0 || op.BN % 16 != 0 || op.BN > 256) return fail(FSE_EINVAL, "conv_gemm: bad BN %d for N %d", op.BN, p.N);
      if (!op.mA0 || !op.mW) return fail(FSE_ESTATE, "conv_gemm: tensor maps missing");
      if (p.KB == 64) {
        if (op.BN % 32 == 0) return launch_tc<TOp, 64, 32, Epi>(p, op, epi, st);
        return launch_tc<TOp, 64, 16, Epi>(p, op, epi, st);
      } else {
        if (op.BN % 32 == 0) return launch_tc<TOp, 32, 32, Epi>(p, op, epi, st);
        return launch_tc<TOp, 32, 16, Epi>(p, op, epi, st);
      }
    }
  } else {
    if (mode == FSE_MODE_TC_TF32) {        // fp32 operands on the tensor cores (kind::tf32): k-blocks of 32 (128 B) or 16 (64 B) channels
      if (p.N % op.BN != 0 || op.BN % 16 != 0 || op.BN > 256) return fail(FSE_EINVAL, "conv_gemm: bad BN %d for N %d", op.BN, p.N);
      if (!op.mA0 || !op.mW) return fail(FSE_ESTATE, "conv_gemm: tensor maps missing");
      if (p.KB != 32 && p.KB != 16) return fail(FSE_EINVAL, "conv_gemm(tf32): k-block must be 32 or 16 channels, got %d", p.KB);
      if (p.KB == 32) {
        if (op.BN % 32 == 0) return launch_tc<TOp, 32, 32, Epi>(p, op, epi, st);
        return launch_tc<TOp, 32, 16, Epi>(p, op, epi, st);
      } else {
        if (op.BN % 32 == 0) return launch_tc<TOp, 16, 32, Epi>(p, op, epi, st);
        return launch_tc<TOp, 16, 16, Epi>(p, op, epi, st);
      }
    }
  }
  const int tiles = (p.Trows + 63) / 64;
  dim3 grid(p.B * tiles, (p.N + 63) / 64);
  conv_gemm_simt_kernel<TOp, Epi><<<grid, 256, 0, st>>>(p, static_cast<const TOp*>(op.A0), static_cast<const TOp*>(op.A1),
                                                       static_cast<const TOp*>(op.W), epi);
  FSE_CUDA(cudaGetLastError());
  return FSE_OK;
}

template <typename TOp, class Epi>
inline int run_conv_gemm(int mode, const ConvGemmParams& p, const GemmOperands& op, const Epi& epi, cudaStream_t st,
                         const LaunchCtx& ctx) {
  if (ctx.launches) ++*ctx.launches;
  if (ctx.prof) ctx.prof->begin(ctx.kind, st);
  const int rc = run_conv_gemm_impl<TOp>(mode, p, op, epi, st);
  if (ctx.prof) ctx.prof->end(st);
  return rc;
}

}  // namespace fse
