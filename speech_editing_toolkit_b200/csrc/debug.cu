// Test hook: the bare conv_gemm primitive with a store-only epilogue, so tests/ can check the
// tensor-core path (TMA swizzle, UMMA descriptors, TMEM readout) in isolation at arbitrary shapes.
#include "fse_common.cuh"

namespace fse {
struct EpiStore {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  float* out;
  int N, T;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    st_vec<NV>(out + (static_cast<size_t>(b) * T + t) * N + n0, acc);
  }
};
}  // namespace fse

using namespace fse;

extern "C" int fse_debug_conv_gemm(int32_t mode, const void* A0, const void* W, float* out, int32_t B, int32_t T, int32_t C0,
                                   int32_t ntaps, const int32_t* offs, int32_t N, int32_t BN, int32_t KB, void* stream, int64_t* dbg_stamps,
                                   int32_t shared_a /* 0 = one A load per tap, 1 = shared-A schedule (2 = same with descriptor base-offset set: known wrong, kept as a probe) */) {
  if (!A0 || !W || !out || !offs) return fail(FSE_EINVAL, "null argument");
  if (ntaps <= 0 || ntaps > kMaxTaps) return fail(FSE_EINVAL, "ntaps out of range");
  if (KB != 64 && KB != 32) return fail(FSE_EINVAL, "KB must be 32 or 64");
  if (mode != FSE_MODE_TC_BF16 && mode != FSE_MODE_SIMT_BF16) return fail(FSE_EINVAL, "debug gemm takes bf16 operands");
  int o[kMaxTaps];
  for (int i = 0; i < ntaps; ++i) o[i] = offs[i];
  ConvGemmParams p = make_params(B, T, T, C0, ntaps, o, 0, N, KB);
  p.dbg = reinterpret_cast<long long*>(dbg_stamps);
  GemmOperands op; op.A0 = A0; op.W = W; op.BN = BN;
  CUtensorMap mA{}, mW{};
  if (mode == FSE_MODE_TC_BF16) {
    int rows = kTileM;
    // shared_a = schedule + 16 * (MT - 1): MT 128-frame sub-tiles per job; schedule 0 = one A load per tap, 1 = shared-A,
    // 2 = shared-A with the descriptor base-offset set (known wrong, kept as a probe)
    const int mt = 1 + (shared_a >> 4), sched = shared_a & 15;
    if (sched) {
      if (!enable_shared_a(p, mt)) return fail(FSE_EINVAL, "shared-A schedule needs >= 2 taps and a job that fits shared memory");
      p.bo_mode = sched == 2 ? 1 : 0;
      rows = p.Rbox;
    } else {
      p.MT = mt;
    }
    FSE_TRY(make_map_act(&mA, A0, C0, T, B, KB, rows));
    FSE_TRY(make_map_w(&mW, W, p.Kp, N, KB, BN));
    op.mA0 = &mA; op.mW = &mW;
  }
  EpiStore epi{out, N, T};
  return run_conv_gemm<__nv_bfloat16>(mode, p, op, epi, static_cast<cudaStream_t>(stream), LaunchCtx{});
}
