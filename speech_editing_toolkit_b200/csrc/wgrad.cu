// Weight-gradient GEMM of the training step on tcgen05 with MN-major operands: fse_wgrad.
//
//   Out[m, n, j] = sum over b, t of  P[b, t, m] * Q[b, t + offs[j], n]          (rows of Q outside [0, T) read as zero)
//
// which is what torch.autograd computes for the weight of every Conv1d / Linear of DiffNet (diffnet.py:60-132): P = the gradient of the
// layer's output, Q = the layer's input, offs = the conv's tap offsets (the padding rows contribute nothing), K = ALL frames of the batch.
// Both operands lie in memory with the frame index as the slow dimension, i.e. the reduction dimension K is the OUTER one: "MN-major"
// operands in UMMA terms.  No transposed copies are made: TMA loads [KR frames x 128 bytes of channels] boxes (128-byte swizzle), which is
// exactly the canonical MN-major SWIZZLE_128B atom (8 k-rows x 128 B; for tf32 the 32-byte-granular variant, see make_desc_mn_sw128),
// and the shared-memory descriptors / the instruction descriptor (a_major = b_major = 1) tell the tensor core to read them that way:
//     descriptor: start address, LBO = bytes between consecutive 128-byte channel chunks (one TMA box each), SBO = 1024 = bytes between
//     8-frame groups inside a box; one instruction consumes 16 frames (bf16) or 8 frames (tf32) of K.
// The tap shift is a row offset of Q's TMA coordinate (out-of-range frames are zero-filled by TMA, which is the conv's zero padding), and a
// strided view (the per-layer slice of the [B*T, L*2C] gradient buffer) is just a tensor map with a wider row pitch.
//
// Work split: (128-row tile of m) x (<= 256-column tile of n) x tap = a unit; the frames are cut into S contiguous slices so that
// units x S CTAs fill the GPU once.  Slices are combined WITHOUT floating-point atomics: every CTA stores its partial tile and a second
// small kernel, spread over all SMs, sums the S partials of every output element in slice order and writes the result with the caller's
// strides, so the output is bit-reproducible.  (First version: the last CTA of a unit to arrive summed them itself — 1.5 MB through one
// SM at the very end of the launch, a serial chain of L2 round trips: 313 us per GEMM instead of ~15.)
// Several GEMMs over the same (B, T) grid can share one launch (fse_wgrad_group: the four weight gradients of a residual layer): their
// tiles are numbered through, one S is chosen for all, and one reduction kernel follows - per GEMM launched alone the fixed costs (prologue,
// epilogue, 19 MB of partial tiles whatever the shape) were as large as the MMA time.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (one elected lane), warps 2-5 epilogue (thread = TMEM lane = output row m).
#include <cuda.h>
#include <cuda_runtime.h>

#include "fse_common.cuh"
#include "ptx_sm100.cuh"

namespace fse {
namespace {

constexpr int kWgThreads = 192;
constexpr int kWgStages = 4;
constexpr int kWgTileM = 128, kWgMaxN = 256;
constexpr int kWgPBytes = 16 * 1024;        // P stage: 128 channels x KR frames
constexpr int kWgQBytes = 32 * 1024;        // Q stage: up to 256 channels x KR frames
constexpr int kWgStageBytes = kWgPBytes + kWgQBytes;
constexpr size_t kWgSmemBytes = 1024 + static_cast<size_t>(kWgStages) * kWgStageBytes + 256;
constexpr int kWgMaxTaps = 16;
constexpr int kWgMaxProblems = 4;

struct WgProblem {
  int M, N, BN, ntaps, tiles_n, unit0, units;
  int offs[kWgMaxTaps];
  float* out;
  long long ld_m, ld_n, ld_j;
};
struct WgradParams {
  int B, T, S, total_chunks, cpi, nprob, total_units, bn_max;       // cpi: K chunks per utterance
  WgProblem pr[kWgMaxProblems];
  float* partials;                                                   // [unit][S][128][bn_max]
};
struct WgMaps { CUtensorMap p[kWgMaxProblems], q[kWgMaxProblems]; };

// MN-major operand tile written by TMA: chunks of 128 B of channels x KR frames, chunk c at `addr + c * lbo`.
//   16-bit operands: CU_TENSOR_MAP_SWIZZLE_128B, canonical atom 8 frames x 128 B (layout code 2, SBO = 1024 between 8-frame groups);
//   32-bit operands (tf32): the ONLY MN-major layout the tensor core reads is SWIZZLE_128B_BASE32B (layout code 1): 32-byte chunks
//   XOR-ed with (frame mod 4), atom 4 frames x 128 B, SBO = 512 between 4-frame groups - what TMA writes with
//   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  (First contact with the 16-byte swizzle for tf32: every product came out zero.)
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, bool base32) {
  const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((base32 ? 512u : 1024u) >> 4) | (1u << 14) | ((base32 ? 1u : 2u) << 29);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

__device__ __forceinline__ int wg_problem_of(const WgradParams& p, int unit) {
  int g = 0;
#pragma unroll
  for (int i = 1; i < kWgMaxProblems; ++i)
    if (i < p.nprob && unit >= p.pr[i].unit0) g = i;
  return g;
}

template <typename TOp>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgMaps maps, const __grid_constant__ WgradParams p) {
  constexpr bool kTF32 = std::is_same<TOp, float>::value;
  constexpr int ES = static_cast<int>(sizeof(TOp));
  constexpr int CC = 128 / ES;                        // channels per 128-byte chunk: 64 (bf16) or 32 (fp32)
  constexpr int KR = kWgPBytes / (kWgTileM * ES);      // frames per stage: 64 (bf16) or 32 (fp32)
  constexpr int KI = 32 / ES;                         // frames per MMA instruction: 16 or 8
  constexpr uint32_t kLbo = KR * 128;                  // bytes between channel chunks (one TMA box each)
  extern __shared__ uint8_t wg_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wg_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
  uint64_t* empty = full + kWgStages;
  uint64_t* acc_full = empty + kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int unit = blockIdx.x / p.S, slice = blockIdx.x % p.S;
  const int g = wg_problem_of(p, unit);
  const WgProblem& pr = p.pr[g];
  const int lu = unit - pr.unit0;
  const int tap = lu % pr.ntaps, tn = (lu / pr.ntaps) % pr.tiles_n, tm = lu / (pr.ntaps * pr.tiles_n);
  const int m0 = tm * kWgTileM, n0 = tn * pr.BN;
  const int c_begin = static_cast<int>(static_cast<long long>(p.total_chunks) * slice / p.S);
  const int c_end = static_cast<int>(static_cast<long long>(p.total_chunks) * (slice + 1) / p.S);
  const int nchunks = c_end - c_begin;
  uint32_t ncols = 32;
  while (static_cast<int>(ncols) < pr.BN) ncols <<= 1;
  const CUtensorMap* mapP = &maps.p[g];
  const CUtensorMap* mapQ = &maps.q[g];

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(mapP);
    ptx::prefetch_tensormap(mapQ);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kWgStages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
      ptx::mbar_init(acc_full, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, ncols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nq = (pr.BN + CC - 1) / CC;                // channel chunks of the Q tile

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = static_cast<uint32_t>((kWgTileM / CC + nq) * KR * 128);
      const int off = pr.offs[tap];
      for (int i = 0; i < nchunks; ++i) {
        const int s = i % kWgStages, u = i / kWgStages;
        const int c = c_begin + i, b = c / p.cpi, t0 = (c % p.cpi) * KR;
        ptx::mbar_wait(&empty[s], (u & 1) ^ 1u);
        ptx::mbar_arrive_expect_tx(&full[s], tx);
        uint8_t* sp = smem + s * kWgStageBytes;
        uint8_t* sq = sp + kWgPBytes;
#pragma unroll 1
        for (int k = 0; k < kWgTileM / CC; ++k) ptx::tma_load_3d(sp + k * kLbo, mapP, &full[s], m0 + k * CC, t0, b);
#pragma unroll 1
        for (int k = 0; k < nq; ++k) ptx::tma_load_3d(sq + k * kLbo, mapQ, &full[s], n0 + k * CC, t0 + off, b);
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: D = f32, A = B = bf16 / tf32, BOTH MN-major (bits 15, 16), N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (kTF32 ? ptx::make_idesc_tf32_f32(kWgTileM, pr.BN) : ptx::make_idesc_bf16_f32(kWgTileM, pr.BN)) | (1u << 15) | (1u << 16);
    const bool el = ptx::elect_one();
    for (int i = 0; i < nchunks; ++i) {
      const int s = i % kWgStages, u = i / kWgStages;
      ptx::mbar_wait(&full[s], u & 1);
      ptx::tc_fence_after();
      const uint32_t pa = ptx::smem_u32(smem + s * kWgStageBytes), qa = pa + kWgPBytes;
#pragma unroll
      for (int k = 0; k < KR / KI; ++k) {
        const uint64_t da = make_desc_mn_sw128(pa + k * KI * 128, kLbo, kTF32), db = make_desc_mn_sw128(qa + k * KI * 128, kLbo, kTF32);
        if (el) {
          if constexpr (kTF32) ptx::mma_tf32_ss(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
          else ptx::mma_f16_ss(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
        }
      }
      if (el) ptx::mma_commit(&empty[s]);
      __syncwarp();
    }
    if (el) ptx::mma_commit(acc_full);
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- epilogue: thread = output row m0 + q*32 + lane
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* part = p.partials + ((static_cast<size_t>(unit) * p.S + slice) * kWgTileM + row) * p.bn_max;
    float* orow = pr.out + static_cast<long long>(m0 + row) * pr.ld_m + static_cast<long long>(tap) * pr.ld_j;
    const bool row_ok = m0 + row < pr.M;
    if (nchunks > 0) {
      ptx::mbar_wait(acc_full, 0);
      ptx::tc_fence_after();
    }
    for (int c0 = 0; c0 < pr.BN; c0 += 32) {
      uint32_t r[32];
      if (nchunks > 0) {
        ptx::tmem_ld_32x32b_x32(taddr + c0, r);
        ptx::tmem_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      if (p.S == 1) {
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (n0 + c0 + i < pr.N) orow[static_cast<long long>(n0 + c0 + i) * pr.ld_n] = __uint_as_float(r[i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<uint4*>(part + c0 + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, ncols);
  }
}

// Out[m, n, j] = sum over the S slices, in slice order, of the partial tiles; coalesced float4 reads, strided scalar writes
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ WgradParams p) {
  const int unit = blockIdx.x;
  const int g = wg_problem_of(p, unit);
  const WgProblem& pr = p.pr[g];
  const int lu = unit - pr.unit0;
  const int tap = lu % pr.ntaps, tn = (lu / pr.ntaps) % pr.tiles_n, tm = lu / (pr.ntaps * pr.tiles_n);
  const int vpr = pr.BN / 4;                           // float4 per row of this problem's tile
  const int vec_per_tile = kWgTileM * vpr;
  const size_t tile = static_cast<size_t>(kWgTileM) * p.bn_max;
  const float* base = p.partials + static_cast<size_t>(unit) * p.S * tile;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < vec_per_tile; i += gridDim.y * blockDim.x) {
    const int row = i / vpr, col = (i % vpr) * 4;
    const float* src = base + static_cast<size_t>(row) * p.bn_max + col;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    for (; s + 4 <= p.S; s += 4) {
      const float4 v0 = __ldcs(reinterpret_cast<const float4*>(src + (s + 0) * tile)), v1 = __ldcs(reinterpret_cast<const float4*>(src + (s + 1) * tile));
      const float4 v2 = __ldcs(reinterpret_cast<const float4*>(src + (s + 2) * tile)), v3 = __ldcs(reinterpret_cast<const float4*>(src + (s + 3) * tile));
      acc.x = (((acc.x + v0.x) + v1.x) + v2.x) + v3.x; acc.y = (((acc.y + v0.y) + v1.y) + v2.y) + v3.y;
      acc.z = (((acc.z + v0.z) + v1.z) + v2.z) + v3.z; acc.w = (((acc.w + v0.w) + v1.w) + v2.w) + v3.w;
    }
    for (; s < p.S; ++s) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(src + s * tile));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const int m = tm * kWgTileM + row, n = tn * pr.BN + col;
    if (m >= pr.M) continue;
    float* o = pr.out + static_cast<long long>(m) * pr.ld_m + static_cast<long long>(tap) * pr.ld_j;
    const float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (n + k < pr.N) o[static_cast<long long>(n + k) * pr.ld_n] = v[k];
  }
}

// dims (cols, T, B) with row pitch ld (elements); box (128 bytes of channels, rows, 1); out-of-range rows / columns read 0
int make_map_rows(CUtensorMap* m, const void* ptr, int cols, long long ld, int T, int B, int box_rows, int es) {
  PFN_tmapEncodeTiled enc;
  FSE_TRY(get_encode_fn(&enc));
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * es, static_cast<cuuint64_t>(T) * ld * es};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(128 / es), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, es == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(FSE_ECUDA, "cuTensorMapEncodeTiled(wgrad operand cols=%d ld=%lld T=%d B=%d) failed: %d", cols, ld, T, B, (int)r);
  return FSE_OK;
}

// tiles of every problem, the common number of frame slices and the scratch size
int plan_wgrad(int es, int B, int T, const fse_wgrad_problem* pb, int n, int num_sms, WgradParams* p, size_t* ws_bytes) {
  const int CC = 128 / es, KR = kWgPBytes / (kWgTileM * es);
  if (n < 1 || n > kWgMaxProblems) return fail(FSE_EINVAL, "1 <= number of GEMMs per launch <= %d", kWgMaxProblems);
  p->B = B; p->T = T; p->nprob = n; p->bn_max = 0;
  int units = 0;
  for (int g = 0; g < n; ++g) {
    if (pb[g].M <= 0 || pb[g].N <= 0 || pb[g].ntaps <= 0 || pb[g].ntaps > kWgMaxTaps) return fail(FSE_EINVAL, "M, N must be positive and 1 <= ntaps <= %d", kWgMaxTaps);
    WgProblem& pr = p->pr[g];
    int bn = (pb[g].N + CC - 1) / CC * CC;
    pr.tiles_n = (bn + kWgMaxN - 1) / kWgMaxN;
    pr.BN = ((pb[g].N + pr.tiles_n - 1) / pr.tiles_n + CC - 1) / CC * CC;       // equal tiles, whole channel chunks (a multiple of 16 columns)
    pr.M = pb[g].M; pr.N = pb[g].N; pr.ntaps = pb[g].ntaps;
    pr.unit0 = units;
    pr.units = (pb[g].M + kWgTileM - 1) / kWgTileM * pr.tiles_n * pb[g].ntaps;
    units += pr.units;
    for (int j = 0; j < kWgMaxTaps; ++j) pr.offs[j] = (j < pb[g].ntaps && pb[g].offs) ? pb[g].offs[j] : 0;
    pr.out = pb[g].out; pr.ld_m = pb[g].ld_m; pr.ld_n = pb[g].ld_n; pr.ld_j = pb[g].ld_j;
    if (pr.BN > p->bn_max) p->bn_max = pr.BN;
  }
  p->total_units = units;
  p->cpi = (T + KR - 1) / KR;
  p->total_chunks = B * p->cpi;
  int S = num_sms / units;
  if (S < 1) S = 1;
  if (S > p->total_chunks) S = p->total_chunks;
  p->S = S;
  *ws_bytes = S > 1 ? static_cast<size_t>(units) * S * kWgTileM * p->bn_max * sizeof(float) : 16;
  return FSE_OK;
}

int check_device_sms(int* num_sms, int* dev_out) {
  int dev = 0;
  FSE_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(FSE_EINVAL, "device ordinal %d out of range", dev);
  int major = 0;
  FSE_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(FSE_ECUDA, "this library is built for sm_100a only (no fallback)");
  *num_sms = device_sm_count(dev);
  if (*num_sms <= 0) return fail(FSE_ECUDA, "cannot query the SM count of device %d", dev);
  *dev_out = dev;
  return FSE_OK;
}

template <typename TOp>
int launch_wgrad(const fse_wgrad_problem* pb, int n, int B, int T, void* ws, long long ws_bytes, cudaStream_t st) {
  constexpr int ES = static_cast<int>(sizeof(TOp));
  constexpr int KR = kWgPBytes / (kWgTileM * ES);
  int num_sms = 0, dev = 0;
  FSE_TRY(check_device_sms(&num_sms, &dev));
  WgradParams p{};
  size_t need = 0;
  FSE_TRY(plan_wgrad(ES, B, T, pb, n, num_sms, &p, &need));
  if (ws_bytes < static_cast<long long>(need)) return fail(FSE_EINVAL, "wgrad workspace too small: %lld < %lld bytes", ws_bytes, static_cast<long long>(need));
  WgMaps maps{};
  for (int g = 0; g < n; ++g) {
    if (!pb[g].P || !pb[g].Q || !pb[g].out) return fail(FSE_EINVAL, "null argument");
    if (pb[g].ldp < pb[g].M || pb[g].ldq < pb[g].N) return fail(FSE_EINVAL, "row pitch smaller than the number of columns");
    if ((pb[g].ldp * ES) % 16 || (pb[g].ldq * ES) % 16 || reinterpret_cast<uintptr_t>(pb[g].P) % 16 || reinterpret_cast<uintptr_t>(pb[g].Q) % 16)
      return fail(FSE_EINVAL, "operands and their row pitch (bytes) must be 16-byte aligned");
    FSE_TRY(make_map_rows(&maps.p[g], pb[g].P, pb[g].M, pb[g].ldp, T, B, KR, ES));
    FSE_TRY(make_map_rows(&maps.q[g], pb[g].Q, pb[g].N, pb[g].ldq, T, B, KR, ES));
  }
  p.partials = static_cast<float*>(ws);
  static bool attr_set[kMaxDevices] = {};
  auto kern = wgrad_tc_kernel<TOp>;
  if (!attr_set[dev]) {
    FSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWgSmemBytes)));
    attr_set[dev] = true;
  }
  kern<<<p.total_units * p.S, kWgThreads, kWgSmemBytes, st>>>(maps, p);
  FSE_CUDA(cudaGetLastError());
  if (p.S > 1) {
    int by = (num_sms * 2 + p.total_units - 1) / p.total_units;
    const int max_by = (kWgTileM * p.bn_max / 4 + 255) / 256;
    if (by > max_by) by = max_by;
    wgrad_reduce_kernel<<<dim3(p.total_units, by), 256, 0, st>>>(p);
    FSE_CUDA(cudaGetLastError());
  }
  return FSE_OK;
}

}  // namespace
}  // namespace fse

using namespace fse;

extern "C" {

int64_t fse_wgrad_group_workspace_bytes(int32_t mode, const fse_wgrad_problem* problems, int32_t n, int32_t B, int32_t T) {
  if (!problems || B <= 0 || T <= 0) return 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  const int num_sms = device_sm_count(dev);
  if (num_sms <= 0) return 0;
  WgradParams p{};
  size_t need = 0;
  if (plan_wgrad(mode == FSE_MODE_TC_BF16 ? 2 : 4, B, T, problems, n, num_sms, &p, &need) != FSE_OK) return 0;
  return static_cast<int64_t>(need);
}

int fse_wgrad_group(int32_t mode, const fse_wgrad_problem* problems, int32_t n, int32_t B, int32_t T, void* workspace, int64_t workspace_bytes,
                    void* stream) {
  if (!problems || !workspace) return fail(FSE_EINVAL, "null argument");
  if (B <= 0 || T <= 0) return fail(FSE_EINVAL, "B and T must be positive");
  if (mode != FSE_MODE_TC_BF16 && mode != FSE_MODE_TC_TF32) return fail(FSE_EINVAL, "fse_wgrad runs in the tensor-core modes (FSE_MODE_TC_BF16 / FSE_MODE_TC_TF32)");
  if (reinterpret_cast<uintptr_t>(workspace) % 16) return fail(FSE_EINVAL, "the workspace must be 16-byte aligned");
  if (n < 1 || n > kWgMaxProblems) return fail(FSE_EINVAL, "1 <= number of GEMMs per launch <= %d", kWgMaxProblems);
  const int es = mode == FSE_MODE_TC_BF16 ? 2 : 4;
  for (int g = 0; g < n; ++g) {                          // every argument error is reported before the device is touched
    const fse_wgrad_problem& q = problems[g];
    if (!q.P || !q.Q || !q.out) return fail(FSE_EINVAL, "null argument (GEMM %d)", g);
    if (q.M <= 0 || q.N <= 0 || q.ntaps <= 0 || q.ntaps > kWgMaxTaps) return fail(FSE_EINVAL, "M, N must be positive and 1 <= ntaps <= %d (GEMM %d)", kWgMaxTaps, g);
    if (q.ldp < q.M || q.ldq < q.N) return fail(FSE_EINVAL, "row pitch smaller than the number of columns (GEMM %d)", g);
    if ((q.ldp * es) % 16 || (q.ldq * es) % 16 || reinterpret_cast<uintptr_t>(q.P) % 16 || reinterpret_cast<uintptr_t>(q.Q) % 16)
      return fail(FSE_EINVAL, "operands and their row pitch (bytes) must be 16-byte aligned (GEMM %d)", g);
  }
  auto st = static_cast<cudaStream_t>(stream);
  if (mode == FSE_MODE_TC_BF16) return launch_wgrad<__nv_bfloat16>(problems, n, B, T, workspace, workspace_bytes, st);
  return launch_wgrad<float>(problems, n, B, T, workspace, workspace_bytes, st);
}

int64_t fse_wgrad_workspace_bytes(int32_t mode, int32_t B, int32_t T, int32_t M, int32_t N, int32_t ntaps) {
  fse_wgrad_problem pb{};
  pb.M = M; pb.N = N; pb.ntaps = ntaps;
  return fse_wgrad_group_workspace_bytes(mode, &pb, 1, B, T);
}

int fse_wgrad(int32_t mode, const void* P, int64_t ldp, const void* Q, int64_t ldq, int32_t B, int32_t T, int32_t M, int32_t N, const int32_t* offs,
              int32_t ntaps, float* out, int64_t ld_m, int64_t ld_n, int64_t ld_j, void* workspace, int64_t workspace_bytes, void* stream) {
  fse_wgrad_problem pb{P, ldp, Q, ldq, M, N, offs, ntaps, out, ld_m, ld_n, ld_j};
  return fse_wgrad_group(mode, &pb, 1, B, T, workspace, workspace_bytes, stream);
}

}  // extern "C"
