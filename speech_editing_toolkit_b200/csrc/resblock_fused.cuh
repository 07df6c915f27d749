// One HiFi-GAN ResBlock1 conv pair as ONE kernel (modules/vocoder/hifigan/hifigan.py:51-58):
//
//     y = x + conv2( lrelu( conv1( lrelu(x) ) ) )          conv1: k taps, dilation d;  conv2: k taps, dilation 1
//
// The two-kernel form (conv_gemm_tc_kernel twice) writes the intermediate lrelu(conv1(.)) to HBM and reads it back: 4 of the
// 16 bytes per element the pair moves with bf16 operands, 8 of 24 with fp32 (tf32) operands - and the narrow late stages
// (C = 32 / 64, 3/4 of all elements) are HBM-bound.  Here the intermediate never leaves the SM:
//
//   job = (item b, Rout = 128 MT - (k-1) consecutive output frames [o0, o0 + Rout))
//   C1  conv1 for the 128 MT intermediate frames [o0 - h2, o0 - h2 + 128 MT), h2 = (k-1)/2: MT sub-tiles of 128 frames against
//       the same weight tiles, every (tap, sub-tile) operand a row-shifted UMMA descriptor of ONE shared-memory copy of the
//       input rows [o0 - h2 - h2 d, ...) per channel block (the shared-A schedule of conv_gemm_tc_kernel; TMA zero fill outside
//       [0, T) is conv1's zero padding)                                                           -> TMEM accumulator 1
//   E1  bias + lrelu, frames outside [0, T) forced to 0 (conv2 pads its INPUT with zeros), operand type (bf16 / tf32-rounded
//       fp32) -> shared memory in the UMMA K-major swizzled layout (the tile `U`)
//   C2  conv2 over U: tap j of sub-tile mt reads U rows [128 mt + j, 128 mt + j + 128) - again row-shifted descriptors - so the
//       last k-1 rows of the job have no valid output (that is the 2 h2 / (128 MT) overlap between jobs)  -> TMEM accumulator 2
//   E2  the pair's residual epilogue (EpiResAdd of hifigan.cu: + bias + fp32 residual, running sum / mean over the parallel
//       blocks, activation for the next layer), transposed through shared memory so that global accesses are coalesced.
//
// Warp roles as in conv_gemm_tc_kernel (warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, NE epilogue warps); C1 of job i+1
// runs on the tensor pipe while the epilogue warps are in E2 of job i.  Within one CTA the four phases of a job are a chain
// (C1 -> E1 -> C2 -> E2), so the kernel is shaped to run TWO CTAs per SM (NE = 4, <= 112 KB shared memory, <= 256 TMEM columns
// each): while one CTA's epilogue warps wait for memory the other one's tensor-core phase runs (measured first with one
// 8-epilogue-warp CTA per SM: no faster than the two-kernel form).  The E2 transposition scratch aliases U, which is dead between
// C2's completion and the next E1 (a named barrier keeps a fast warp's next E1 away from a slow warp's scratch).
// Both operand kinds of the library: bf16 (kind::f16) and fp32 containers (kind::tf32).
#pragma once
#include "conv_gemm.cuh"

namespace fse {

struct PairParams {
  int B, T;                 // items, frames per item (the pair keeps the length)
  int C;                    // channels: C_in = C_out of both convs
  int k, dil;               // taps of both convs, dilation of conv1
  int nkb;                  // k-blocks per tap = ceil(C / KB)
  int MT;                   // 128-frame sub-tiles per job
  int Rout;                 // valid output frames per job = 128 MT - (k - 1)
  int Rbox, nload;          // conv1's input rows arrive as nload TMA boxes of Rbox rows laid end to end
  int a_slots, a_slot_bytes;
  int stages, w_stage_bytes;
  int u_kb_bytes;           // bytes of one k-block of U: (128 MT + k - 1, rounded up to 8) rows x row bytes, 1024-aligned
  float slope1;             // leaky_relu slope between the convs (0.1)
};

__device__ __forceinline__ float lrelu_f(float x, float slope) { return x >= 0.f ? x : x * slope; }

template <typename TOp, int KB, int NE, class Epi>
__global__ void __launch_bounds__(64 + 32 * NE, NE == 4 ? 2 : 1)
resblock_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW1,
                     const __grid_constant__ CUtensorMap mapW2, PairParams p, const float* __restrict__ bias1, Epi epi) {
  constexpr int ES = static_cast<int>(sizeof(TOp));
  constexpr int RB = KB * ES;                       // bytes per k-block row: 128 (128B swizzle) or 64 (64B swizzle)
  constexpr bool kTF32 = std::is_same<TOp, float>::value;
  constexpr int CH = 32;
  static_assert(RB == 128 || RB == 64, "k-block row must be 128 or 64 bytes");
  static_assert(epi_transposed<Epi>::value, "the pair's second epilogue is the transposed one");
  static_assert(NE == 4 || NE == 8, "epilogue warps: one or two per TMEM lane quarter");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sW = sA + p.a_slots * p.a_slot_bytes;
  uint8_t* sU = sW + p.stages * p.w_stage_bytes;
  uint8_t* sScratch = sU;                        // E2's per-warp 4 KB transposition scratch aliases U (NE * 4 KB <= nkb * u_kb_bytes)
  uint64_t* full = reinterpret_cast<uint64_t*>(sU + p.nkb * p.u_kb_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* a_full = empty + p.stages;
  uint64_t* a_empty = a_full + p.a_slots;
  uint64_t* acc_full = a_empty + p.a_slots;     // [2]: accumulator 1 (conv1), accumulator 2 (conv2)
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint64_t* u_full = acc_empty + 2;             // E1 has written U (and released accumulator 1)
  uint64_t* u_empty = u_full + 1;               // conv2's MMAs have read U
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(u_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int C = p.C, MT = p.MT, k = p.k;
  const int h2 = (k - 1) / 2, h1 = h2 * p.dil;
  const int jobs_per_item = (p.T + p.Rout - 1) / p.Rout;
  const int total_jobs = p.B * jobs_per_item;
  uint32_t ncols = 32;
  while (static_cast<int>(ncols) < 2 * MT * C) ncols <<= 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&mapA);
    ptx::prefetch_tensormap(&mapW1);
    ptx::prefetch_tensormap(&mapW2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
      for (int i = 0; i < p.a_slots; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
      for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], NE); }
      ptx::mbar_init(u_full, NE);
      ptx::mbar_init(u_empty, 1);
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
  asm volatile("griddepcontrol.wait;" ::: "memory");              // PDL: everything above overlapped the previous kernel's tail
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int kbg = 0, ga = 0;
      const uint32_t box_bytes = static_cast<uint32_t>(p.Rbox) * RB;
      for (int job = blockIdx.x; job < total_jobs; job += gridDim.x) {
        const int b = job / jobs_per_item, o0 = (job % jobs_per_item) * p.Rout;
        const int f0 = o0 - h2 - h1;                               // first input frame of the job (may be negative: zero fill)
        for (int g = 0; g < p.nkb; ++g, ++ga) {
          const int slot = ga % p.a_slots;
          ptx::mbar_wait(&a_empty[slot], ((ga / p.a_slots) & 1) ^ 1u);
          ptx::mbar_arrive_expect_tx(&a_full[slot], static_cast<uint32_t>(p.nload) * box_bytes);
          for (int i = 0; i < p.nload; ++i)
            ptx::tma_load_3d(sA + slot * p.a_slot_bytes + i * box_bytes, &mapA, &a_full[slot], g * KB, f0 + i * p.Rbox, b);
          for (int j = 0; j < k; ++j, ++kbg) {                    // conv1 weights: k-blocks packed tap-major
            const int s = kbg % p.stages;
            ptx::mbar_wait(&empty[s], ((kbg / p.stages) & 1) ^ 1u);
            ptx::mbar_arrive_expect_tx(&full[s], static_cast<uint32_t>(C * RB));
            ptx::tma_load_2d(sW + s * p.w_stage_bytes, &mapW1, &full[s], (j * p.nkb + g) * KB, 0);
          }
        }
        for (int g = 0; g < p.nkb; ++g)
          for (int j = 0; j < k; ++j, ++kbg) {                    // conv2 weights (its activation operand is U)
            const int s = kbg % p.stages;
            ptx::mbar_wait(&empty[s], ((kbg / p.stages) & 1) ^ 1u);
            ptx::mbar_arrive_expect_tx(&full[s], static_cast<uint32_t>(C * RB));
            ptx::tma_load_2d(sW + s * p.w_stage_bytes, &mapW2, &full[s], (j * p.nkb + g) * KB, 0);
          }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = kTF32 ? ptx::make_idesc_tf32_f32(kTileM, C) : ptx::make_idesc_bf16_f32(kTileM, C);
    auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
      if constexpr (kTF32) ptx::mma_tf32_ss(d, da, db, idesc, acc); else ptx::mma_f16_ss(d, da, db, idesc, acc);
    };
    auto desc = [&](uint32_t addr) { return (RB == 128) ? ptx::make_desc_k_sw128(addr) : ptx::make_desc_k_sw64(addr); };
    const bool el = ptx::elect_one();
    int kbg = 0, ga = 0, it = 0;
    for (int job = blockIdx.x; job < total_jobs; job += gridDim.x, ++it) {
      // ---- C1
      ptx::mbar_wait(&acc_empty[0], (it & 1) ^ 1u);
      ptx::tc_fence_after();
      uint32_t accum = 0;
      for (int g = 0; g < p.nkb; ++g, ++ga) {
        const int slot = ga % p.a_slots;
        ptx::mbar_wait(&a_full[slot], (ga / p.a_slots) & 1);
        for (int j = 0; j < k; ++j, ++kbg) {
          const int s = kbg % p.stages;
          ptx::mbar_wait(&full[s], (kbg / p.stages) & 1);
          ptx::tc_fence_after();
          {
            // descriptors by the whole warp (uniform registers), only the tcgen05 instructions under the one-lane predicate
            const uint64_t db = desc(ptx::smem_u32(sW + s * p.w_stage_bytes));
            const uint32_t a_base = ptx::smem_u32(sA + slot * p.a_slot_bytes) + static_cast<uint32_t>(j * p.dil * RB);
            for (int mt = 0; mt < MT; ++mt) {
              const uint64_t da = desc(a_base + static_cast<uint32_t>(mt * kTileM * RB));
              const uint32_t td = tmem_base + static_cast<uint32_t>(mt * C);
#pragma unroll
              for (int kk = 0; kk < RB / 32; ++kk)
                if (el) mma(td, da + 2 * kk, db + 2 * kk, accum | (kk != 0 ? 1u : 0u));
            }
            if (el) ptx::mma_commit(&empty[s]);
          }
          accum = 1;
          __syncwarp();
        }
        if (el) ptx::mma_commit(&a_empty[slot]);
        __syncwarp();
      }
      if (el) ptx::mma_commit(&acc_full[0]);
      __syncwarp();
      // ---- C2
      ptx::mbar_wait(u_full, it & 1);
      ptx::mbar_wait(&acc_empty[1], (it & 1) ^ 1u);
      ptx::tc_fence_after();
      accum = 0;
      for (int g = 0; g < p.nkb; ++g) {
        for (int j = 0; j < k; ++j, ++kbg) {
          const int s = kbg % p.stages;
          ptx::mbar_wait(&full[s], (kbg / p.stages) & 1);
          ptx::tc_fence_after();
          {
            const uint64_t db = desc(ptx::smem_u32(sW + s * p.w_stage_bytes));
            const uint32_t a_base = ptx::smem_u32(sU + g * p.u_kb_bytes) + static_cast<uint32_t>(j * RB);
            for (int mt = 0; mt < MT; ++mt) {
              const uint64_t da = desc(a_base + static_cast<uint32_t>(mt * kTileM * RB));
              const uint32_t td = tmem_base + static_cast<uint32_t>((MT + mt) * C);
#pragma unroll
              for (int kk = 0; kk < RB / 32; ++kk)
                if (el) mma(td, da + 2 * kk, db + 2 * kk, accum | (kk != 0 ? 1u : 0u));
            }
            if (el) ptx::mma_commit(&empty[s]);
          }
          accum = 1;
          __syncwarp();
        }
      }
      if (el) {
        ptx::mma_commit(u_empty);
        ptx::mma_commit(&acc_full[1]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    constexpr int kSplit = NE / 4;           // warps per lane quarter: they split the column chunks
    const int half = ew >> 2;
    const int cpb = C / CH;                  // chunks per 128-frame sub-tile
    const int nchunks = MT * cpb;
    float* stg = reinterpret_cast<float*>(sScratch) + ew * 1024;
    const int cq = lane & 7, r0 = lane >> 3;
    constexpr int AP = Epi::kAux > 0 ? 4 * Epi::kAux : 1;
    constexpr int AX = Epi::kAux + epi_late<Epi>::value > 0 ? 4 * (Epi::kAux + epi_late<Epi>::value) : 1;
    int it = 0;
    for (int job = blockIdx.x; job < total_jobs; job += gridDim.x, ++it) {
      const int b = job / jobs_per_item, o0 = (job % jobs_per_item) * p.Rout;
      // ---- E1: accumulator 1 -> bias + lrelu -> operand type -> U (lane = frame = TMEM lane, 32 channels per tcgen05.ld)
      ptx::mbar_wait(&acc_full[0], it & 1);
      ptx::tc_fence_after();
      if (it > 0) {
        ptx::mbar_wait(u_empty, (it - 1) & 1);                     // conv2 of the previous job has read U
        asm volatile("bar.sync 1, %0;" ::"n"(NE * 32) : "memory");   // ... and every epilogue warp is done with its E2 scratch (= U)
      }
      for (int c = half; c < nchunks; c += kSplit) {
        const int mt = c / cpb, n0 = (c % cpb) * CH;
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mt * C + n0), r);
        ptx::tmem_wait_ld();
        const int row = mt * kTileM + q * 32 + lane;               // row of U = intermediate frame o0 - h2 + row
        const int frame = o0 - h2 + row;
        const bool ok = frame >= 0 && frame < p.T;                 // conv2 zero-pads its input: frames outside the item are 0
        float v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(bias1 + n0) + i);
          v[4 * i] = ok ? lrelu_f(__uint_as_float(r[4 * i]) + bv.x, p.slope1) : 0.f;
          v[4 * i + 1] = ok ? lrelu_f(__uint_as_float(r[4 * i + 1]) + bv.y, p.slope1) : 0.f;
          v[4 * i + 2] = ok ? lrelu_f(__uint_as_float(r[4 * i + 2]) + bv.z, p.slope1) : 0.f;
          v[4 * i + 3] = ok ? lrelu_f(__uint_as_float(r[4 * i + 3]) + bv.w, p.slope1) : 0.f;
        }
        // K-major rows of RB bytes; 16-byte chunks XOR-swizzled by the address bits [7, 10) (128B rows) / [7, 9) (64B rows)
        uint8_t* ub = sU + (n0 / KB) * p.u_kb_bytes + row * RB;
        const uint32_t sw = (static_cast<uint32_t>(row * RB) >> 7) & (RB / 16 - 1);
        const uint32_t chunk0 = static_cast<uint32_t>((n0 % KB) * ES) >> 4;
        if constexpr (kTF32) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint4*>(ub + (((chunk0 + i) ^ sw) << 4)) =
                make_uint4(__float_as_uint(ptx::round_tf32(v[4 * i])), __float_as_uint(ptx::round_tf32(v[4 * i + 1])),
                           __float_as_uint(ptx::round_tf32(v[4 * i + 2])), __float_as_uint(ptx::round_tf32(v[4 * i + 3])));
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(ub + (((chunk0 + i) ^ sw) << 4)) =
                make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                           pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&acc_empty[0]);
        ptx::mbar_arrive(u_full);
      }
      // ---- E2: accumulator 2 -> the pair's residual epilogue, transposed (8 lanes cover the 32 channels of one frame)
      const int fq = q * 32 + r0;                                  // job-local row of iteration 0 in sub-tile 0
      auto R_OF = [&](int c, int i) { return (c / cpb) * kTileM + fq + 4 * i; };
      auto N_OF = [&](int c) { return (c % cpb) * CH + cq * 4; };
      auto VALID = [&](int c, int i) { const int rl = R_OF(c, i); return rl < p.Rout && o0 + rl < p.T; };
      const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(MT * C);
      float aux[8][AX], aux_next[8][AP];
      if constexpr (Epi::kAux > 0) {
        if (half < nchunks) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (VALID(half, i)) epi.template load_aux<4>(b, o0 + R_OF(half, i), N_OF(half), aux[i]);
        }
      }
      ptx::mbar_wait(&acc_full[1], it & 1);
      ptx::tc_fence_after();
      for (int c = half; c < nchunks; c += kSplit) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(lane_base + c * CH, r);
        ptx::tmem_wait_ld();
        if constexpr (Epi::kAux > 0) {
          if (c + kSplit < nchunks) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (VALID(c + kSplit, i)) epi.template load_aux<4>(b, o0 + R_OF(c + kSplit, i), N_OF(c + kSplit), aux_next[i]);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        const int nn = N_OF(c);
        if constexpr (epi_late<Epi>::value > 0) {     // all of the chunk's late reads before its first store
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (VALID(c, i)) epi.template load_late<4>(b, o0 + R_OF(c, i), nn, aux[i] + 4 * Epi::kAux);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rloc = 4 * i + r0;
          const float4 a = *reinterpret_cast<const float4*>(stg + rloc * 32 + ((cq ^ (rloc & 7)) << 2));
          if (VALID(c, i)) {
            const float v[4] = {a.x, a.y, a.z, a.w};
            epi.template apply<4>(b, o0 + R_OF(c, i), nn, v, aux[i]);
          }
        }
        __syncwarp();                                 // all lanes have read the scratch before the next chunk overwrites it
        if constexpr (Epi::kAux > 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < AP; ++j) aux[i][j] = aux_next[i][j];
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[1]);
    }
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, ncols);
  }
}

}  // namespace fse
