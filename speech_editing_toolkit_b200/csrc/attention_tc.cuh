// Tensor-core attention for sm_100a: O = softmax(Q K^T) V per (item, head), head_dim 96, bf16 operands, fp32 softmax.
// (MultiheadAttention.forward of the reference, modules/speech_editing/commons/transformer.py:365-407; q carries head_dim^-0.5.)
//
// One CTA = 128 queries of one (item, head); keys / values stream through shared memory in tiles of 128 keys.
//   warp 0   TMA producer: Q once; per tile K [128 keys x 96] and V^T [96 x 128 keys] (two 64-column boxes each, 128B swizzle,
//            out-of-range rows / keys zero-filled), double buffered
//   warp 1   tcgen05.mma issuer + TMEM owner:  S_j = Q K_j^T (M 128 x N 128, K = 96 as 4 + 2 UMMA_K steps) into one of two
//            TMEM buffers, then O_j = P_j V_j (M 128 x N 96, K = 128 keys) into a third; S_{j+1} is issued before P_j is awaited
//   warps 2-5 softmax: thread = query row = TMEM lane.  Two passes over the S row in TMEM (max, then exp / sum), P written to
//            shared memory as bf16 in the UMMA K-major 128B-swizzled layout (the A operand of the P.V MMA), running
//            (max, sum) per row, O accumulated in registers: o = o * alpha + O_j read back from TMEM (no TMEM read-modify-write)
// Masking as in the reference: padded keys get -1e8 (exp underflows to exactly 0), keys past Tk are excluded.
// V^T ([B * heads * 96, Tkp] bf16, keys contiguous) is written by the q/k/v projection's epilogue (EpiScaleCols::vt), so both
// MMA operands are K-major and the verified descriptor forms of conv_gemm.cuh apply unchanged.
#pragma once
#include <cuda_bf16.h>

#include "fse_common.cuh"

namespace fse {
namespace {

constexpr int kAttD = 96;            // head_dim (hidden 192 / 2 heads)
constexpr int kAtcM = 128;           // queries per CTA (TMEM lanes)
constexpr int kAtcN = 128;           // keys per tile
constexpr int kAtcKB = 16384;        // one 64-column k-block of a 128-row operand tile: 128 rows x 128 B
constexpr int kAtcVB = 12288;        // one 64-key k-block of the V^T tile: 96 rows x 128 B
constexpr int kAtcThreads = 192;
constexpr int kAtcSmemBytes = 1024 /*alignment slack*/ + 2 * kAtcKB /*Q*/ + 2 * 2 * kAtcKB /*K x2*/ + 2 * 2 * kAtcVB /*V^T x2*/ +
                              2 * kAtcKB /*P*/ + 2 * kAtcN * 4 /*key flags x2*/ + 256 /*barriers + TMEM slot*/;

struct AttnTcParams {
  int Tq, Tk, heads;
  int qoff, koff;              // first column of head 0's queries / keys inside their rows
  int ldo;                     // row stride of O in elements
  const float* key_keep;       // [B, Tk] 1 = attend, 0 = padded key; or null
  __nv_bfloat16* O;            // [B, Tq, ldo], head h at columns h*96
  float* probs = nullptr;      // tc2 kernel, Tk <= 128 only: [B, Tq, Tk] (zeroed by the caller) += softmax row / heads
};

__global__ void __launch_bounds__(kAtcThreads, 1) camp_attention_tc_kernel(const __grid_constant__ CUtensorMap mapQ,
                                                                           const __grid_constant__ CUtensorMap mapK,
                                                                           const __grid_constant__ CUtensorMap mapVt, AttnTcParams p) {
  extern __shared__ uint8_t atc_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(atc_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = base;
  uint8_t* sK = sQ + 2 * kAtcKB;
  uint8_t* sV = sK + 4 * kAtcKB;
  uint8_t* sP = sV + 4 * kAtcVB;
  float* sKeep = reinterpret_cast<float*>(sP + 2 * kAtcKB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKeep + 2 * kAtcN);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;     // [2]
  uint64_t* s_empty = bars + 7;    // [2]
  uint64_t* p_full = bars + 9;
  uint64_t* p_empty = bars + 10;
  uint64_t* o_full = bars + 11;
  uint64_t* o_empty = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAtcM, h = blockIdx.y, b = blockIdx.z;
  const int ntiles = (p.Tk + kAtcN - 1) / kAtcN;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&mapQ);
    ptx::prefetch_tensormap(&mapK);
    ptx::prefetch_tensormap(&mapVt);
  }
  if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&kv_full[i], 1);
        ptx::mbar_init(&kv_empty[i], 1);
        ptx::mbar_init(&s_full[i], 1);
        ptx::mbar_init(&s_empty[i], 4);      // one arrival per softmax warp
      }
      ptx::mbar_init(p_full, 4);
      ptx::mbar_init(p_empty, 1);
      ptx::mbar_init(o_full, 1);
      ptx::mbar_init(o_empty, 4);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);         // S buffers at columns 0 and 128, O_j at 256
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, 2 * kAtcKB);
      ptx::tma_load_3d(sQ, &mapQ, q_full, p.qoff + h * kAttD, q0, b);
      ptx::tma_load_3d(sQ + kAtcKB, &mapQ, q_full, p.qoff + h * kAttD + 64, q0, b);     // columns 64..95 are read, the rest ignored
      for (int j = 0; j < ntiles; ++j) {
        const int buf = j & 1, u = j >> 1, k0 = j * kAtcN;
        ptx::mbar_wait(&kv_empty[buf], (u & 1) ^ 1u);
        ptx::mbar_arrive_expect_tx(&kv_full[buf], 2 * kAtcKB + 2 * kAtcVB);
        uint8_t* dk = sK + buf * 2 * kAtcKB;
        uint8_t* dv = sV + buf * 2 * kAtcVB;
        ptx::tma_load_3d(dk, &mapK, &kv_full[buf], p.koff + h * kAttD, k0, b);
        ptx::tma_load_3d(dk + kAtcKB, &mapK, &kv_full[buf], p.koff + h * kAttD + 64, k0, b);
        ptx::tma_load_2d(dv, &mapVt, &kv_full[buf], k0, (b * p.heads + h) * kAttD);
        ptx::tma_load_2d(dv + kAtcVB, &mapVt, &kv_full[buf], k0 + 64, (b * p.heads + h) * kAttD);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idS = ptx::make_idesc_bf16_f32(kAtcM, kAtcN), idO = ptx::make_idesc_bf16_f32(kAtcM, kAttD);
    const bool el = ptx::elect_one();
    ptx::mbar_wait(q_full, 0);
    ptx::tc_fence_after();
    auto issue_S = [&](int j) {
      const int buf = j & 1, u = j >> 1;
      ptx::mbar_wait(&kv_full[buf], u & 1);
      ptx::mbar_wait(&s_empty[buf], (u & 1) ^ 1u);
      ptx::tc_fence_after();
      const uint32_t td = tmem_base + static_cast<uint32_t>(buf * kAtcN);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t da = ptx::make_desc_k_sw128(ptx::smem_u32(sQ + kb * kAtcKB));
        const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(sK + buf * 2 * kAtcKB + kb * kAtcKB));
        const int nsteps = kb == 0 ? 4 : 2;                       // head_dim 96 = 64 + 32
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < nsteps && el) ptx::mma_f16_ss(td, da + 2 * k, db + 2 * k, idS, (kb | k) != 0 ? 1u : 0u);
      }
      if (el) ptx::mma_commit(&s_full[buf]);
      __syncwarp();
    };
    issue_S(0);
    for (int j = 0; j < ntiles; ++j) {
      if (j + 1 < ntiles) issue_S(j + 1);
      const int buf = j & 1;
      ptx::mbar_wait(p_full, j & 1);
      ptx::mbar_wait(o_empty, (j & 1) ^ 1u);
      ptx::tc_fence_after();
      const uint32_t td = tmem_base + 256u;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t da = ptx::make_desc_k_sw128(ptx::smem_u32(sP + kb * kAtcKB));
        const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(sV + buf * 2 * kAtcVB + kb * kAtcVB));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (el) ptx::mma_f16_ss(td, da + 2 * k, db + 2 * k, idO, (kb | k) != 0 ? 1u : 0u);
      }
      if (el) {
        ptx::mma_commit(o_full);
        ptx::mma_commit(p_empty);
        ptx::mma_commit(&kv_empty[buf]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ softmax + output: thread = query row
    const int qw = warp & 3;                       // TMEM lane quarter this warp may read
    const int r = qw * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(qw * 32) << 16);
    float o[kAttD];
#pragma unroll
    for (int i = 0; i < kAttD; ++i) o[i] = 0.f;
    float mrow = -INFINITY, lrow = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const int buf = j & 1, u = j >> 1, k0 = j * kAtcN;
      {
        const int k = k0 + r;
        sKeep[buf * kAtcN + r] = k >= p.Tk ? -1.f : (p.key_keep ? p.key_keep[static_cast<size_t>(b) * p.Tk + k] : 1.f);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");            // the four softmax warps
      const float* flags = sKeep + buf * kAtcN;
      ptx::mbar_wait(&s_full[buf], u & 1);
      ptx::tc_fence_after();
      const uint32_t s_addr = lane_addr + static_cast<uint32_t>(buf * kAtcN);
      uint32_t rr[32];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ptx::tmem_ld_32x32b_x32(s_addr + c * 32, rr);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float f = flags[c * 32 + i];
          const float sv = f < 0.f ? -INFINITY : (f == 0.f ? -1e8f : __uint_as_float(rr[i]));
          mx = fmaxf(mx, sv);
        }
      }
      const float mnew = fmaxf(mrow, mx);          // finite: key k0 of every tile is < Tk
      const float alpha = expf(mrow - mnew);       // exp(-inf) = 0 on the first tile
      ptx::mbar_wait(p_empty, (j & 1) ^ 1u);       // the previous P.V MMA has read sP
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ptx::tmem_ld_32x32b_x32(s_addr + c * 32, rr);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float f = flags[c * 32 + g * 8 + i];
            const float sv = f == 0.f ? -1e8f : __uint_as_float(rr[g * 8 + i]);
            pv[i] = f < 0.f ? 0.f : expf(sv - mnew);
            sum += pv[i];
          }
          const int kc = c * 32 + g * 8;             // first key column of this 16-byte chunk
          uint8_t* dst = sP + (kc >> 6) * kAtcKB + r * 128 + ((((kc & 63) >> 3) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]),
                                                      pack_bf16x2(pv[6], pv[7]));
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&s_empty[buf]);
        ptx::mbar_arrive(p_full);
      }
      lrow = lrow * alpha + sum;
      mrow = mnew;
      ptx::mbar_wait(o_full, j & 1);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        ptx::tmem_ld_32x32b_x32(lane_addr + 256u + c * 32, rr);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha, __uint_as_float(rr[i]));
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(o_empty);
    }
    const int q = q0 + r;
    if (q < p.Tq) {
      const float inv = 1.0f / lrow;
      __nv_bfloat16* dst = p.O + (static_cast<size_t>(b) * p.Tq + q) * p.ldo + h * kAttD;
#pragma unroll
      for (int c = 0; c < kAttD / 8; ++c)
        reinterpret_cast<uint4*>(dst)[c] = make_uint4(pack_bf16x2(o[8 * c] * inv, o[8 * c + 1] * inv), pack_bf16x2(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                                                     pack_bf16x2(o[8 * c + 4] * inv, o[8 * c + 5] * inv), pack_bf16x2(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Second schedule: TWO query tiles (256 queries) per CTA with two softmax warp groups that share the K / V^T tiles.
//   warp 0      TMA producer: Q (two tiles) once; K double-buffered, V^T single-buffered (it is only needed after a softmax pass)
//   warp 1      MMA issuer: per key tile  PV_A(j), S_A(j+1), PV_B(j), S_B(j+1)  — a group's next S tile is computed while that
//               group accumulates O, and one group's softmax overlaps the other's MMAs
//   warps 2-5   softmax group A (query tile 0), warps 6-9 group B (query tile 1): thread = query row = TMEM lane
// TMEM (512 columns): S_A 0, S_B 128, O_A 256, O_B 384.  S is single-buffered per group: P_g(j) being published implies the
// group has finished reading S_g(j), so the issuer may overwrite it with S_g(j+1) right after PV_g(j).
// Softmax work per element is cut to ~6 instructions: tiles without padded / out-of-range keys (every tile of the decoder's
// self-attention) skip the key-flag path entirely, and exp is ex2.approx on log2e-prescaled scores (P is rounded to bf16 anyway).
constexpr int kAtc2Threads = 320;
constexpr int kAtc2SmemBytes = 1024 /*alignment slack*/ + 2 * 2 * kAtcKB /*Q x2 tiles*/ + 2 * 2 * kAtcKB /*K x2*/ + 2 * kAtcVB /*V^T*/ +
                               2 * 2 * kAtcKB /*P x2 groups*/ + 2 * 2 * kAtcN * 4 /*key flags: 2 groups x 2 buffers*/ + 256;

__global__ void __launch_bounds__(kAtc2Threads, 1) camp_attention_tc2_kernel(const __grid_constant__ CUtensorMap mapQ,
                                                                             const __grid_constant__ CUtensorMap mapK,
                                                                             const __grid_constant__ CUtensorMap mapVt, AttnTcParams p) {
  extern __shared__ uint8_t atc_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(atc_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = base;                          // [2 tiles][2 k-blocks]
  uint8_t* sK = sQ + 4 * kAtcKB;               // [2 buffers][2 k-blocks]
  uint8_t* sV = sK + 4 * kAtcKB;               // [2 k-blocks]
  uint8_t* sP = sV + 2 * kAtcVB;               // [2 groups][2 k-blocks]
  float* sKeep = reinterpret_cast<float*>(sP + 4 * kAtcKB);     // [2 groups][2 buffers][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKeep + 4 * kAtcN);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* k_empty = bars + 3;    // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;     // [2 groups]
  uint64_t* p_full = bars + 9;     // [2]
  uint64_t* p_empty = bars + 11;   // [2]
  uint64_t* o_full = bars + 13;    // [2]
  uint64_t* o_empty = bars + 15;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * kAtcM, h = blockIdx.y, b = blockIdx.z;
  const int ntiles = (p.Tk + kAtcN - 1) / kAtcN;
  const bool tileB = q0 + kAtcM < p.Tq;        // the second query tile exists (uniform per CTA)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&mapQ);
    ptx::prefetch_tensormap(&mapK);
    ptx::prefetch_tensormap(&mapVt);
  }
  if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_init(q_full, 1);
      ptx::mbar_init(v_full, 1);
      ptx::mbar_init(v_empty, 1);
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&k_full[i], 1);
        ptx::mbar_init(&k_empty[i], 1);
        ptx::mbar_init(&s_full[i], 1);
        ptx::mbar_init(&p_full[i], 4);       // one arrival per softmax warp of the group
        ptx::mbar_init(&p_empty[i], 1);
        ptx::mbar_init(&o_full[i], 1);
        ptx::mbar_init(&o_empty[i], 4);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, 4 * kAtcKB);
      for (int g = 0; g < 2; ++g) {            // rows past Tq are zero-filled, so the second tile is loaded unconditionally
        ptx::tma_load_3d(sQ + g * 2 * kAtcKB, &mapQ, q_full, p.qoff + h * kAttD, q0 + g * kAtcM, b);
        ptx::tma_load_3d(sQ + g * 2 * kAtcKB + kAtcKB, &mapQ, q_full, p.qoff + h * kAttD + 64, q0 + g * kAtcM, b);
      }
      for (int j = 0; j < ntiles; ++j) {
        const int buf = j & 1, u = j >> 1, k0 = j * kAtcN;
        ptx::mbar_wait(&k_empty[buf], (u & 1) ^ 1u);
        ptx::mbar_arrive_expect_tx(&k_full[buf], 2 * kAtcKB);
        ptx::tma_load_3d(sK + buf * 2 * kAtcKB, &mapK, &k_full[buf], p.koff + h * kAttD, k0, b);
        ptx::tma_load_3d(sK + buf * 2 * kAtcKB + kAtcKB, &mapK, &k_full[buf], p.koff + h * kAttD + 64, k0, b);
        ptx::mbar_wait(v_empty, (j & 1) ^ 1u);
        ptx::mbar_arrive_expect_tx(v_full, 2 * kAtcVB);
        ptx::tma_load_2d(sV, &mapVt, v_full, k0, (b * p.heads + h) * kAttD);
        ptx::tma_load_2d(sV + kAtcVB, &mapVt, v_full, k0 + 64, (b * p.heads + h) * kAttD);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idS = ptx::make_idesc_bf16_f32(kAtcM, kAtcN), idO = ptx::make_idesc_bf16_f32(kAtcM, kAttD);
    const bool el = ptx::elect_one();
    const int ng = tileB ? 2 : 1;
    auto issue_S = [&](int g, int kbuf) {       // S_g = Q_g K^T into TMEM columns g*128
      const uint32_t td = tmem_base + static_cast<uint32_t>(g * kAtcN);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t da = ptx::make_desc_k_sw128(ptx::smem_u32(sQ + g * 2 * kAtcKB + kb * kAtcKB));
        const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(sK + kbuf * 2 * kAtcKB + kb * kAtcKB));
        const int nsteps = kb == 0 ? 4 : 2;                       // head_dim 96 = 64 + 32
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < nsteps && el) ptx::mma_f16_ss(td, da + 2 * k, db + 2 * k, idS, (kb | k) != 0 ? 1u : 0u);
      }
      if (el) ptx::mma_commit(&s_full[g]);
      __syncwarp();
    };
    ptx::mbar_wait(q_full, 0);
    ptx::mbar_wait(&k_full[0], 0);
    ptx::tc_fence_after();
    for (int g = 0; g < ng; ++g) issue_S(g, 0);
    if (el) ptx::mma_commit(&k_empty[0]);
    __syncwarp();
    for (int j = 0; j < ntiles; ++j) {
      const bool more = j + 1 < ntiles;
      const int knext = (j + 1) & 1;
      for (int g = 0; g < ng; ++g) {
        ptx::mbar_wait(&p_full[g], j & 1);                        // P_g(j) published (=> S_g(j) fully read)
        ptx::mbar_wait(&o_empty[g], (j & 1) ^ 1u);                // O_g(j-1) read back
        if (g == 0) ptx::mbar_wait(v_full, j & 1);
        ptx::tc_fence_after();
        const uint32_t td = tmem_base + 256u + static_cast<uint32_t>(g * kAtcN);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t da = ptx::make_desc_k_sw128(ptx::smem_u32(sP + g * 2 * kAtcKB + kb * kAtcKB));
          const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(sV + kb * kAtcVB));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (el) ptx::mma_f16_ss(td, da + 2 * k, db + 2 * k, idO, (kb | k) != 0 ? 1u : 0u);
        }
        if (el) {
          ptx::mma_commit(&o_full[g]);
          ptx::mma_commit(&p_empty[g]);
          if (g == ng - 1) ptx::mma_commit(v_empty);              // V^T(j) consumed by every group
        }
        __syncwarp();
        if (more) {
          if (g == 0) {
            ptx::mbar_wait(&k_full[knext], ((j + 1) >> 1) & 1);
            ptx::tc_fence_after();
          }
          issue_S(g, knext);
          if (g == ng - 1) {
            if (el) ptx::mma_commit(&k_empty[knext]);
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + output: group g, thread = query row
    const int g = (warp - 2) >> 2;
    const int qw = warp & 3;                       // TMEM lane quarter this warp may read
    const int r = qw * 32 + lane;
    const int qrow = q0 + g * kAtcM + r;
    if (g == 0 || tileB) {
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(qw * 32) << 16);
      const uint32_t s_addr = lane_addr + static_cast<uint32_t>(g * kAtcN);
      const uint32_t o_addr = lane_addr + 256u + static_cast<uint32_t>(g * kAtcN);
      uint8_t* myP = sP + g * 2 * kAtcKB + r * 128;
      constexpr float kLog2e = 1.4426950408889634f;
      float o[kAttD];
#pragma unroll
      for (int i = 0; i < kAttD; ++i) o[i] = 0.f;
      float mrow = -INFINITY, lrow = 0.f;          // running max (natural units) and sum
      for (int j = 0; j < ntiles; ++j) {
        const int k0 = j * kAtcN;
        const bool masked = p.key_keep != nullptr || k0 + kAtcN > p.Tk;       // uniform per tile
        float* flags = sKeep + (g * 2 + (j & 1)) * kAtcN;
        if (masked) {
          const int k = k0 + r;
          flags[r] = k >= p.Tk ? -1.f : (p.key_keep ? p.key_keep[static_cast<size_t>(b) * p.Tk + k] : 1.f);
          if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
        }
        ptx::mbar_wait(&s_full[g], j & 1);
        ptx::tc_fence_after();
        uint32_t rr[32];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          ptx::tmem_ld_32x32b_x32(s_addr + c * 32, rr);
          ptx::tmem_wait_ld();
          if (masked) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float f = flags[c * 32 + i];
              mx = fmaxf(mx, f < 0.f ? -INFINITY : (f == 0.f ? -1e8f : __uint_as_float(rr[i])));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(rr[i]));
          }
        }
        const float mnew = fmaxf(mrow, mx);          // finite: key k0 of every tile is < Tk
        const float alpha = exp2f((mrow - mnew) * kLog2e);
        const float moff = mnew * kLog2e;
        ptx::mbar_wait(&p_empty[g], (j & 1) ^ 1u);   // the previous P.V MMA of this group has read sP
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          ptx::tmem_ld_32x32b_x32(s_addr + c * 32, rr);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int q8 = 0; q8 < 4; ++q8) {
            float pv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float sv = __uint_as_float(rr[q8 * 8 + i]);
              if (masked) {
                const float f = flags[c * 32 + q8 * 8 + i];
                sv = f == 0.f ? -1e8f : sv;
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(sv, kLog2e, -moff)));
                pv[i] = f < 0.f ? 0.f : e;
              } else {
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pv[i]) : "f"(fmaf(sv, kLog2e, -moff)));
              }
              sum += pv[i];
            }
            const int kc = c * 32 + q8 * 8;          // first key column of this 16-byte chunk
            *reinterpret_cast<uint4*>(myP + (kc >> 6) * kAtcKB + ((((kc & 63) >> 3) ^ (r & 7)) << 4)) =
                make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]), pack_bf16x2(pv[6], pv[7]));
          }
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();               // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[g]);
        lrow = lrow * alpha + sum;
        mrow = mnew;
        ptx::mbar_wait(&o_full[g], j & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          ptx::tmem_ld_32x32b_x32(o_addr + c * 32, rr);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha, __uint_as_float(rr[i]));
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&o_empty[g]);
      }
      if (qrow < p.Tq) {
        const float inv = 1.0f / lrow;
        __nv_bfloat16* dst = p.O + (static_cast<size_t>(b) * p.Tq + qrow) * p.ldo + h * kAttD;
#pragma unroll
        for (int c = 0; c < kAttD / 8; ++c)
          reinterpret_cast<uint4*>(dst)[c] = make_uint4(pack_bf16x2(o[8 * c] * inv, o[8 * c + 1] * inv), pack_bf16x2(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                                                       pack_bf16x2(o[8 * c + 4] * inv, o[8 * c + 5] * inv), pack_bf16x2(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
      }
      if (p.probs != nullptr) {
        // Head-averaged attention probabilities (transformer.py:398-419: the decoder's first layer returns them).  Single key tile
        // only (the host checks Tk <= 128): the score tile is still in TMEM and (mrow, lrow) are final, so the row is one more pass
        // of ex2 over it.  One vector reduction per 4 keys into the zeroed [B, Tq, Tk] buffer; with two heads the sum of the two
        // addends does not depend on their order.
        const float* flags = sKeep + (g * 2) * kAtcN;
        const bool masked = p.key_keep != nullptr || kAtcN > p.Tk;
        const float scale = 1.0f / (lrow * static_cast<float>(p.heads)), moff = mrow * kLog2e;
        float* prow = p.probs + (static_cast<size_t>(b) * p.Tq + (qrow < p.Tq ? qrow : 0)) * p.Tk;
        const bool vec = (p.Tk & 3) == 0;
        uint32_t rr[32];
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          ptx::tmem_ld_32x32b_x32(s_addr + c * 32, rr);          // warp-collective: every lane takes part, stores are predicated
          ptx::tmem_wait_ld();
          float pr[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float sv = __uint_as_float(rr[i]);
            const float f = masked ? flags[c * 32 + i] : 1.f;
            sv = f == 0.f ? -1e8f : sv;
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(sv, kLog2e, -moff)));
            pr[i] = f < 0.f ? 0.f : e * scale;
          }
          if (qrow < p.Tq) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const int k = c * 32 + i;
              if (vec && k + 3 < p.Tk) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(prow + k), "f"(pr[i]), "f"(pr[i + 1]), "f"(pr[i + 2]), "f"(pr[i + 3]) : "memory");
              } else {
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4)
                  if (k + e4 < p.Tk) atomicAdd(prow + k + e4, pr[i + e4]);
              }
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace
}  // namespace fse
