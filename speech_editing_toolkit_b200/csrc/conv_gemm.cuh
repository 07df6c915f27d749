// Channels-last 1-D convolution as a shifted GEMM, the one compute primitive of this library.
//
//   Out[(b,t), n] = sum_{tap, c} A0[b, t + off[tap], c] * W[n, (tap, c)]  +  sum_c A1[b, t, c] * W[n, (ntaps, c)]
//
// rows = frames (b,t) of a [B, T, C] channels-last activation, zero outside [0, T) (that IS the
// conv zero padding), columns = output channels, fp32 accumulation, result handed to an epilogue
// functor  epi.apply<NV>(b, t, n0, acc[NV])  (NV consecutive columns of one frame).
//
// Two back ends behind the same operand layout:
//   * conv_gemm_tc_kernel  — sm_100a tensor cores: TMA (cp.async.bulk.tensor, 128B/64B swizzle,
//     out-of-bounds zero fill = conv padding) -> shared memory ring -> tcgen05.mma (bf16 x bf16 -> fp32
//     in TMEM, M=128 frames, N<=256 channels) -> tcgen05.ld -> epilogue.  Warp-specialised:
//     warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator, warps 2..5 epilogue.
//   * conv_gemm_simt_kernel — plain CUDA-core tiled GEMM over fp32 or bf16 operands; the exact-fp32
//     mode of the library and the on-device cross-check of the tensor-core path.
#pragma once
#include <type_traits>
#include <cuda_bf16.h>

#include "ptx_sm100.cuh"

namespace fse {

constexpr int kMaxTaps = 12;

struct ConvGemmParams {
  int B;       // items processed by this launch
  int b_off;   // index of the first item (rows / TMA batch coordinate are b_off + local index)
  int Trows;   // output rows (frames) per item; tiles never straddle items
  int Tsrc;    // frames per item of the sources (SIMT bounds check; TC relies on TMA OOB fill)
  int C0, C1;  // channels of source 0 (with taps) and source 1 (single tap, offset 0; 0 = unused)
  int c_off0;  // first channel of source 0 inside its rows (a [.., ld0]-wide buffer holding several tensors)
  int ld0;     // row stride of source 0 in elements (>= c_off0 + C0)
  int ntaps;
  int tap_off[kMaxTaps];
  int KB;      // k-block width in channels: 64 (128B swizzle) or 32 (64B swizzle)
  int nkb0;    // k-blocks per tap for source 0 = ceil(C0/KB)
  int nkb1;    // k-blocks for source 1
  int N;       // output columns
  int Kp;      // padded K of the packed weight = (ntaps*nkb0 + nkb1)*KB
  int MT;        // 128-frame sub-tiles per job: one weight tile feeds MT activation tiles and the accumulator buffer
                 //    holds MT x BN columns - amortises the per-job handshakes and the weight reloads when BN is small
  int shared_a;  // 1: load each channel block of source 0 ONCE per job (Rrows = 128*MT + tap span rows) and feed every
                 //    (tap, sub-tile) from row-shifted UMMA descriptors of that one smem copy (activation ingest / ntaps)
  int off_min;   // smallest tap offset; Rrows = 128*MT + max(off) - min(off)
  int Rrows;
  int Rbox;      // shared-A: the job's rows arrive as nload TMA boxes of Rbox rows (<= 256, multiple of 8) laid end to end
  int nload;
  int bo_mode;   // descriptor base-offset convention for row-shifted starts (test hook; 1 = (addr>>7)&7)
  long long* dbg;  // optional [32] clock64 phase stamps of CTA 0 (test hook), else null
};

// Epilogue functors that set `static constexpr bool kTransposed = true` are run through a per-warp shared-memory
// transposition (conv_gemm_tc_kernel, CH == 32): 8 lanes cover 32 consecutive channels of one frame, so every global
// load / store instruction touches 4 row segments of 64-128 contiguous bytes instead of 32 rows x 16 bytes.
template <class E, class = void>
struct epi_transposed : std::false_type {};
template <class E>
struct epi_transposed<E, std::void_t<decltype(E::kTransposed)>> : std::bool_constant<E::kTransposed> {};
// `static constexpr int kLate = n`: the functor has n more operand sets that depend on stores of this very launch pattern
// (read-modify-write buffers) and are therefore fetched right before apply instead of one chunk ahead:
// load_late<NV>(b, t, n0, dst) fills aux[NV*kAux .. NV*(kAux+kLate)).
template <class E, class = void>
struct epi_late : std::integral_constant<int, 0> {};
template <class E>
struct epi_late<E, std::void_t<decltype(E::kLate)>> : std::integral_constant<int, E::kLate> {};
constexpr int kEpiScratchBytes = 8 * 32 * 32 * 4;     // kEpiWarps x [32 frames][32 fp32]

// ------------------------------------------------------------------------------------------
// small vector store helpers
// ------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void st_vec(float* p, const float* v) {
  static_assert(NV % 2 == 0, "NV");
  if constexpr (NV % 4 == 0) {
#pragma unroll
    for (int i = 0; i < NV / 4; ++i)
      reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < NV / 2; ++i) reinterpret_cast<float2*>(p)[i] = make_float2(v[2 * i], v[2 * i + 1]);
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <int NV>
__device__ __forceinline__ void st_vec(__nv_bfloat16* p, const float* v) {
  static_assert(NV % 2 == 0, "NV");
  if constexpr (NV % 8 == 0) {
#pragma unroll
    for (int i = 0; i < NV / 8; ++i)
      reinterpret_cast<uint4*>(p)[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                                  pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
  } else if constexpr (NV % 4 == 0) {
#pragma unroll
    for (int i = 0; i < NV / 4; ++i)
      reinterpret_cast<uint2*>(p)[i] = make_uint2(pack_bf16x2(v[4 * i], v[4 * i + 1]), pack_bf16x2(v[4 * i + 2], v[4 * i + 3]));
  } else {
#pragma unroll
    for (int i = 0; i < NV / 2; ++i) reinterpret_cast<uint32_t*>(p)[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
  }
}
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// ------------------------------------------------------------------------------------------
// CUDA-core back end
// ------------------------------------------------------------------------------------------
template <typename TOp, class Epi>
__global__ void __launch_bounds__(256) conv_gemm_simt_kernel(ConvGemmParams p, const TOp* __restrict__ A0,
                                                             const TOp* __restrict__ A1, const TOp* __restrict__ W,
                                                             Epi epi) {
  __shared__ float As[16][68];
  __shared__ float Ws[16][68];
  const int tiles_per_item = (p.Trows + 63) / 64;
  const int b = p.b_off + blockIdx.x / tiles_per_item;
  const int t0 = (blockIdx.x % tiles_per_item) * 64;
  const int n0 = blockIdx.y * 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int sub = p.KB / 16;
  const int nk16 = p.Kp / 16;
  const int nkb_src0 = p.ntaps * p.nkb0;
  for (int j = 0; j < nk16; ++j) {
    const int kb = j / sub;
    const TOp* src;
    int C, cb, off, ld;
    if (kb < nkb_src0) {
      const int tap = kb / p.nkb0;
      src = A0 + p.c_off0; C = p.C0; off = p.tap_off[tap]; ld = p.ld0;
      cb = (kb % p.nkb0) * p.KB + (j % sub) * 16;
    } else {
      src = A1; C = p.C1; off = 0; ld = p.C1;
      cb = (kb - nkb_src0) * p.KB + (j % sub) * 16;
    }
    {
      const int tt = t0 + lrow + off;
      const bool ok = (tt >= 0) && (tt < p.Tsrc);
      const TOp* rp = src + (static_cast<size_t>(b) * p.Tsrc + (ok ? tt : 0)) * ld;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int ch = cb + lk + q;
        As[lk + q][lrow] = (ok && ch < C) ? to_f32(rp[ch]) : 0.f;
      }
      const int n = n0 + lrow;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        Ws[lk + q][lrow] = (n < p.N) ? to_f32(W[static_cast<size_t>(n) * p.Kp + j * 16 + lk + q]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(av[i], wv[jj], acc[i][jj]);
    }
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n < p.N) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + ty * 4 + i;
      if (t < p.Trows) {
        float aux[4 * (Epi::kAux + epi_late<Epi>::value > 0 ? Epi::kAux + epi_late<Epi>::value : 1)];
        if constexpr (Epi::kAux > 0) epi.template load_aux<4>(b, t, n, aux);
        if constexpr (epi_late<Epi>::value > 0) epi.template load_late<4>(b, t, n, aux + 4 * Epi::kAux);
        epi.template apply<4>(b, t, n, acc[i], aux);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// tcgen05 back end: persistent, warp-specialised, double-buffered accumulators
// ------------------------------------------------------------------------------------------
//   warp 0      TMA producer: streams (A tile, W tile) k-blocks of successive output tiles through an
//               smem ring (full/empty mbarriers)
//   warp 1      MMA issuer + TMEM owner: tcgen05.mma into accumulator buffer (tile & 1); tcgen05.commit
//               releases smem slots and signals acc_full
//   warps 2..9  epilogue: tcgen05.ld the finished buffer (lane = frame, CH channels at a time), run the
//               functor with prefetched residual operands, release the buffer (acc_empty)
// so the epilogue of tile i overlaps the loads and MMAs of tile i+1.  One CTA per SM, grid = min(tiles, SMs).
constexpr int kEpiWarps = 8;
constexpr int kTcThreads = 64 + 32 * kEpiWarps;
constexpr int kTileM = 128;       // frames per tile = TMEM lanes

// rb = bytes of one k-block row = KB elements x sizeof(operand): 128 (128B swizzle) or 64 (64B swizzle)
__host__ __device__ inline int tc_b_stage_bytes(int BN, int rb) { return ((BN * rb + 1023) / 1024) * 1024; }
__host__ __device__ inline int tc_a_stage_bytes(int rb) { return kTileM * rb; }
__host__ inline size_t tc_smem_bytes(int BN, int rb, int stages, int a_slots, int a_slot_bytes, int scratch_bytes = 0) {
  return 1024 + static_cast<size_t>(a_slots) * a_slot_bytes + static_cast<size_t>(stages) * tc_b_stage_bytes(BN, rb) + scratch_bytes +
         8 * (2 * stages + 4 + 2 * a_slots) + 16;
}

// TOp = __nv_bfloat16: tcgen05 kind::f16 (bf16 x bf16 -> fp32), K-blocks of 64 / 32 channels.
// TOp = float:         tcgen05 kind::tf32 (the reference's own GPU arithmetic: cuDNN TF32 convolutions), fp32 operands in
//                      HBM and shared memory, K-blocks of 32 / 16 channels.  Same bytes per k-block row (128 / 64), same
//                      swizzle atoms, same descriptors; one MMA instruction covers 32 bytes of K in both kinds.
template <typename TOp, int KB, int CH, class Epi>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                    const __grid_constant__ CUtensorMap mapW, ConvGemmParams p, int BN, int stages, int a_slots,
                    int a_slot_bytes, int scratch_bytes, Epi epi) {
  constexpr int ES = static_cast<int>(sizeof(TOp));
  constexpr int RB = KB * ES;                 // bytes per k-block row
  constexpr bool kTF32 = std::is_same<TOp, float>::value;
  static_assert(RB == 128 || RB == 64, "k-block row must be 128 or 64 bytes");
  static_assert(CH == 32 || CH == 16, "epilogue chunk");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int A_BYTES = a_slot_bytes;        // non-shared mode: a_slots == stages, one A tile per W stage
  const int B_BYTES = tc_b_stage_bytes(BN, RB);
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_slots * A_BYTES;
  uint8_t* sScratch = sB + stages * B_BYTES;       // transposed epilogue: [kEpiWarps][32][32] fp32
  uint64_t* full = reinterpret_cast<uint64_t*>(sScratch + scratch_bytes);
  uint64_t* empty = full + stages;
  uint64_t* acc_full = empty + stages;     // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint64_t* a_full = acc_empty + 2;        // [a_slots] (shared-A mode)
  uint64_t* a_empty = a_full + a_slots;    // [a_slots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + a_slots);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  long long* dbg = (p.dbg && blockIdx.x == 0) ? p.dbg : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = clock64();
  const int MT = p.MT > 0 ? p.MT : 1;
  const int tile_rows = kTileM * MT;
  const int tiles_per_item = (p.Trows + tile_rows - 1) / tile_rows;
  const int n_tiles = p.N / BN;
  const int total_tiles = p.B * tiles_per_item * n_tiles;
  const int nkb_src0 = p.ntaps * p.nkb0;
  const int nkb = nkb_src0 + p.nkb1;
  uint32_t ncols = 32;
  // two accumulator buffers (the epilogue of job i overlaps the MMAs of job i+1) when they fit the 512 TMEM columns, else one
  const int nacc = 2 * MT * BN <= 512 ? 2 : 1;
  while (static_cast<int>(ncols) < nacc * MT * BN) ncols <<= 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&mapA0);
    if (p.nkb1 > 0) ptx::prefetch_tensormap(&mapA1);
    ptx::prefetch_tensormap(&mapW);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        ptx::mbar_init(&full[s], 1);
        ptx::mbar_init(&empty[s], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&acc_full[i], 1);
        ptx::mbar_init(&acc_empty[i], kEpiWarps);
      }
      for (int i = 0; i < a_slots; ++i) {
        ptx::mbar_init(&a_full[i], 1);
        ptx::mbar_init(&a_empty[i], 1);
      }
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
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may
  // overlap the tail of the previous kernel in the stream; no global memory is touched before this point.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0 && p.shared_a) {
      int kbg = 0, ga = 0;
      const int ngroups = p.nkb0 + p.nkb1;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m = tile / n_tiles, n0 = (tile % n_tiles) * BN;
        const int b = p.b_off + m / tiles_per_item, t0 = (m % tiles_per_item) * tile_rows;
        for (int g = 0; g < ngroups; ++g, ++ga) {
          const int slot = ga % a_slots;
          ptx::mbar_wait(&a_empty[slot], ((ga / a_slots) & 1) ^ 1u);
          const bool src0 = g < p.nkb0;
          const int box_bytes = p.Rbox * RB;
          ptx::mbar_arrive_expect_tx(&a_full[slot], static_cast<uint32_t>(src0 ? p.nload * box_bytes : kTileM * RB));
          if (src0) {
            for (int i = 0; i < p.nload; ++i)
              ptx::tma_load_3d(sA + slot * A_BYTES + i * box_bytes, &mapA0, &a_full[slot], p.c_off0 + g * KB, t0 + p.off_min + i * p.Rbox, b);
          } else {
            ptx::tma_load_3d(sA + slot * A_BYTES, &mapA1, &a_full[slot], (g - p.nkb0) * KB, t0, b);   // source 1: MT == 1 only
          }
          const int ntap = src0 ? p.ntaps : 1;
          for (int j = 0; j < ntap; ++j, ++kbg) {
            const int s = kbg % stages;
            ptx::mbar_wait(&empty[s], ((kbg / stages) & 1) ^ 1u);
            ptx::mbar_arrive_expect_tx(&full[s], static_cast<uint32_t>(BN * RB));
            const int kb = src0 ? j * p.nkb0 + g : nkb_src0 + (g - p.nkb0);
            ptx::tma_load_2d(sB + s * B_BYTES, &mapW, &full[s], kb * KB, n0);
          }
        }
        if (dbg && tile == blockIdx.x) dbg[2] = clock64();
      }
    } else if (lane == 0) {
      constexpr int A1 = kTileM * RB;            // one 128-frame activation tile
      const uint32_t tx_bytes = static_cast<uint32_t>(MT * A1 + BN * RB);
      int kbg = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m = tile / n_tiles, n0 = (tile % n_tiles) * BN;
        const int b = p.b_off + m / tiles_per_item, t0 = (m % tiles_per_item) * tile_rows;
        for (int kb = 0; kb < nkb; ++kb, ++kbg) {
          const int s = kbg % stages;
          const uint32_t ph = (kbg / stages) & 1;
          ptx::mbar_wait(&empty[s], ph ^ 1u);
          ptx::mbar_arrive_expect_tx(&full[s], tx_bytes);
          for (int mt = 0; mt < MT; ++mt) {
            if (kb < nkb_src0) {
              const int tap = kb / p.nkb0;
              ptx::tma_load_3d(sA + s * A_BYTES + mt * A1, &mapA0, &full[s], p.c_off0 + (kb % p.nkb0) * KB,
                               t0 + mt * kTileM + p.tap_off[tap], b);
            } else {
              ptx::tma_load_3d(sA + s * A_BYTES + mt * A1, &mapA1, &full[s], (kb - nkb_src0) * KB, t0 + mt * kTileM, b);
            }
          }
          ptx::tma_load_2d(sB + s * B_BYTES, &mapW, &full[s], kb * KB, n0);
        }
        if (dbg && tile == blockIdx.x) dbg[2] = clock64();      // all loads of the first tile issued
      }
    }
    __syncwarp();   // reconverge: bar.sync below must be reached by whole warps
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = kTF32 ? ptx::make_idesc_tf32_f32(kTileM, BN) : ptx::make_idesc_bf16_f32(kTileM, BN);
    auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
      if constexpr (kTF32) ptx::mma_tf32_ss(d, da, db, idesc, acc); else ptx::mma_f16_ss(d, da, db, idesc, acc);
    };
    const bool el = ptx::elect_one();      // the one lane that issues every tcgen05.mma / commit of this CTA
    int kbg = 0, it = 0, ga = 0;
    if (p.shared_a) {
      const int ngroups = p.nkb0 + p.nkb1;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int buf = nacc == 2 ? (it & 1) : 0, use = nacc == 2 ? (it >> 1) : it;
        ptx::mbar_wait(&acc_empty[buf], (use & 1) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * MT * BN);
        uint32_t accum = 0;
        for (int g = 0; g < ngroups; ++g, ++ga) {
          const int slot = ga % a_slots;
          ptx::mbar_wait(&a_full[slot], (ga / a_slots) & 1);
          const bool src0 = g < p.nkb0;
          const int ntap = src0 ? p.ntaps : 1;
          for (int j = 0; j < ntap; ++j, ++kbg) {
            const int s = kbg % stages;
            ptx::mbar_wait(&full[s], (kbg / stages) & 1);
            ptx::tc_fence_after();
            if (dbg && lane == 0 && kbg == 0) dbg[3] = clock64();
            {
              // row-shifted views of the one smem copy: (tap j, sub-tile mt) reads rows [mt*128 + shift, +128) of the job's rows.
              // Descriptors are computed by the whole warp (convergent code -> uniform registers); only the tcgen05
              // instructions sit under the one-lane predicate (a divergent region would cost ~7 R2UR + a waterfall per MMA).
              const int shift = src0 ? p.tap_off[j] - p.off_min : 0;
              const uint32_t b_addr = ptx::smem_u32(sB + s * B_BYTES);
              const uint64_t db = (RB == 128) ? ptx::make_desc_k_sw128(b_addr) : ptx::make_desc_k_sw64(b_addr);
              const uint32_t a_base = ptx::smem_u32(sA + slot * A_BYTES) + static_cast<uint32_t>(shift * RB);
              for (int mt = 0; mt < MT; ++mt) {
                const uint32_t a_addr = a_base + static_cast<uint32_t>(mt * kTileM * RB);
                const uint32_t bo = p.bo_mode ? ((a_addr >> 7) & (RB == 128 ? 7u : 3u)) : 0u;
                const uint64_t da = ((RB == 128) ? ptx::make_desc_k_sw128(a_addr) : ptx::make_desc_k_sw64(a_addr)) |
                                    (static_cast<uint64_t>(bo) << 49);
                const uint32_t td = tmem_d + static_cast<uint32_t>(mt * BN);
#pragma unroll
                for (int k = 0; k < RB / 32; ++k)
                  if (el) mma(td, da + 2 * k, db + 2 * k, accum | (k != 0 ? 1u : 0u));
              }
              if (el) ptx::mma_commit(&empty[s]);
            }
            accum = 1;
            __syncwarp();
          }
          if (el) {
            ptx::mma_commit(&a_empty[slot]);
            if (g == ngroups - 1) {
              ptx::mma_commit(&acc_full[buf]);
              if (dbg && it == 0) dbg[4] = clock64();
            }
          }
          __syncwarp();
        }
      }
    } else
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int buf = nacc == 2 ? (it & 1) : 0, use = nacc == 2 ? (it >> 1) : it;
      ptx::mbar_wait(&acc_empty[buf], (use & 1) ^ 1u);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * MT * BN);
      for (int kb = 0; kb < nkb; ++kb, ++kbg) {
        const int s = kbg % stages;
        const uint32_t ph = (kbg / stages) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        if (dbg && lane == 0 && kbg == 0) dbg[3] = clock64();   // first k-block landed
        {
          const uint32_t b_addr = ptx::smem_u32(sB + s * B_BYTES);
          const uint64_t db = (RB == 128) ? ptx::make_desc_k_sw128(b_addr) : ptx::make_desc_k_sw64(b_addr);
          const uint32_t a_base = ptx::smem_u32(sA + s * A_BYTES);
          const uint32_t acc = kb != 0 ? 1u : 0u;
          for (int mt = 0; mt < MT; ++mt) {          // the same weight tile against MT activation sub-tiles
            const uint32_t a_addr = a_base + static_cast<uint32_t>(mt * (kTileM * RB));
            const uint64_t da = (RB == 128) ? ptx::make_desc_k_sw128(a_addr) : ptx::make_desc_k_sw64(a_addr);
            const uint32_t td = tmem_d + static_cast<uint32_t>(mt * BN);
#pragma unroll
            for (int k = 0; k < RB / 32; ++k)   // +32 bytes along K inside the swizzle atom = +2 in the addr>>4 field
              if (el) mma(td, da + 2 * k, db + 2 * k, k != 0 ? 1u : acc);
          }
          if (el) {
            ptx::mma_commit(&empty[s]);
            if (kb == nkb - 1) ptx::mma_commit(&acc_full[buf]);
            if (dbg && kb == nkb - 1 && it == 0) dbg[4] = clock64();   // all MMAs of the first tile issued
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    // Lane = frame (TMEM lane), CH consecutive channels per tcgen05.ld: every lane touches CH*4 (or CH*2)
    // contiguous bytes of its own row, i.e. whole 32-byte sectors.  Residual operands (Epi::kAux floats per
    // column) of the NEXT chunk are requested before the current chunk is processed, and the first chunk's
    // before the accumulator is even complete, so their latency hides behind the MMAs / the previous chunk.
    const int ew = warp - 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = ew >> 2;               // two warps share a quarter and split the column chunks
    const int cpb = BN / CH;                 // chunks per 128-frame sub-tile
    const int nchunks = MT * cpb;
    if constexpr (CH == 32 && epi_transposed<Epi>::value) {
      // ---- transposed epilogue: accumulator chunk -> registers -> XOR-swizzled scratch -> lanes re-assigned so that 8 lanes
      // cover the 32 channels of one frame (4 channels each) and a warp instruction covers 4 frames.  Residual operands of
      // the NEXT chunk are requested while the current one is processed (first chunk: before the accumulator is awaited).
      float* stg = reinterpret_cast<float*>(sScratch) + ew * 1024;
      const int cq = lane & 7, r0 = lane >> 3;
      constexpr int AP = Epi::kAux > 0 ? 4 * Epi::kAux : 1;                                   // prefetched part
      constexpr int AX = Epi::kAux + epi_late<Epi>::value > 0 ? 4 * (Epi::kAux + epi_late<Epi>::value) : 1;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int buf = nacc == 2 ? (it & 1) : 0, use = nacc == 2 ? (it >> 1) : it;
        const int m = tile / n_tiles, n0 = (tile % n_tiles) * BN;
        const int b = p.b_off + m / tiles_per_item, t0 = (m % tiles_per_item) * tile_rows;
        const int tq = t0 + q * 32 + r0;          // frame of iteration 0 in sub-tile 0; iteration i adds 4*i, sub-tile st adds 128*st
        auto T_OF = [&](int c, int i) { return tq + (c / cpb) * kTileM + 4 * i; };
        auto N_OF = [&](int c) { return n0 + (c % cpb) * CH + cq * 4; };
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * MT * BN);
        float aux[8][AX], aux_next[8][AP];
        if constexpr (Epi::kAux > 0) {
          if (half < nchunks) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (T_OF(half, i) < p.Trows) epi.template load_aux<4>(b, T_OF(half, i), N_OF(half), aux[i]);
          }
        }
        ptx::mbar_wait(&acc_full[buf], use & 1);
        ptx::tc_fence_after();
        if (dbg && ew == 0 && lane == 0 && it == 0) dbg[5] = clock64();   // first accumulator complete
        for (int c = half; c < nchunks; c += 2) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(lane_base + c * CH, r);
          ptx::tmem_wait_ld();
          if constexpr (Epi::kAux > 0) {
            if (c + 2 < nchunks) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (T_OF(c + 2, i) < p.Trows) epi.template load_aux<4>(b, T_OF(c + 2, i), N_OF(c + 2), aux_next[i]);
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
              if (T_OF(c, i) < p.Trows) epi.template load_late<4>(b, T_OF(c, i), nn, aux[i] + 4 * Epi::kAux);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rloc = 4 * i + r0;
            const float4 a = *reinterpret_cast<const float4*>(stg + rloc * 32 + ((cq ^ (rloc & 7)) << 2));
            const int t = T_OF(c, i);
            if (t < p.Trows) {
              const float v[4] = {a.x, a.y, a.z, a.w};
              epi.template apply<4>(b, t, nn, v, aux[i]);
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
        if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
        if (dbg && lane == 0 && it == 0) dbg[6 + ew] = clock64();           // epilogue warp ew done with the first tile
      }
    } else {
    constexpr int AUXN = CH * (Epi::kAux > 0 ? Epi::kAux : 1);
    auto apply_chunk = [&](int bb, int t, int nn, const float* v, const float* auxp) {
      if constexpr (epi_late<Epi>::value > 0) {
        float comb[CH * (Epi::kAux + epi_late<Epi>::value)];
#pragma unroll
        for (int j = 0; j < CH * Epi::kAux; ++j) comb[j] = auxp[j];
        epi.template load_late<CH>(bb, t, nn, comb + CH * Epi::kAux);
        epi.template apply<CH>(bb, t, nn, v, comb);
      } else {
        epi.template apply<CH>(bb, t, nn, v, auxp);
      }
    };
    // Residual operands: when a warp owns at most kMaxPre chunks of a tile and they fit in 64 registers, ALL of
    // them are requested before the accumulator is awaited (one exposed memory latency per tile); otherwise the
    // next chunk's operands are requested while the current chunk is processed (when they fit in 32 registers).
    constexpr int kMaxPre = (Epi::kAux > 0 && AUXN <= 32) ? 64 / AUXN : 0;
    constexpr bool kChunkAhead = Epi::kAux > 0 && AUXN <= 32;
    const int my_chunks = (nchunks - half + 1) / 2;
    const bool pre_all = kMaxPre > 0 && my_chunks <= kMaxPre;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int buf = nacc == 2 ? (it & 1) : 0, use = nacc == 2 ? (it >> 1) : it;
      const int m = tile / n_tiles, n0 = (tile % n_tiles) * BN;
      const int b = p.b_off + m / tiles_per_item, t0 = (m % tiles_per_item) * tile_rows;
      const int tl0 = t0 + q * 32 + lane;      // frame of this lane in sub-tile 0; chunk c lives in sub-tile c / cpb
      auto T_OF = [&](int c) { return tl0 + (c / cpb) * kTileM; };
      auto N_OF = [&](int c) { return n0 + (c % cpb) * CH; };
      const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * MT * BN);
      float aux[kMaxPre > 0 ? kMaxPre : 1][AUXN];
      if constexpr (kChunkAhead) {
        if (pre_all) {
#pragma unroll
          for (int i = 0; i < kMaxPre; ++i)
            if (i < my_chunks && T_OF(half + 2 * i) < p.Trows) epi.template load_aux<CH>(b, T_OF(half + 2 * i), N_OF(half + 2 * i), aux[i]);
        } else if (half < nchunks && T_OF(half) < p.Trows) {
          epi.template load_aux<CH>(b, T_OF(half), N_OF(half), aux[0]);
        }
      }
      ptx::mbar_wait(&acc_full[buf], use & 1);
      ptx::tc_fence_after();
      if (dbg && ew == 0 && lane == 0 && it == 0) dbg[5] = clock64();   // first accumulator complete
      int ci = 0;
      for (int c = half; c < nchunks; c += 2, ++ci) {
        uint32_t r[CH];
        const int t = T_OF(c);
        const bool row_ok = t < p.Trows;
        const bool stamp = dbg && ew == 0 && lane == 0 && it == 0 && c < 8;
        if (stamp) dbg[16 + 4 * (c >> 1)] = clock64();
        float aux_here[(Epi::kAux > 0 && !kChunkAhead) ? AUXN : 1];
        if constexpr (Epi::kAux > 0 && !kChunkAhead) {
          if (row_ok) epi.template load_aux<CH>(b, t, N_OF(c), aux_here);
        }
        if constexpr (CH == 32) ptx::tmem_ld_32x32b_x32(lane_base + c * CH, r);
        else ptx::tmem_ld_32x32b_x16(lane_base + c * CH, r);
        ptx::tmem_wait_ld();
        if (stamp) dbg[17 + 4 * (c >> 1)] = clock64();
        float aux_next[kChunkAhead ? AUXN : 1];
        if constexpr (kChunkAhead) {
          if (!pre_all && c + 2 < nchunks && T_OF(c + 2) < p.Trows) epi.template load_aux<CH>(b, T_OF(c + 2), N_OF(c + 2), aux_next);
        }
        if (stamp) dbg[18 + 4 * (c >> 1)] = clock64();
        if (row_ok) {
          float v[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(r[i]);
          if constexpr (kChunkAhead) {
            if (pre_all) {
              // static register indexing: pick the pre-loaded set of this chunk
#pragma unroll
              for (int i = 0; i < kMaxPre; ++i)
                if (i == ci) apply_chunk(b, t, N_OF(c), v, aux[i]);
            } else {
              apply_chunk(b, t, N_OF(c), v, aux[0]);
            }
          } else {
            apply_chunk(b, t, N_OF(c), v, aux_here);
          }
        }
        if constexpr (kChunkAhead) {
          if (!pre_all) {
#pragma unroll
            for (int j = 0; j < AUXN; ++j) aux[0][j] = aux_next[j];
          }
        }
        if (stamp) dbg[19 + 4 * (c >> 1)] = clock64();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
      if (dbg && lane == 0 && it == 0) dbg[6 + ew] = clock64();           // epilogue warp ew done with the first tile
    }
    }
  }
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[14] = clock64();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, ncols);
  }
}

}  // namespace fse
