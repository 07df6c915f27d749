// All residual layers of one DiffNet evaluation in ONE persistent launch (diffnet.py:68-81 x L).
//
// Per layer l and 128-frame tile, one CTA runs three tensor-core jobs back to back, alternating between two
// 256-column TMEM accumulators (job j uses buffer j & 1, epilogue j must have drained it before job j+2):
//   G1a, G1b : y[:, half] = conv_k3(hb_l) W_dc^T + cond W_cp^T        K = 3*256 + H, N = 2 x 256 (gate/filter interleaved)
//              epilogue: u = sigmoid(gate) * tanh(filter) (+ fp32 timestep bias) -> u_all[:, l*C ...] in HBM (for the
//              folded skip GEMM) AND into shared memory in the UMMA K-major 128B-swizzled layout
//   G2       : o = u W_op[:C]^T                                          K = 256 (A operand = the smem copy of u), N = 256
//              epilogue: h <- (h + o + b) / sqrt(2)  (fp32, in place)  and hb_{l+1} (bf16, ping-pong buffer)
// so u never round-trips through HBM between the two GEMMs, the residual epilogue hides behind the next tile's
// MMAs, and the 40 launches per step (with their drain / fill / TMEM alloc) collapse into one.  The activation tile of
// each 64-channel block is loaded once (130 rows) and the three conv taps are row-shifted UMMA descriptors of it.
// Layers are separated by a grid-wide barrier (the conv halo of layer l+1 reads hb rows written by neighbour CTAs);
// the launch is cooperative so that all CTAs are co-resident.
#pragma once
#include "epilogues.cuh"

namespace fse {

constexpr int kFC = 256;                       // residual channels handled by the fused kernel
constexpr int kFusedASlots = 3, kFusedWStages = 3;
constexpr int kFusedASlotBytes = 17408;        // 130 rows x 128 B rounded up to 1024
constexpr int kFusedWStageBytes = 256 * 128;   // 256 weight rows x 64 bf16
constexpr int kFusedUBytes = 4 * 128 * 128;    // u tile: 4 k-blocks of [128 rows x 64 bf16]
constexpr int kFusedBiasBytes = 3 * 512 * 4 + 256 * 4;   // per-layer timestep tables (m, a, c) + residual bias
constexpr size_t kFusedSmemBytes = 1024 + kFusedASlots * kFusedASlotBytes + kFusedWStages * kFusedWStageBytes + kFusedUBytes + kFusedBiasBytes + 256;

struct FusedParams {
  int B, b_off, T, L, H;          // items of this launch, first item, frames per item, layers, cond channels
  float* h;                       // [*, 256] fp32 residual stream, updated in place
  __nv_bfloat16* hb0;             // operand copy read by even layers / written by odd layers
  __nv_bfloat16* hb1;             // ... and vice versa
  float* hf0;                     // tf32 streamed kernel: fp32 residual stream = operand, read by even / written by odd layers
  float* hf1;                     // ... and vice versa
  __nv_bfloat16* u_all;           // [*, L*256]
  const float* dbias;             // timestep tables of this call: [.., L, 3, 512]
  long long dbias_bstride;        // elements between consecutive items' tables (0: shared)
  const float* b2;                // [L, 256] residual half of output_projection.bias
  unsigned int* grid_bar;         // zeroed before the launch (lock-step kernel)
  unsigned int* done;             // [L, units] publication counters, zero when the launch starts (denoiser_stream.cuh)
  unsigned int* done_clear;       // optional: the OTHER counter array, zeroed by this launch for the next one (programmatic dependent launch)
  unsigned int done_clear_n;
  const CUtensorMap* mW1;         // [L] in global memory: [512, 960] gate weights, box 64 x 256
  const CUtensorMap* mW2;         // [L]: [256, 256] residual weights, box 64 x 256
  const CUtensorMap* mW1p;        // [L] same tensors with box 64 x 128 (CTA-pair mode: each CTA loads half of every tile)
  const CUtensorMap* mW2p;        // [L]
  long long* dbg;                 // optional [64] clock64 stamps of CTA 0 in layer 3 (developer aid)
};

// fp32 residual stream: read once per layer, never reused from L1
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void fused_grid_barrier(unsigned int* bar, unsigned int target) {
  __threadfence();
  atomicAdd(bar, 1u);
  const long long t0 = clock64();
  unsigned int v;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    if (clock64() - t0 > 4000000000LL) {
      printf("fse: grid barrier timed out (block %d, target %u, seen %u)\n", blockIdx.x, target, v);
      __trap();
    }
  } while (v < target);
}

// kPair: the two CTAs of a cluster process two adjacent tiles with M = 256 tcgen05.mma.cta_group::2 instructions
// issued by the leader; each CTA loads only HALF of every weight tile (the per-SM L2->SM ingest is the limiter).
template <bool kPair, bool kSharedA>
__global__ void __launch_bounds__(kTcThreads, 1)
denoiser_layers_kernel(const __grid_constant__ CUtensorMap mapHb0, const __grid_constant__ CUtensorMap mapHb1,
                       const __grid_constant__ CUtensorMap mapCond, FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int WB = kPair ? 128 * 128 : 256 * 128;     // bytes of one weight stage (128 or 256 rows x 64 bf16)
  // kSharedA: one 130-row activation load per channel block, taps = row-shifted descriptors (less ingest, but the
  // misaligned operand fetch slows the MMA); otherwise one 128-row load per (tap, channel block).
  constexpr int AS = kSharedA ? kFusedASlots : 4;
  constexpr int AB = kSharedA ? kFusedASlotBytes : 128 * 128;
  constexpr int WS = kPair ? (kSharedA ? 6 : 5) : kFusedWStages;
  static_assert(AS * AB + WS * WB <= kFusedASlots * kFusedASlotBytes + kFusedWStages * kFusedWStageBytes, "smem budget");
  constexpr uint32_t kMul = kPair ? 2u : 1u;
  uint8_t* sA = smem;
  uint8_t* sW = sA + AS * AB;
  uint8_t* sU = sW + WS * WB;
  float* sBias = reinterpret_cast<float*>(sU + kFusedUBytes);          // [3][512] timestep tables + [256] residual bias of the layer
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sU + kFusedUBytes + kFusedBiasBytes);
  uint64_t* a_empty = a_full + AS;
  uint64_t* w_full = a_empty + AS;
  uint64_t* w_empty = w_full + WS;
  uint64_t* acc_full = w_empty + WS;              // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint64_t* u_full = acc_empty + 2;
  uint64_t* u_empty = u_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(u_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int tiles_per_item = (p.T + kTileM - 1) / kTileM;
  const int total_tiles = p.B * tiles_per_item;
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  // work units: a tile (or, in pair mode, two adjacent tiles 2u, 2u+1 handled by the CTAs of a cluster)
  const int unit0 = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int unit_stride = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int total_units = kPair ? (total_tiles + 1) / 2 : total_tiles;
  const int nkbH = (p.H + 63) / 64;
  const int nhb = kSharedA ? 4 : 12;               // hb groups: 4 channel blocks (3 taps each) or 12 (tap, block) pairs
  const int ngroups = nhb + nkbH;                  // + cond blocks (1 tap)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&mapHb0);
    ptx::prefetch_tensormap(&mapHb1);
    ptx::prefetch_tensormap(&mapCond);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < AS; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
      for (int i = 0; i < WS; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
      for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], kEpiWarps * kMul); }
      ptx::mbar_init(u_full, 2 * kEpiWarps * kMul);
      ptx::mbar_init(u_empty, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    if constexpr (kPair) { ptx::tmem_alloc_pair(tmem_slot, 512); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (kPair) ptx::cluster_sync_all();     // the peer's barriers are initialised before any remote arrive / TMA signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // monotonic pipeline counters (continue across tiles and layers)
  int ga = 0, kw = 0, it = 0;

  for (int l = 0; l < p.L; ++l) {
    // Per-layer bias tables -> shared memory (L1 is streamed through by the fp32 residual traffic, so per-chunk __ldg of
    // the tables kept missing).  With per-item tables (dbias_bstride != 0) only the residual bias is staged.
    if (warp >= 2) {
      const int e = threadIdx.x - 64;                      // 0..255
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0;
      const float4* tab = reinterpret_cast<const float4*>(p.dbias + static_cast<size_t>(l) * 3 * 512);
      if (p.dbias_bstride == 0) {                          // 384 float4: all loads in flight before the first store
        v0 = __ldg(tab + e);
        if (e < 128) v1 = __ldg(tab + 256 + e);
      }
      if (e < 64) v2 = __ldg(reinterpret_cast<const float4*>(p.b2 + static_cast<size_t>(l) * kFC) + e);
      if (p.dbias_bstride == 0) {
        reinterpret_cast<float4*>(sBias)[e] = v0;
        if (e < 128) reinterpret_cast<float4*>(sBias)[256 + e] = v1;
      }
      if (e < 64) reinterpret_cast<float4*>(sBias + 3 * 512)[e] = v2;
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");      // epilogue warps only
    }
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      if (lane == 0) {
        const CUtensorMap* mHb = (l & 1) ? &mapHb1 : &mapHb0;
        const CUtensorMap* mW1 = (kPair ? p.mW1p : p.mW1) + l;
        const CUtensorMap* mW2 = (kPair ? p.mW2p : p.mW2) + l;
        const int nrow = kPair ? static_cast<int>(rank) * 128 : 0;       // this CTA's half of every weight tile
        for (int unit = unit0; unit < total_units; unit += unit_stride) {
          const int tile = kPair ? 2 * unit + static_cast<int>(rank) : unit;
          const int b = p.b_off + tile / tiles_per_item, t0 = (tile % tiles_per_item) * kTileM;
          for (int half = 0; half < 2; ++half) {
            for (int g = 0; g < ngroups; ++g, ++ga) {
              const int slot = ga % AS;
              ptx::mbar_wait(&a_empty[slot], ((ga / AS) & 1) ^ 1u);
              // pair mode: both CTAs' loads signal the LEADER's barrier, which expects the bytes of both
              const bool is_hb = g < nhb;
              const uint32_t rows = (kSharedA && is_hb) ? 130u : 128u;
              const int c0 = is_hb ? (kSharedA ? g : (g & 3)) * 64 : (g - nhb) * 64;
              const int tt = is_hb ? (kSharedA ? t0 - 1 : t0 + (g >> 2) - 1) : t0;      // non-shared: tap g/4 has offset g/4 - 1
              if (leader) ptx::mbar_arrive_expect_tx(&a_full[slot], rows * 128u * kMul);
              if constexpr (kPair) {
                ptx::tma_load_3d_pair(sA + slot * AB, is_hb ? mHb : &mapCond, ptx::mapa_u32(ptx::smem_u32(&a_full[slot]), 0), c0, tt, b);
              } else {
                ptx::tma_load_3d(sA + slot * AB, is_hb ? mHb : &mapCond, &a_full[slot], c0, tt, b);
              }
              const int ntap = (kSharedA && is_hb) ? 3 : 1;
              for (int j = 0; j < ntap; ++j, ++kw) {
                const int s = kw % WS;
                ptx::mbar_wait(&w_empty[s], ((kw / WS) & 1) ^ 1u);
                if (leader) ptx::mbar_arrive_expect_tx(&w_full[s], static_cast<uint32_t>(WB) * kMul);
                const int kb = kSharedA ? (g < 4 ? j * 4 + g : 12 + (g - 4)) : g;     // weight k-blocks are packed tap-major
                if constexpr (kPair) ptx::tma_load_2d_pair(sW + s * WB, mW1, ptx::mapa_u32(ptx::smem_u32(&w_full[s]), 0), kb * 64, half * 256 + nrow);
                else ptx::tma_load_2d(sW + s * WB, mW1, &w_full[s], kb * 64, half * 256);
              }
            }
          }
          for (int kb = 0; kb < 4; ++kb, ++kw) {          // residual GEMM weights (its A operand is the smem copy of u)
            const int s = kw % WS;
            ptx::mbar_wait(&w_empty[s], ((kw / WS) & 1) ^ 1u);
            if (leader) ptx::mbar_arrive_expect_tx(&w_full[s], static_cast<uint32_t>(WB) * kMul);
            if constexpr (kPair) ptx::tma_load_2d_pair(sW + s * WB, mW2, ptx::mapa_u32(ptx::smem_u32(&w_full[s]), 0), kb * 64, nrow);
            else ptx::tma_load_2d(sW + s * WB, mW2, &w_full[s], kb * 64, 0);
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer
      const uint32_t idesc = ptx::make_idesc_bf16_f32(kPair ? 2 * kTileM : kTileM, 256);
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
        if constexpr (kPair) ptx::mma_f16_ss_pair(d, da, db, idesc, acc); else ptx::mma_f16_ss(d, da, db, idesc, acc);
      };
      auto commit = [&](uint64_t* bar) {
        if constexpr (kPair) ptx::mma_commit_pair(bar); else ptx::mma_commit(bar);
      };
      if (leader)
      for (int unit = unit0; unit < total_units; unit += unit_stride, ++it) {
        const int tl = (unit - unit0) / unit_stride;   // local unit index in this layer
        long long* dm = (p.dbg && blockIdx.x == 0 && l == 3 && lane == 0 && tl < 2) ? p.dbg + 1 + tl * 8 : nullptr;
        for (int half = 0; half < 2; ++half) {
          const int job = 3 * it + half, buf = job & 1;
          ptx::mbar_wait(&acc_empty[buf], ((job >> 1) & 1) ^ 1u);
          ptx::tc_fence_after();
          if (dm) dm[half * 2] = clock64();                 // job may start (buffer free)
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * 256);
          uint32_t accum = 0;
          for (int g = 0; g < ngroups; ++g, ++ga) {
            const int slot = ga % AS;
            ptx::mbar_wait(&a_full[slot], (ga / AS) & 1);
            const int ntap = (kSharedA && g < nhb) ? 3 : 1;
            for (int j = 0; j < ntap; ++j, ++kw) {
              const int s = kw % WS;
              ptx::mbar_wait(&w_full[s], (kw / WS) & 1);
              ptx::tc_fence_after();
              if (lane == 0) {
                // hb tile holds frames t0-1 .. t0+128; tap j (offset j-1) starts at row j
                const uint32_t a_addr = ptx::smem_u32(sA + slot * AB) + ((kSharedA && g < nhb) ? static_cast<uint32_t>(j * 128) : 0u);
                const uint64_t da = ptx::make_desc_k_sw128(a_addr);
                const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(sW + s * WB));
#pragma unroll
                for (int k = 0; k < 4; ++k) mma(tmem_d, da + 2 * k, db + 2 * k, accum | (k != 0 ? 1u : 0u));
                commit(&w_empty[s]);
              }
              accum = 1;
              __syncwarp();
            }
            if (lane == 0) commit(&a_empty[slot]);
            __syncwarp();
          }
          if (lane == 0) commit(&acc_full[buf]);
          if (dm) dm[half * 2 + 1] = clock64();             // all MMAs of the job issued
          __syncwarp();
        }
        {
          const int job = 3 * it + 2, buf = job & 1;
          ptx::mbar_wait(&acc_empty[buf], ((job >> 1) & 1) ^ 1u);
          if (dm) dm[4] = clock64();
          ptx::mbar_wait(u_full, it & 1);                 // both halves of u are in shared memory
          ptx::tc_fence_after();
          if (dm) dm[5] = clock64();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * 256);
          for (int kb = 0; kb < 4; ++kb, ++kw) {
            const int s = kw % WS;
            ptx::mbar_wait(&w_full[s], (kw / WS) & 1);
            ptx::tc_fence_after();
            if (lane == 0) {
              const uint64_t da = ptx::make_desc_k_sw128(ptx::smem_u32(sU + kb * 16384));
              const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(sW + s * WB));
#pragma unroll
              for (int k = 0; k < 4; ++k) mma(tmem_d, da + 2 * k, db + 2 * k, (kb | k) != 0 ? 1u : 0u);
              commit(&w_empty[s]);
            }
            __syncwarp();
          }
          if (lane == 0) {
            commit(u_empty);
            commit(&acc_full[buf]);
          }
          if (dm) dm[6] = clock64();
          __syncwarp();
        }
      }
    } else {
      // ---------------------------------------------------------------- epilogue warps
      const int ew = warp - 2;
      const int q = warp & 3;
      const int half2 = ew >> 2;
      const float* b2 = p.b2 + static_cast<size_t>(l) * kFC;
      __nv_bfloat16* hb_out = (l & 1) ? p.hb0 : p.hb1;
      // arrivals that the leader's MMA warp waits for: local barrier, or (pair mode) the leader's barrier through the cluster window
      auto arrive_leader = [&](uint64_t* bar) {
        if constexpr (kPair) ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(bar), 0)); else ptx::mbar_arrive(bar);
      };
      for (int unit = unit0; unit < total_units; unit += unit_stride, ++it) {
        const int tile = kPair ? 2 * unit + static_cast<int>(rank) : unit;
        const int b = p.b_off + tile / tiles_per_item, t0 = (tile % tiles_per_item) * kTileM;
        const int r = q * 32 + lane;                    // row inside the tile = TMEM lane
        const int t = tile < total_tiles ? t0 + r : p.T;   // a dummy tile (odd tile count in pair mode) has no valid row
        const bool row_ok = t < p.T;
        const size_t row = static_cast<size_t>(b) * p.T + t;
        const float* db = p.dbias + static_cast<size_t>(b) * p.dbias_bstride + static_cast<size_t>(l) * 3 * 512;
        const bool e0 = t < 1, e2 = t >= p.T - 1;       // dilation 1: the taps that fell on the zero padding
        const int tl = (unit - unit0) / unit_stride;
        long long* de = (p.dbg && blockIdx.x == 0 && l == 3 && ew == 0 && lane == 0 && tl < 2) ? p.dbg + 20 + tl * 8 : nullptr;
        if (it > 0) ptx::mbar_wait(u_empty, (it - 1) & 1);   // the previous tile's residual GEMM has finished reading u
        for (int half = 0; half < 2; ++half) {
          const int job = 3 * it + half, buf = job & 1;
          ptx::mbar_wait(&acc_full[buf], (job >> 1) & 1);
          ptx::tc_fence_after();
          if (de) de[half * 2] = clock64();                 // accumulator ready
          const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * 256);
          for (int ci = 0; ci < 4; ++ci) {
            const int c = half2 * 4 + ci;                 // contiguous ownership: this warp writes u k-block (2*half + half2) only
            uint32_t rr[32];
            ptx::tmem_ld_32x32b_x32(lane_base + c * 32, rr);
            ptx::tmem_wait_ld();
            const int n0 = half * 256 + c * 32;           // first of 32 interleaved (gate, filter) columns
            const bool shared_tab = p.dbias_bstride == 0;
            const float* m = shared_tab ? sBias + n0 : db + n0;
            float y[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 mv = shared_tab ? *(reinterpret_cast<const float4*>(m) + i) : __ldg(reinterpret_cast<const float4*>(m) + i);
              y[4 * i] = __uint_as_float(rr[4 * i]) + mv.x; y[4 * i + 1] = __uint_as_float(rr[4 * i + 1]) + mv.y;
              y[4 * i + 2] = __uint_as_float(rr[4 * i + 2]) + mv.z; y[4 * i + 3] = __uint_as_float(rr[4 * i + 3]) + mv.w;
            }
            if (e0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) y[i] -= m[512 + i];
            }
            if (e2) {
#pragma unroll
              for (int i = 0; i < 32; ++i) y[i] -= m[1024 + i];
            }
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              pk[j] = pack_bf16x2(sigmoid_f<true>(y[4 * j]) * tanh_f<true>(y[4 * j + 1]),
                                  sigmoid_f<true>(y[4 * j + 2]) * tanh_f<true>(y[4 * j + 3]));
            // u columns [ucol, ucol+16): HBM copy for the folded skip GEMM ...
            const int ucol = half * 128 + c * 16;
            if (row_ok) {
              __nv_bfloat16* up = p.u_all + row * static_cast<size_t>(p.L * kFC) + l * kFC + ucol;
              asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(up), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
              asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(up + 8), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
            }
            // ... and the A operand of the residual GEMM: K-major, 128-byte rows, 16-byte chunks XOR-swizzled by (row & 7)
            uint8_t* ub = sU + (ucol >> 6) * 16384 + r * 128;
            const int ch = (ucol & 63) >> 3;               // first of the two 16-byte chunks
            *reinterpret_cast<uint4*>(ub + (((ch) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(ub + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          ptx::tc_fence_before();
          ptx::fence_proxy_async_smem();                  // generic-proxy smem writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) {
            arrive_leader(&acc_empty[buf]);
            arrive_leader(u_full);
          }
          if (de) de[half * 2 + 1] = clock64();             // gate epilogue of this half done
        }
        {
          // residual epilogue: h <- (h + o + b) / sqrt(2).
          // The fp32 residual stream is accessed TRANSPOSED: the 32x32 accumulator chunk goes through a 4 KB
          // XOR-swizzled scratch (the rows of the u tile that only this warp writes; u is dead between the residual
          // GEMM and the next tile's gate epilogue), so that 8 lanes cover one 128-byte row segment: every global
          // load/store instruction touches 4 cache lines instead of 32 (the lane-per-row form was LSU-bound).
          const int job = 3 * it + 2, buf = job & 1;
          float* stg = reinterpret_cast<float*>(sU + half2 * 16384 + q * 4096);
          const int cq = lane & 7, r0 = lane >> 3;
          const int tq = tile < total_tiles ? t0 + q * 32 + r0 : p.T;   // frame of iteration 0; iteration i adds 4*i
          const size_t rowq = static_cast<size_t>(b) * p.T + tq;
          float* hq = p.h + rowq * kFC + cq * 4;
          __nv_bfloat16* hbq = hb_out + rowq * kFC + cq * 4;
          float4 hv[8], hn[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            hv[i] = (tq + 4 * i < p.T) ? ld_stream_f4(hq + static_cast<size_t>(4 * i) * kFC + half2 * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
          ptx::mbar_wait(&acc_full[buf], (job >> 1) & 1);
          ptx::tc_fence_after();
          if (de) de[4] = clock64();
          const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * 256);
          for (int ci = 0; ci < 4; ++ci) {
            const int c = half2 * 4 + ci;                       // this warp owns columns [half2*128, half2*128+128)
            uint32_t rr[32];
            ptx::tmem_ld_32x32b_x32(lane_base + c * 32, rr);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_uint4(rr[4 * j], rr[4 * j + 1], rr[4 * j + 2], rr[4 * j + 3]);
            __syncwarp();
            if (ci + 1 < 4) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                hn[i] = (tq + 4 * i < p.T) ? ld_stream_f4(hq + static_cast<size_t>(4 * i) * kFC + (c + 1) * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float4 bv = *reinterpret_cast<const float4*>(sBias + 3 * 512 + c * 32 + cq * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rloc = 4 * i + r0;
              const float4 a = *reinterpret_cast<const float4*>(stg + rloc * 32 + ((cq ^ (rloc & 7)) << 2));
              if (tq + 4 * i < p.T) {
                float v[4];
                v[0] = (hv[i].x + (a.x + bv.x)) * 0.70710678118654752440f;
                v[1] = (hv[i].y + (a.y + bv.y)) * 0.70710678118654752440f;
                v[2] = (hv[i].z + (a.z + bv.z)) * 0.70710678118654752440f;
                v[3] = (hv[i].w + (a.w + bv.w)) * 0.70710678118654752440f;
                st_vec<4>(hq + static_cast<size_t>(4 * i) * kFC + c * 32, v);
                st_vec<4>(hbq + static_cast<size_t>(4 * i) * kFC + c * 32, v);
              }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) hv[i] = hn[i];
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader(&acc_empty[buf]);
          if (de) de[5] = clock64();
        }
      }
    }
    // ------------------------------------------------------------------ layer boundary
    // hb_{l+1} rows written by neighbour CTAs are read by the next layer's TMA loads (conv halo): grid-wide barrier.
    if (l + 1 < p.L) {
      if (p.dbg && blockIdx.x == 0 && l == 2 && threadIdx.x == 64) p.dbg[0] = clock64();    // layer 3 starts after this barrier
      __syncthreads();
      if (p.dbg && blockIdx.x == 0 && l == 3 && threadIdx.x == 0) p.dbg[40] = clock64();
      if (threadIdx.x == 0) fused_grid_barrier(p.grid_bar, static_cast<unsigned int>(l + 1) * gridDim.x);
      if (p.dbg && blockIdx.x == 0 && (l == 3 || l == 2) && threadIdx.x == 0) p.dbg[41 + (l == 2)] = clock64();
      __syncwarp();
      __syncthreads();
      asm volatile("fence.proxy.async;" ::: "memory");      // other CTAs' generic-proxy stores -> our TMA (async proxy) loads
    }
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __syncthreads();
  if constexpr (kPair) ptx::cluster_sync_all();     // the peer may still be reading our smem / signalling our barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (kPair) ptx::tmem_dealloc_pair(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fse
