// All residual layers of one DiffNet evaluation as ONE stream of (layer, unit) work items (diffnet.py:68-81 x L).
//
// denoiser_fused.cuh runs the layers in lock step: every CTA (pair) owns the same tiles in every layer and a grid-wide
// barrier separates the layers.  That leaves three holes per layer: the drain (the last tile's residual epilogue with an
// idle tensor pipe), the barrier itself, and the refill of the TMA/MMA pipeline; and with 128 units on 74 CTA pairs, 20
// pairs idle through every second half-layer.  Here the L x units items are numbered g = l * units + unit and dealt
// round-robin (item g belongs to pair g mod npairs), so
//   * every pair gets ceil/floor(L * units / npairs) items instead of L * ceil(units / npairs)  (35 instead of 40 at C2),
//   * the TMA producer, the MMA issuer and the epilogue warps each walk their item list without ever meeting at a
//     block- or grid-wide barrier: the tensor pipe starts item k+1 (usually of the next layer) while the epilogue warps
//     are still in the residual epilogue of item k.
// The only cross-CTA dependency is the k=3 conv halo plus the in-place residual stream: item (l, u) reads hb_l rows of
// units u-1, u, u+1 and h rows of unit u, all written by the residual epilogues of layer l-1.  A CTA publishes its rows
// with  stores -> mbarrier (all epilogue warps) -> publisher thread: __threadfence -> atomicAdd(done[l][u])  and the
// producer thread of a CTA acquires done[l-1][u-1 .. u+1] (== epilogue warps of the unit) followed by
// fence.proxy.async before it issues the item's TMA loads.  Dependencies point to strictly lower g, every pair works in increasing g and all CTAs are
// co-resident (grid <= SM count, 1 CTA/SM), so the item with the smallest unfinished g can always run: no deadlock.
// WAR on the hb ping-pong buffers: hb_{l+2} (same buffer as hb_l) is written by item (l+1, u), which waited for
// done[l][u-1 .. u+1], i.e. for every reader of hb_l rows of unit u.
//
// Per item the three tensor-core jobs, the shared-memory u tile and the two epilogues are those of denoiser_fused.cuh.
#pragma once
#include "denoiser_fused.cuh"

namespace fse {

// h rows may have been written by another SM in the previous layer: read them from L2
__device__ __forceinline__ float4 ld_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void stream_wait_done(const unsigned int* flag, unsigned int target) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
  if (v >= target) return;
  const long long t0 = clock64();
  do {
    __nanosleep(40);
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (clock64() - t0 > 4000000000LL) {
      printf("fse: layer dependency wait timed out (block %d, flag %p, target %u, seen %u)\n", blockIdx.x, flag, target, v);
      __trap();
    }
  } while (v < target);
}

constexpr int kStreamThreads = kTcThreads + 32;     // + the publisher warp

// kTF32 (CTA pairs only): the same schedule on tcgen05 kind::tf32 — the reference's own GPU arithmetic (cuDNN TF32 convolutions).
// Every operand is fp32 in HBM and shared memory, so a k-block is 32 channels (128-byte rows as before) and there are twice as
// many of them: 8 hb blocks + H/32 cond blocks per gate job, 8 u blocks for the residual GEMM.  The residual stream IS the
// operand (no bf16 copy): layer l reads hf[l & 1] (TMA halo tiles + the epilogue's own rows) and writes hf[(l + 1) & 1]; u is
// rounded to tf32 once (cvt.rna); u_all is fp32.  Gate non-linearities use ex2/rcp at fp32 accuracy instead of tanh.approx.
// Shared memory: a whole fp32 u tile would be 128 KB and leave room for only 2 activation slots + 3 weight stages; measured
// (FSE_DBG_STAMPS): the MMA warp then waits 60 % of the launch for operands (82 KB in flight against ~3 k cycles of loaded TMA
// latency).  So the tile holds ONE half of u (64 KB): the residual GEMM is split in two K halves, its first half (u of gate
// job a) is issued in the middle of gate job b (after the conv k-blocks, before the cond k-blocks: the first gate epilogue has
// long finished by then), and gate epilogue b overwrites the tile for the second half.  That restores the bf16 schedule's ring:
// 3 activation slots + 6 weight stages.
constexpr int kStreamTf32ASlots = 3, kStreamTf32WStages = 6;
constexpr int kStreamTf32UBytes = 4 * 128 * 128;
constexpr size_t kStreamTf32SmemBytes = 1024 + kStreamTf32ASlots * kFusedASlotBytes + kStreamTf32WStages * 128 * 128 + kStreamTf32UBytes + kFusedBiasBytes + 256;
static_assert(kStreamTf32SmemBytes <= 227 * 1024, "tf32 stream kernel: shared memory");

// sigmoid(g) * tanh(f) = (1 - E2) / ((1 + E1) (1 + E2)) with E1 = exp(-g), E2 = exp(-2 f): two ex2 and ONE reciprocal per element at
// fp32 accuracy (ex2.approx / rcp.approx are good to ~1 ulp); the arguments are clamped where both functions have long saturated in
// fp32 (|f| > 9 => tanh = +-1, sigmoid(-30) = 9e-14), which keeps every intermediate finite.
__device__ __forceinline__ float gate_acc(float g, float f) {
  g = fminf(fmaxf(g, -30.f), 30.f);
  f = fminf(fmaxf(f, -15.f), 15.f);
  float e1, e2;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(g * -1.4426950408889634f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(f * -2.8853900817779268f));
  return __fdividef(1.0f - e2, (1.0f + e1) * (1.0f + e2));
}

template <bool kPair, bool kTF32 = false>
__global__ void __launch_bounds__(kStreamThreads, 1)
denoiser_stream_kernel(const __grid_constant__ CUtensorMap mapHb0, const __grid_constant__ CUtensorMap mapHb1,
                       const __grid_constant__ CUtensorMap mapCond, const __grid_constant__ CUtensorMap mapU, FusedParams p) {
  static_assert(kPair || !kTF32, "the tf32 schedule exists for CTA pairs only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int WB = kPair ? 128 * 128 : 256 * 128;     // bytes of one weight stage (128 or 256 rows x 128 B)
  constexpr int AS = kTF32 ? kStreamTf32ASlots : kFusedASlots;     // 130-row activation tiles; taps = row-shifted descriptors
  constexpr int AB = kFusedASlotBytes;
  constexpr int WS = kTF32 ? kStreamTf32WStages : (kPair ? 6 : kFusedWStages);
  constexpr int KC = kTF32 ? 32 : 64;                   // channels per k-block (128 bytes)
  constexpr int NHB = kFC / KC;                         // hb channel blocks = u k-blocks: 4 (bf16) or 8 (tf32)
  constexpr int UB = kTF32 ? kStreamTf32UBytes : kFusedUBytes;
  static_assert(kTF32 || AS * AB + WS * WB <= kFusedASlots * kFusedASlotBytes + kFusedWStages * kFusedWStageBytes, "smem budget");
  constexpr uint32_t kMul = kPair ? 2u : 1u;
  uint8_t* sA = smem;
  uint8_t* sW = sA + AS * AB;
  uint8_t* sU = sW + WS * WB;
  float* sBias = reinterpret_cast<float*>(sU + UB);          // [3][512] timestep tables + [256] residual bias of the layer
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sU + UB + kFusedBiasBytes);
  uint64_t* a_empty = a_full + AS;
  uint64_t* w_full = a_empty + AS;
  uint64_t* w_empty = w_full + WS;
  uint64_t* acc_full = w_empty + WS;              // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint64_t* u_full = acc_empty + 2;               // [2]: u k-blocks 0,1 (first gate half) / 2,3 (second half) are in shared memory
  uint64_t* u_empty = u_full + 2;
  uint64_t* pub_bar = u_empty + 1;                // this CTA's epilogue warps have stored their h / hb rows of the item
  uint64_t* ust_full = pub_bar + 1;               // [2]: this CTA's half of the u tile is complete in shared memory -> TMA store to u_all
  uint64_t* ust_done = ust_full + 2;              // the TMA stores have read the u tile (bf16: once per item; tf32: once per half)
  uint64_t* ua_empty = ust_done + 1;              // tf32: the first half of the residual GEMM has read the u tile (gate epilogue b may overwrite it)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ua_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int tiles_per_item = (p.T + kTileM - 1) / kTileM;
  const int total_tiles = p.B * tiles_per_item;
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  // work units: a tile (or, in pair mode, two adjacent tiles 2u, 2u+1 handled by the CTAs of a cluster)
  const int pair0 = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int npairs = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int total_units = kPair ? (total_tiles + 1) / 2 : total_tiles;
  const int total_items = p.L * total_units;
  const unsigned int done_target = kEpiWarps * kMul;     // epilogue warps that store one unit
  const int nkbH = (p.H + KC - 1) / KC;
  const int ngroups = NHB + nkbH;                  // hb channel blocks (3 taps each) + cond blocks (1 tap)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&mapHb0);
    ptx::prefetch_tensormap(&mapHb1);
    ptx::prefetch_tensormap(&mapCond);
    ptx::prefetch_tensormap(&mapU);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < AS; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
      for (int i = 0; i < WS; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
      for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], kEpiWarps * kMul); }
      for (int i = 0; i < 2; ++i) ptx::mbar_init(&u_full[i], kEpiWarps * kMul);
      ptx::mbar_init(u_empty, 1);
      ptx::mbar_init(pub_bar, kEpiWarps);
      for (int i = 0; i < 2; ++i) ptx::mbar_init(&ust_full[i], kEpiWarps);
      ptx::mbar_init(ust_done, 1);
      ptx::mbar_init(ua_empty, 1);
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
  // the counters the NEXT launch will use (last touched by the previous launch, which completed before this kernel's predecessor did)
  if (p.done_clear)
    for (unsigned int i = blockIdx.x * kStreamThreads + threadIdx.x; i < p.done_clear_n; i += gridDim.x * kStreamThreads) p.done_clear[i] = 0u;
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[0] = clock64();
  constexpr int kDbgItem = 6;                       // items 6 and 7 of CTA 0 are stamped (developer aid)

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int ga = 0, kw = 0, it = 0;
      const int nrow = kPair ? static_cast<int>(rank) * 128 : 0;       // this CTA's half of every weight tile
      for (int g = pair0; g < total_items; g += npairs, ++it) {
        const int l = g / total_units, unit = g - l * total_units;
        const CUtensorMap* mHb = (l & 1) ? &mapHb1 : &mapHb0;
        const CUtensorMap* mW1 = (kPair ? p.mW1p : p.mW1) + l;
        const CUtensorMap* mW2 = (kPair ? p.mW2p : p.mW2) + l;
        const int tile = kPair ? 2 * unit + static_cast<int>(rank) : unit;
        const int b = p.b_off + tile / tiles_per_item, t0 = (tile % tiles_per_item) * kTileM;
        if (l > 0) {
          // layer l-1 of this unit and of its two neighbours (conv halo) has been published
          long long* dp = (p.dbg && blockIdx.x == 0 && it >= kDbgItem && it < kDbgItem + 2) ? p.dbg + 40 + (it - kDbgItem) * 2 : nullptr;
          if (dp) dp[0] = clock64();
          const unsigned int* f = p.done + static_cast<size_t>(l - 1) * total_units;
          const int u_lo = unit > 0 ? unit - 1 : 0, u_hi = unit + 1 < total_units ? unit + 1 : total_units - 1;
          for (int u = u_lo; u <= u_hi; ++u) stream_wait_done(f + u, done_target);
          asm volatile("fence.proxy.async;" ::: "memory");      // other CTAs' generic-proxy stores -> our TMA (async proxy) loads
          if (dp) dp[1] = clock64();
        }
        auto load_w2 = [&](int kb0, int kb1) {          // residual GEMM weights (its A operand is the smem copy of u)
          for (int kb = kb0; kb < kb1; ++kb, ++kw) {
            const int s = kw % WS;
            ptx::mbar_wait(&w_empty[s], ((kw / WS) & 1) ^ 1u);
            if (leader) ptx::mbar_arrive_expect_tx(&w_full[s], static_cast<uint32_t>(WB) * kMul);
            if constexpr (kPair) ptx::tma_load_2d_pair(sW + s * WB, mW2, ptx::mapa_u32(ptx::smem_u32(&w_full[s]), 0), kb * KC, nrow);
            else ptx::tma_load_2d(sW + s * WB, mW2, &w_full[s], kb * KC, 0);
          }
        };
        for (int half = 0; half < 2; ++half) {
          for (int grp = 0; grp < ngroups; ++grp, ++ga) {
            if constexpr (kTF32) { if (half == 1 && grp == NHB) load_w2(0, NHB / 2); }     // same order as the MMA warp consumes the ring
            const int slot = ga % AS;
            ptx::mbar_wait(&a_empty[slot], ((ga / AS) & 1) ^ 1u);
            // pair mode: both CTAs' loads signal the LEADER's barrier, which expects the bytes of both
            const bool is_hb = grp < NHB;
            const uint32_t rows = is_hb ? 130u : 128u;
            const int c0 = is_hb ? grp * KC : (grp - NHB) * KC;
            const int tt = is_hb ? t0 - 1 : t0;
            if (leader) ptx::mbar_arrive_expect_tx(&a_full[slot], rows * 128u * kMul);
            if constexpr (kPair) {
              ptx::tma_load_3d_pair(sA + slot * AB, is_hb ? mHb : &mapCond, ptx::mapa_u32(ptx::smem_u32(&a_full[slot]), 0), c0, tt, b);
            } else {
              ptx::tma_load_3d(sA + slot * AB, is_hb ? mHb : &mapCond, &a_full[slot], c0, tt, b);
            }
            const int ntap = is_hb ? 3 : 1;
            for (int j = 0; j < ntap; ++j, ++kw) {
              const int s = kw % WS;
              ptx::mbar_wait(&w_empty[s], ((kw / WS) & 1) ^ 1u);
              if (leader) ptx::mbar_arrive_expect_tx(&w_full[s], static_cast<uint32_t>(WB) * kMul);
              const int kb = is_hb ? j * NHB + grp : 3 * NHB + (grp - NHB);      // weight k-blocks are packed tap-major
              if constexpr (kPair) ptx::tma_load_2d_pair(sW + s * WB, mW1, ptx::mapa_u32(ptx::smem_u32(&w_full[s]), 0), kb * KC, half * 256 + nrow);
              else ptx::tma_load_2d(sW + s * WB, mW1, &w_full[s], kb * KC, half * 256);
            }
          }
        }
        load_w2(kTF32 ? NHB / 2 : 0, NHB);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = kTF32 ? ptx::make_idesc_tf32_f32(kPair ? 2 * kTileM : kTileM, 256) : ptx::make_idesc_bf16_f32(kPair ? 2 * kTileM : kTileM, 256);
    auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
      if constexpr (kTF32) ptx::mma_tf32_ss_pair(d, da, db, idesc, acc);
      else if constexpr (kPair) ptx::mma_f16_ss_pair(d, da, db, idesc, acc);
      else ptx::mma_f16_ss(d, da, db, idesc, acc);
    };
    auto commit = [&](uint64_t* bar) {
      if constexpr (kPair) ptx::mma_commit_pair(bar); else ptx::mma_commit(bar);
    };
    if (leader) {
      const bool el = ptx::elect_one();      // the one lane that issues every tcgen05.mma / commit of the pair
      int ga = 0, kw = 0, it = 0;
      // developer aid (FSE_DBG_STAMPS=1): cycles CTA 0's MMA warp spends waiting for accumulators / activation tiles / weight stages / u
      const bool acct = p.dbg && blockIdx.x == 0;
      long long w_acc = 0, w_a = 0, w_w = 0, w_u = 0;
      auto timed_wait = [&](uint64_t* bar, uint32_t parity, long long& sum) {
        if (acct) { const long long c0 = clock64(); ptx::mbar_wait(bar, parity); sum += clock64() - c0; }
        else ptx::mbar_wait(bar, parity);
      };
      for (int g = pair0; g < total_items; g += npairs, ++it) {
        long long* dm = (p.dbg && blockIdx.x == 0 && lane == 0 && it >= kDbgItem && it < kDbgItem + 2) ? p.dbg + 1 + (it - kDbgItem) * 8 : nullptr;
        // residual GEMM k-blocks [kb0, kb1): A = the u tile in shared memory (tf32: the tile holds one K half at a time)
        const int job2 = 3 * it + 2, buf2 = job2 & 1;
        const uint32_t tmem_d2 = tmem_base + static_cast<uint32_t>(buf2 * 256);
        auto res_gemm = [&](int kb0, int kb1) {
          for (int kb = kb0; kb < kb1; ++kb, ++kw) {
            if (kb % (NHB / 2) == 0) {
              // the first half of the residual GEMM (first half of the u k-blocks) only needs the FIRST gate epilogue, which
              // finished while the second gate job was running; only the second half waits for the second gate epilogue
              timed_wait(&u_full[kb / (NHB / 2)], it & 1, w_u);
              ptx::tc_fence_after();
              if (dm && kb == NHB / 2) dm[5] = clock64();
            }
            const int s = kw % WS;
            timed_wait(&w_full[s], (kw / WS) & 1, w_w);
            ptx::tc_fence_after();
            {
              const int ukb = kTF32 ? kb % (NHB / 2) : kb;
              const uint64_t da = ptx::make_desc_k_sw128(ptx::smem_u32(sU + ukb * 16384));
              const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(sW + s * WB));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (el) mma(tmem_d2, da + 2 * k, db + 2 * k, (kb | k) != 0 ? 1u : 0u);
              if (el) commit(&w_empty[s]);
            }
            __syncwarp();
          }
        };
        for (int half = 0; half < 2; ++half) {
          const int job = 3 * it + half, buf = job & 1;
          timed_wait(&acc_empty[buf], ((job >> 1) & 1) ^ 1u, w_acc);
          ptx::tc_fence_after();
          if (dm) dm[half * 2] = clock64();                 // job may start (buffer free)
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * 256);
          uint32_t accum = 0;
          for (int grp = 0; grp < ngroups; ++grp, ++ga) {
            if constexpr (kTF32) {
              if (half == 1 && grp == NHB) {
                // first K half of the residual GEMM, between the conv and the cond k-blocks of gate job b: its accumulator (the
                // buffer gate job a used) and the u half tile were both released by gate epilogue a several thousand cycles ago
                timed_wait(&acc_empty[buf2], ((job2 >> 1) & 1) ^ 1u, w_acc);
                ptx::tc_fence_after();
                res_gemm(0, NHB / 2);
                if (el) commit(ua_empty);                   // -> gate epilogue b may overwrite the tile (both CTAs)
                __syncwarp();
              }
            }
            const int slot = ga % AS;
            timed_wait(&a_full[slot], (ga / AS) & 1, w_a);
            const int ntap = grp < NHB ? 3 : 1;
            for (int j = 0; j < ntap; ++j, ++kw) {
              const int s = kw % WS;
              timed_wait(&w_full[s], (kw / WS) & 1, w_w);
              ptx::tc_fence_after();
              {
                // hb tile holds frames t0-1 .. t0+128; tap j (offset j-1) starts at row j.  Descriptors are computed by the
                // whole warp (convergent code -> uniform registers); only the tcgen05 instructions are under the one-lane
                // predicate (inside a divergent region every operand would need an R2UR and a waterfall loop per MMA).
                const uint32_t a_addr = ptx::smem_u32(sA + slot * AB) + (grp < NHB ? static_cast<uint32_t>(j * 128) : 0u);
                const uint64_t da = ptx::make_desc_k_sw128(a_addr);
                const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(sW + s * WB));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (el) mma(tmem_d, da + 2 * k, db + 2 * k, accum | (k != 0 ? 1u : 0u));
                if (el) commit(&w_empty[s]);
              }
              accum = 1;
              __syncwarp();
            }
            if (el) commit(&a_empty[slot]);
            __syncwarp();
          }
          if (el) commit(&acc_full[buf]);
          if (dm) dm[half * 2 + 1] = clock64();             // all MMAs of the job issued
          __syncwarp();
        }
        {
          if constexpr (!kTF32) {
            timed_wait(&acc_empty[buf2], ((job2 >> 1) & 1) ^ 1u, w_acc);
            ptx::tc_fence_after();
          }
          if (dm) dm[4] = clock64();
          res_gemm(kTF32 ? NHB / 2 : 0, NHB);
          if (el) {
            commit(u_empty);
            commit(&acc_full[buf2]);
          }
          if (dm) dm[6] = clock64();
          __syncwarp();
        }
      }
      if (acct && lane == 0) { p.dbg[48] = w_acc; p.dbg[49] = w_a; p.dbg[50] = w_w; p.dbg[51] = w_u; p.dbg[52] = clock64() - p.dbg[0]; p.dbg[53] = it; }
    }
  } else if (warp == 2 + kEpiWarps) {
    // ---------------------------------------------------------------- publisher
    // The gpu-scope fence that publishes an item's h / hb rows waits until the stores have reached L2 (~2 k cycles under
    // load).  The epilogue warps only arrive on a CTA-scope mbarrier; this otherwise idle thread takes the wait:
    //   epilogue stores -> mbarrier.arrive (release.cta)  =>  try_wait (acquire.cta) -> fence.acq_rel.gpu -> atomicAdd
    // (the cumulativity pattern of a grid sync: bar.sync, then ONE thread fences and signals for the whole CTA).
    // The consumers need the flag ~1.7 item periods later, and an item lasts >10x the fence, so the parity wait cannot lap.
    // The same thread moves the item's gate outputs u (needed later by the folded skip GEMM) from the shared-memory u tile to
    // u_all with TMA stores: the tile is already in the 128B-swizzled box layout of the tensor map, so the epilogue warps issue
    // no global store for it at all (they used to spend a quarter of the gate epilogue in the LSU queue on 16-byte row pieces).
    if (lane == 0) {
      const uint64_t pol = ptx::l2_policy_evict_first();          // write-once data, read once by the skip GEMM after the launch
      int it = 0;
      for (int g = pair0; g < total_items; g += npairs, ++it) {
        const int l = g / total_units, unit = g - l * total_units;
        const int tile = kPair ? 2 * unit + static_cast<int>(rank) : unit;
        const int b = p.b_off + tile / tiles_per_item, t0 = (tile % tiles_per_item) * kTileM;
        for (int half = 0; half < 2; ++half) {
          ptx::mbar_wait(&ust_full[half], it & 1);                // epilogue warps: st.shared + fence.proxy.async + arrive
          if (tile < total_tiles) {
            for (int j = 0; j < NHB / 2; ++j) {
              const int kb = half * (NHB / 2) + j;                // u k-block; tf32: the tile holds the current half only
              ptx::tma_store_3d(&mapU, sU + (kTF32 ? j : kb) * 16384, l * kFC + kb * KC, t0, b, pol);      // rows >= T are clipped
            }
          }
          if (kTF32 || half == 1) {
            ptx::bulk_commit_group();
            ptx::bulk_wait_group_read0();                         // the tile has been read: it may be overwritten / used as scratch
            ptx::mbar_arrive(ust_done);                           // tf32: phase 2 it (first half), 2 it + 1 (second half); bf16: phase it
          }
        }
        ptx::mbar_wait(pub_bar, it & 1);
        __threadfence();
        atomicAdd(p.done + g, static_cast<unsigned int>(kEpiWarps));     // g == l * total_units + unit
      }
      ptx::bulk_wait_group0();                                    // all u_all writes performed before the CTA retires
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- epilogue warps
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half2 = ew >> 2;
    // arrivals that the leader's MMA warp waits for: local barrier, or (pair mode) the leader's barrier through the cluster window
    auto arrive_leader = [&](uint64_t* bar) {
      if constexpr (kPair) ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(bar), 0)); else ptx::mbar_arrive(bar);
    };
    int it = 0, cur_l = -1;
    for (int g = pair0; g < total_items; g += npairs, ++it) {
      const int l = g / total_units, unit = g - l * total_units;
      if (l != cur_l) {
        // Per-layer bias tables -> shared memory (with per-item tables, dbias_bstride != 0, only the residual bias is staged).
        if (cur_l >= 0) asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");   // everyone is done with the old tables
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
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        cur_l = l;
      }
      __nv_bfloat16* hb_out = (l & 1) ? p.hb0 : p.hb1;
      // tf32: the fp32 residual stream ping-pongs (layer l reads hf[l & 1], writes hf[(l + 1) & 1]); bf16: h in place + bf16 copy
      const float* h_in = kTF32 ? ((l & 1) ? p.hf1 : p.hf0) : p.h;
      float* h_out = kTF32 ? ((l & 1) ? p.hf0 : p.hf1) : p.h;
      const int tile = kPair ? 2 * unit + static_cast<int>(rank) : unit;
      const int b = p.b_off + tile / tiles_per_item, t0 = (tile % tiles_per_item) * kTileM;
      const int r = q * 32 + lane;                    // row inside the tile = TMEM lane
      const int t = tile < total_tiles ? t0 + r : p.T;   // a dummy tile (odd tile count in pair mode) has no valid row
      const bool row_ok = t < p.T;
      const size_t row = static_cast<size_t>(b) * p.T + t;
      const float* db = p.dbias + static_cast<size_t>(b) * p.dbias_bstride + static_cast<size_t>(l) * 3 * 512;
      const bool e0 = t < 1, e2 = t >= p.T - 1;       // dilation 1: the taps that fell on the zero padding
      long long* de = (p.dbg && blockIdx.x == 0 && ew == 0 && lane == 0 && it >= kDbgItem && it < kDbgItem + 2) ? p.dbg + 20 + (it - kDbgItem) * 8 : nullptr;
      if (it > 0) ptx::mbar_wait(u_empty, (it - 1) & 1);   // the previous item's residual GEMM has finished reading u
      for (int half = 0; half < 2; ++half) {
        const int job = 3 * it + half, buf = job & 1;
        ptx::mbar_wait(&acc_full[buf], (job >> 1) & 1);
        ptx::tc_fence_after();
        if constexpr (kTF32) {
          if (half == 1) {          // the half tile is rewritten: the residual GEMM's first K half and the TMA store have read it
            ptx::mbar_wait(ua_empty, it & 1);
            ptx::mbar_wait(ust_done, 0);
          }
        }
        if (de) de[half * 2] = clock64();                 // accumulator ready
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * 256);
        // accumulator chunks are fetched one ahead: the tcgen05.ld of chunk c+1 is in flight while chunk c is processed
        uint32_t ra[32], rb[32];
        ptx::tmem_ld_32x32b_x32(lane_base + half2 * 128, ra);
        ptx::tmem_wait_ld();
        auto gate_chunk = [&](int ci, const uint32_t (&rr)[32]) {
          const int c = half2 * 4 + ci;                 // contiguous ownership: this warp writes u k-block (2*half + half2) only
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
          const int ucol = half * 128 + c * 16;
          if constexpr (kTF32) {
            uint32_t uf[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) uf[j] = __float_as_uint(ptx::round_tf32(gate_acc(y[2 * j], y[2 * j + 1])));
            // A operand of the residual GEMM: k-block = 32 fp32 channels, 128-byte rows, 16-byte chunks XOR-swizzled by (row & 7);
            // the tile holds this half's 128 u channels (4 k-blocks)
            uint8_t* ub = sU + ((ucol & 127) >> 5) * 16384 + r * 128;
            const int ch = (ucol & 31) >> 2;             // first of the four 16-byte chunks
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(ub + (((ch + j) ^ (r & 7)) << 4)) = make_uint4(uf[4 * j], uf[4 * j + 1], uf[4 * j + 2], uf[4 * j + 3]);
            return;
          }
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            pk[j] = pack_bf16x2(sigmoid_f<true>(y[4 * j]) * tanh_f<true>(y[4 * j + 1]),
                                sigmoid_f<true>(y[4 * j + 2]) * tanh_f<true>(y[4 * j + 3]));
          // u columns [ucol, ucol+16) -> the A operand of the residual GEMM (and, from there, to u_all by TMA store: publisher warp): K-major, 128-byte rows, 16-byte chunks XOR-swizzled by (row & 7)
          uint8_t* ub = sU + (ucol >> 6) * 16384 + r * 128;
          const int ch = (ucol & 63) >> 3;               // first of the two 16-byte chunks
          *reinterpret_cast<uint4*>(ub + (((ch) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(ub + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        };
        ptx::tmem_ld_32x32b_x32(lane_base + half2 * 128 + 32, rb);
        gate_chunk(0, ra);
        ptx::tmem_wait_ld();
        ptx::tmem_ld_32x32b_x32(lane_base + half2 * 128 + 64, ra);
        gate_chunk(1, rb);
        ptx::tmem_wait_ld();
        ptx::tmem_ld_32x32b_x32(lane_base + half2 * 128 + 96, rb);
        gate_chunk(2, ra);
        ptx::tmem_wait_ld();
        gate_chunk(3, rb);
        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();                  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) {
          arrive_leader(&acc_empty[buf]);
          arrive_leader(&u_full[half]);
          ptx::mbar_arrive(&ust_full[half]);              // this CTA's own publisher thread stores the half tile to u_all
        }
        if (de) de[half * 2 + 1] = clock64();             // gate epilogue of this half done
      }
      {
        // residual epilogue: h <- (h + o + b) / sqrt(2), accessed TRANSPOSED through 4 KB XOR-swizzled scratches (rows of
        // the u tile that only this warp writes; u is dead between the residual GEMM and the next item's gate epilogue):
        // 8 lanes cover one 128-byte row segment, so a global load/store instruction touches 4 cache lines instead of 32.
        // Two scratches (this warp's rows of u k-blocks half2 and 2 + half2) let the accumulator leave TMEM two chunks
        // ahead of the global traffic: the buffer goes back to the MMA warp after ~40 % of this epilogue (the next item's
        // second gate job was waiting for exactly that).
        const int job = 3 * it + 2, buf = job & 1;
        // (tf32: this warp wrote rows q*32.. of k-blocks 2*half2 and 2*half2 + 1 of the half tile)
        float* stgA = reinterpret_cast<float*>(sU + (kTF32 ? 2 * half2 : half2) * 16384 + q * 4096);
        float* stgB = reinterpret_cast<float*>(sU + (kTF32 ? 2 * half2 + 1 : 2 + half2) * 16384 + q * 4096);
        const int cq = lane & 7, r0 = lane >> 3;
        const int tq = tile < total_tiles ? t0 + q * 32 + r0 : p.T;   // frame of iteration 0; iteration i adds 4*i
        const size_t rowq = static_cast<size_t>(b) * p.T + tq;
        const float* hq = h_in + rowq * kFC + cq * 4;
        float* hoq = h_out + rowq * kFC + cq * 4;
        __nv_bfloat16* hbq = hb_out + rowq * kFC + cq * 4;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * 256);
        auto load_h = [&](int c, float4 (&hx)[8]) {                 // this lane's part of residual-stream chunk c
#pragma unroll
          for (int i = 0; i < 8; ++i)
            hx[i] = (tq + 4 * i < p.T) ? ld_cg_f4(hq + static_cast<size_t>(4 * i) * kFC + c * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto stage = [&](int c, float* stg) {                       // accumulator chunk c: TMEM -> registers -> scratch
          uint32_t rr[32];
          ptx::tmem_ld_32x32b_x32(lane_base + c * 32, rr);
          ptx::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_uint4(rr[4 * j], rr[4 * j + 1], rr[4 * j + 2], rr[4 * j + 3]);
          __syncwarp();
        };
        auto finish = [&](int c, const float* stg, const float4 (&hx)[8]) {
          const float4 bv = *reinterpret_cast<const float4*>(sBias + 3 * 512 + c * 32 + cq * 4);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rloc = 4 * i + r0;
            const float4 a = *reinterpret_cast<const float4*>(stg + rloc * 32 + ((cq ^ (rloc & 7)) << 2));
            if (tq + 4 * i < p.T) {
              float v[4];
              v[0] = (hx[i].x + (a.x + bv.x)) * 0.70710678118654752440f;
              v[1] = (hx[i].y + (a.y + bv.y)) * 0.70710678118654752440f;
              v[2] = (hx[i].z + (a.z + bv.z)) * 0.70710678118654752440f;
              v[3] = (hx[i].w + (a.w + bv.w)) * 0.70710678118654752440f;
              st_vec<4>(hoq + static_cast<size_t>(4 * i) * kFC + c * 32, v);
              if constexpr (!kTF32) st_vec<4>(hbq + static_cast<size_t>(4 * i) * kFC + c * 32, v);
            }
          }
          __syncwarp();                                             // all lanes have read the scratch
        };
        const int cb = half2 * 4;                                   // this warp owns columns [half2*128, half2*128+128)
        // The residual-stream loads are what this epilogue waits for (L2 latency under the weight traffic): three register
        // sets keep two chunks in flight, and the first two are issued before the wait for the residual GEMM.
        float4 h0[8], h1[8], h2[8];
        load_h(cb, h0);
        load_h(cb + 1, h1);
        ptx::mbar_wait(&acc_full[buf], (job >> 1) & 1);
        ptx::tc_fence_after();
        ptx::mbar_wait(ust_done, kTF32 ? 1u : static_cast<uint32_t>(it & 1));   // the TMA stores to u_all have read the u tile: it is scratch now
        if (de) de[4] = clock64();
        {
        stage(cb, stgA);
        stage(cb + 1, stgB);
        load_h(cb + 2, h2);
        finish(cb, stgA, h0);
        stage(cb + 2, stgA);
        load_h(cb + 3, h0);
        finish(cb + 1, stgB, h1);
        stage(cb + 3, stgB);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(&acc_empty[buf]);              // accumulator drained: early release
        if (de) de[6] = clock64();
        finish(cb + 2, stgA, h2);
        finish(cb + 3, stgB, h0);
        }
        if (lane == 0) ptx::mbar_arrive(pub_bar);                   // (finish ends with __syncwarp) -> publisher warp
        if (de) de[5] = clock64();
      }
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
