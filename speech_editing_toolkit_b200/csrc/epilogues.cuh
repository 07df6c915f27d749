// Epilogue functors of the FluentSpeech denoiser step (one per conv_gemm launch).
// Each gets NV consecutive output columns [n0, n0+NV) of frame (b, t) as fp32 accumulators.
// TOp is the operand type the NEXT GEMM reads (bf16 for tensor cores, float for the exact mode).
// Reference semantics: modules/speech_editing/spec_denoiser/diffnet.py:68-81,110-132 and
// spec_denoiser.py:86-108 (file:line relative to the reference tree).
#pragma once
#include "conv_gemm.cuh"

namespace fse {

// ------------------------------------------------------------------ scalar math
template <bool Fast>
__device__ __forceinline__ float tanh_f(float x) {
  if constexpr (Fast) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  } else {
    return tanhf(x);
  }
}
template <bool Fast>
__device__ __forceinline__ float sigmoid_f(float x) {
  if constexpr (Fast) return fmaf(0.5f, tanh_f<true>(0.5f * x), 0.5f);
  else return 1.0f / (1.0f + expf(-x));
}

// Philox4x32-10 (counter-based RNG) + Box-Muller: 4 standard normals per call.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t step, uint32_t row, uint32_t grp, float (&z)[4]) {
  uint32_t r[4];
  philox4x32_10(row, grp, step, 0x46534542u, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  const float u0 = (static_cast<float>(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u1 = (static_cast<float>(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = (static_cast<float>(r[2] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u3 = (static_cast<float>(r[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  // fast intrinsics: the uniforms are 24-bit, so |log| and the phase are well inside the accurate range of the SFU paths
  const float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
  float s, c;
  __sincosf(6.283185307179586f * u1, &s, &c);
  z[0] = ra * c; z[1] = ra * s;
  __sincosf(6.283185307179586f * u3, &s, &c);
  z[2] = rb * c; z[3] = rb * s;
}

// Streaming (evict-first) stores for write-once data that is not re-read soon (the per-layer gate outputs),
// so they do not push the fp32 residual stream out of L2.
template <int NV>
__device__ __forceinline__ void st_vec_stream(float* p, const float* v) { st_vec<NV>(p, v); }
template <int NV>
__device__ __forceinline__ void st_vec_stream(__nv_bfloat16* p, const float* v) {
  if constexpr (NV % 8 == 0) {
#pragma unroll
    for (int i = 0; i < NV / 8; ++i)
      asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p + 8 * i), "r"(pack_bf16x2(v[8 * i], v[8 * i + 1])),
                   "r"(pack_bf16x2(v[8 * i + 2], v[8 * i + 3])), "r"(pack_bf16x2(v[8 * i + 4], v[8 * i + 5])),
                   "r"(pack_bf16x2(v[8 * i + 6], v[8 * i + 7]))
                   : "memory");
  } else {
    st_vec<NV>(p, v);
  }
}

// ------------------------------------------------------------------ input projection
// h = relu(W_in x + b_in)  (diffnet.py:117-120); writes the fp32 residual stream and its operand copy.
template <typename TOp>
struct EpiIn {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  const float* bias;   // [C]
  float* h;            // [B*T, C] fp32 residual stream
  TOp* hb;             // [B*T, C] operand copy
  int C, T;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    const size_t o = (static_cast<size_t>(b) * T + t) * C + n0;
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = fmaxf(acc[i] + __ldg(bias + n0 + i), 0.f);
    if (h) st_vec<NV>(h + o, v);
    st_vec<NV>(hb + o, v);
  }
};

// ------------------------------------------------------------------ gated activation
// y = conv_k3(h + d) + conv_1x1(cond) (+ biases); u = sigmoid(gate) * tanh(filter)  (diffnet.py:69-77).
// Columns are interleaved at weight-pack time: column 2j = gate channel j, 2j+1 = filter channel j.
// The timestep shift d enters as an exact fp32 bias  m - [t<dil] a - [t>=T-dil] c  with
// m = (W0+W1+W2) d + b_dc + b_cp,  a = W0 d,  c = W2 d  (zero padding is applied AFTER adding d in the
// reference, so the taps that fall outside [0,T) must not see d).
template <typename TOp, bool Fast>
struct EpiGate {
  static constexpr int kAux = 0;
  const float* dbias;       // [.., 3, N] for this layer; row 0 = m, 1 = a, 2 = c
  long long bstride;        // elements between consecutive batch items' tables (0: shared by the batch)
  TOp* u;                   // [B*T, ldu] rows; this layer's N/2 channels start at column u_off
  int N, T, dil;
  int ldu, u_off;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    const float* m = dbias + static_cast<size_t>(b) * bstride + n0;
    const bool e0 = t < dil, e2 = t >= T - dil;
    float y[NV];
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      const float4 mv = __ldg(reinterpret_cast<const float4*>(m) + i);
      y[4 * i] = acc[4 * i] + mv.x; y[4 * i + 1] = acc[4 * i + 1] + mv.y;
      y[4 * i + 2] = acc[4 * i + 2] + mv.z; y[4 * i + 3] = acc[4 * i + 3] + mv.w;
    }
    if (e0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) y[i] -= __ldg(m + N + i);
    }
    if (e2) {
#pragma unroll
      for (int i = 0; i < NV; ++i) y[i] -= __ldg(m + 2 * N + i);
    }
    float v[NV / 2];
#pragma unroll
    for (int j = 0; j < NV / 2; ++j) v[j] = sigmoid_f<Fast>(y[2 * j]) * tanh_f<Fast>(y[2 * j + 1]);
    st_vec_stream<NV / 2>(u + (static_cast<size_t>(b) * T + t) * ldu + u_off + n0 / 2, v);
  }
};

// ------------------------------------------------------------------ residual
// o_res = W_op[:C] u + b_op[:C];  h <- (h + o_res) / sqrt(2)   (diffnet.py:79-81).
// The skip half of output_projection never materialises: sum_l skip_l / sqrt(L) followed by skip_projection
// (diffnet.py:126-129) is linear in the per-layer gate outputs u_l, so it is evaluated at the end of the step
// as ONE GEMM over the concatenated u_l with the folded weight W_skip W_op,l[C:] / sqrt(L) (EpiSkip below).
template <typename TOp, bool Fast>
struct EpiRes {
  static constexpr int kAux = 1;          // the fp32 residual stream h
  const float* bias;   // [C]
  float* h;            // [B*T, C]
  TOp* hb;             // [B*T, C] operand copy
  int C, T;
  template <int NV>
  __device__ __forceinline__ void load_aux(int b, int t, int n0, float* aux) const {
    const float4* hp = reinterpret_cast<const float4*>(h + (static_cast<size_t>(b) * T + t) * C + n0);
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      const float4 v = hp[i];
      aux[4 * i] = v.x; aux[4 * i + 1] = v.y; aux[4 * i + 2] = v.z; aux[4 * i + 3] = v.w;
    }
  }
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float* aux) const {
    const size_t o = (static_cast<size_t>(b) * T + t) * C + n0;
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n0) + i);
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float sum = aux[4 * i + q] + (acc[4 * i + q] + bb[q]);
        if constexpr (Fast) v[4 * i + q] = sum * 0.70710678118654752440f;
        else v[4 * i + q] = __fdiv_rn(sum, 1.41421356237309504880f);   // torch: (x + residual) / sqrt(2.0)
      }
    }
    st_vec<NV>(h + o, v);
    st_vec<NV>(hb + o, v);
  }
};

// ------------------------------------------------------------------ skip projection
// r = relu(W_skip s + b_skip), s = sum_l skip_l / sqrt(L)  (diffnet.py:126-130), with the sum folded into the
// K axis of this GEMM (A = [u_0 | u_1 | ... | u_{L-1}], W = folded weight, bias = folded bias).
template <typename TOp>
struct EpiSkip {
  static constexpr int kAux = 0;
  static constexpr bool kTransposed = true;
  const float* bias;
  TOp* rb;   // [B*T, C]
  int C, T;
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    float v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = fmaxf(acc[i] + __ldg(bias + n0 + i), 0.f);
    st_vec<NV>(rb + (static_cast<size_t>(b) * T + t) * C + n0, v);
  }
};

// ------------------------------------------------------------------ output projection + posterior sample
// x0 = W_out r + b_out (diffnet.py:131); mode 0 stores x0 only (the DiffNet.forward contract);
// mode 1 fuses q_posterior_sample (spec_denoiser.py:95-101):
//   x_{t-1} = c1 x0 + c2 x_t + sigma z,  sigma = [t != 0] exp(0.5 logvar_clipped[t])
// written as x_out[B,M,T] (reference layout) + operand copy xb[B*T, M] for the next step, and on the
// final step mel_out[B,T,M] (= x[:,0].transpose(1,2), spec_denoiser.py:183) optionally composited with
// the reference mel: mel*mask + ref*(1-mask) (tasks/speech_editing/spec_denoiser.py:53).
template <typename TOp>
struct EpiOut {
  static constexpr int kAux = 0;
  const float* bias;     // [M]
  int M, T;
  int mode;
  const float* x_t;      // [B, M, T]
  float* x_out;          // [B, M, T]  (mode 0: receives x0)
  TOp* xb;               // [B*T, M] or null
  const float* noise;    // [B, M, T] or null -> Philox
  const unsigned long long* seedp;   // Philox seed in device memory (so a captured CUDA graph of the sampling loop is seed-independent)
  unsigned step;
  float c1, c2, sigma;
  float* mel_out;        // [B, T, M] or null
  const float* ref;      // [B, T, M] or null
  const float* mask;     // [B, T] or null
  template <int NV>
  __device__ __forceinline__ void apply(int b, int t, int n0, const float* acc, const float*) const {
    if (n0 >= M) return;
    const size_t row = static_cast<size_t>(b) * T + t;
    float v[NV];
    // All global reads first: x_t / noise may alias x_out as far as the compiler knows, so a load placed after a store of
    // the previous element is not hoisted and every element would pay its own memory round trip.
    float xt[NV], zn[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      xt[i] = 0.f; zn[i] = 0.f;
      if (mode == 1 && n0 + i < M) {
        const size_t idx = (static_cast<size_t>(b) * M + n0 + i) * T + t;
        xt[i] = __ldg(x_t + idx);
        if (noise) zn[i] = __ldg(noise + idx);
      }
    }
#pragma unroll
    for (int g = 0; g < NV / 4; ++g) {
      float z[4] = {0.f, 0.f, 0.f, 0.f};
      if (mode == 1 && noise == nullptr && sigma != 0.f)
        philox_normal4(__ldg(seedp), step, static_cast<uint32_t>(row), static_cast<uint32_t>((n0 >> 2) + g), z);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = 4 * g + q;
        const int m = n0 + i;
        float val = 0.f;
        if (m < M) {
          const float x0 = acc[i] + __ldg(bias + m);
          if (mode == 0) {
            val = x0;
          } else {
            const float zz = noise ? zn[i] : z[q];
            const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xt[i]));
            val = __fadd_rn(mean, __fmul_rn(sigma, zz));
          }
        }
        v[i] = val;
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (n0 + i < M) x_out[(static_cast<size_t>(b) * M + n0 + i) * T + t] = v[i];
    if (n0 + NV <= M) {
      if (xb) st_vec<NV>(xb + row * M + n0, v);
      if (mel_out) {
        if (mask) {
          const float mk = __ldg(mask + row);
#pragma unroll
          for (int i = 0; i < NV; ++i)
            v[i] = __fadd_rn(__fmul_rn(v[i], mk), __fmul_rn(__ldg(ref + row * M + n0 + i), 1.0f - mk));
        }
        st_vec<NV>(mel_out + row * M + n0, v);
      }
    } else {
      for (int i = 0; i < NV && n0 + i < M; ++i) {
        if (xb) {
          if constexpr (sizeof(TOp) == 2) xb[row * M + n0 + i] = __float2bfloat16_rn(v[i]);
          else xb[row * M + n0 + i] = v[i];
        }
        if (mel_out) {
          float o = v[i];
          if (mask) {
            const float mk = __ldg(mask + row);
            o = __fadd_rn(__fmul_rn(o, mk), __fmul_rn(__ldg(ref + row * M + n0 + i), 1.0f - mk));
          }
          mel_out[row * M + n0 + i] = o;
        }
      }
    }
  }
};

}  // namespace fse
