// Shared device/host helpers for the linnaeus_b200 sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/linnaeus_b200.h"

#define LNX_CHECK_LAUNCH()                                \
  do {                                                    \
    cudaError_t e__ = cudaGetLastError();                 \
    if (e__ != cudaSuccess) return lnx_set_cuda_error(e__); \
  } while (0)

#define LNX_REQUIRE(cond, code) \
  do {                          \
    if (!(cond)) return (code); \
  } while (0)

int lnx_set_cuda_error(cudaError_t e);  // records the CUDA error string, returns LNX_ERR_CUDA

static inline bool lnx_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

namespace lnx {

constexpr int kNumSMs = 148;

typedef __nv_bfloat16 bf16;

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector of T: 4 floats or 8 bf16
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ float get(int i) const { return (&raw.x)[i]; }
  __device__ __forceinline__ void set(int i, float v) { (&raw.x)[i] = v; }
};
template <>
struct Vec16<bf16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ float get(int i) const {
    const bf16* p = reinterpret_cast<const bf16*>(&raw);
    return __bfloat162float(p[i]);
  }
  __device__ __forceinline__ void set(int i, float v) {
    bf16* p = reinterpret_cast<bf16*>(&raw);
    p[i] = __float2bfloat16_rn(v);
  }
};

template <typename T>
__device__ __forceinline__ Vec16<T> ld16(const T* p) {
  Vec16<T> v;
  v.raw = *reinterpret_cast<const decltype(v.raw)*>(p);
  return v;
}
template <typename T>
__device__ __forceinline__ void st16(T* p, const Vec16<T>& v) {
  *reinterpret_cast<decltype(v.raw)*>(p) = v.raw;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// erf-GELU (nn.GELU default) and its derivative
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Fast erf-GELU for the bf16 tensor-core epilogues: Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7, far below
// bf16 resolution) with one MUFU.EX2 and one MUFU.RCP issued as raw approx instructions (the CUDA
// intrinsics add range-fixup code that triples the instruction count of the epilogue).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_fast_parts(float x, float& cdf, float& ez) {
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f));
  ez = ex2_approx(x * x * (-0.5f * 1.4426950408889634f));  // exp(-x^2/2)
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  const float q = poly * t * ez;            // 0.5 * erfc(|x|/sqrt2)
  cdf = x >= 0.f ? 1.0f - q : q;            // Phi(x)
}
__device__ __forceinline__ float gelu_fast(float x) {
  float cdf, ez;
  gelu_fast_parts(x, cdf, ez);
  return x * cdf;
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  float cdf, ez;
  gelu_fast_parts(x, cdf, ez);
  return fmaf(x * 0.39894228040143267794f, ez, cdf);
}

// Cheaper erf-GELU for the hot tcgen05 epilogues: Phi(x) = 0.5 (1 + tanh(x (c0 + c1 x^2 + c2 x^4))) with the odd
// polynomial fitted to erf (max |GELU error| 3e-5, max |GELU' error| 1e-4 -- far below bf16 resolution; the
// textbook 0.044715 form is 16x worse) and ONE MUFU.TANH per element; x^2 is clamped at 64 so the quintic
// never turns over.  The derivative is that of the same approximation, so forward and backward are consistent.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kGeluC0 = 7.97482750e-01f, kGeluC1 = 3.69853532e-02f, kGeluC2 = -3.46672663e-04f;
__device__ __forceinline__ float gelu_tanh3(float x) {
  const float x2 = fminf(x * x, 64.f);
  const float p = fmaf(x2, fmaf(x2, kGeluC2, kGeluC1), kGeluC0);
  const float t = tanh_approx(x * p);
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ float gelu_grad_tanh3(float x) {
  const float x2 = fminf(x * x, 64.f);
  const float p = fmaf(x2, fmaf(x2, kGeluC2, kGeluC1), kGeluC0);
  const float dp = fmaf(x2, fmaf(x2, 5.f * kGeluC2, 3.f * kGeluC1), kGeluC0);
  const float t = tanh_approx(x * p);
  const float s = fmaf(-t, t, 1.0f);
  return fmaf(0.5f * x * s, dp, fmaf(0.5f, t, 0.5f));
}

// Packed fp32x2 forms (FFMA2 / FMUL2, sm_100): the 3-register scalar FFMA issues at half rate on Blackwell, so the
// hot epilogues evaluate two elements per instruction.  gelu_grad folds the 0.5 and the sign into the derivative
// polynomial: g'(x) = 0.5 + 0.5 t + (x dpn(x^2)) (t^2 - 1), dpn = -0.5 d/dx[x p(x^2)].
__device__ __forceinline__ float2 f2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 gelu_tanh3_x2(float2 x) {
  float2 x2 = __fmul2_rn(x, x);
  x2.x = fminf(x2.x, 64.f);
  x2.y = fminf(x2.y, 64.f);
  const float2 p = __ffma2_rn(x2, __ffma2_rn(x2, f2(kGeluC2), f2(kGeluC1)), f2(kGeluC0));
  const float2 u = __fmul2_rn(x, p);
  const float2 t = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 hx = __fmul2_rn(x, f2(0.5f));
  return __ffma2_rn(hx, t, hx);
}
__device__ __forceinline__ float2 gelu_grad_tanh3_x2(float2 x) {
  float2 x2 = __fmul2_rn(x, x);
  x2.x = fminf(x2.x, 64.f);
  x2.y = fminf(x2.y, 64.f);
  const float2 p = __ffma2_rn(x2, __ffma2_rn(x2, f2(kGeluC2), f2(kGeluC1)), f2(kGeluC0));
  const float2 dpn = __ffma2_rn(x2, __ffma2_rn(x2, f2(-2.5f * kGeluC2), f2(-1.5f * kGeluC1)), f2(-0.5f * kGeluC0));
  const float2 u = __fmul2_rn(x, p);
  const float2 t = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 hn = __fmul2_rn(x, dpn);
  const float2 sm1 = __ffma2_rn(t, t, f2(-1.0f));
  return __ffma2_rn(hn, sm1, __ffma2_rn(t, f2(0.5f), f2(0.5f)));
}

// gelu and its derivative from ONE tanh (forward epilogue that saves the derivative for the backward pass)
__device__ __forceinline__ void gelu_both_tanh3_x2(float2 x, float2& gl, float2& dg) {
  float2 x2 = __fmul2_rn(x, x);
  x2.x = fminf(x2.x, 64.f);
  x2.y = fminf(x2.y, 64.f);
  const float2 p = __ffma2_rn(x2, __ffma2_rn(x2, f2(kGeluC2), f2(kGeluC1)), f2(kGeluC0));
  const float2 dpn = __ffma2_rn(x2, __ffma2_rn(x2, f2(-2.5f * kGeluC2), f2(-1.5f * kGeluC1)), f2(-0.5f * kGeluC0));
  const float2 u = __fmul2_rn(x, p);
  const float2 t = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 hx = __fmul2_rn(x, f2(0.5f));
  gl = __ffma2_rn(hx, t, hx);
  const float2 hn = __fmul2_rn(x, dpn);
  const float2 sm1 = __ffma2_rn(t, t, f2(-1.0f));
  dg = __ffma2_rn(hn, sm1, __ffma2_rn(t, f2(0.5f), f2(0.5f)));
}

// Leanest erf-GELU for the fused pointwise-pair kernels, where the GELU warps are the critical resource: cubic argument
// x (q0 + q1 x^2) (monotone, so no clamp; max |GELU error| 2.7e-4 before the bf16 rounding of the hidden tile, rms error
// after rounding 1.703e-3 against 1.694e-3 for exact erf-GELU) and one MUFU.TANH per element.  Measured on B200
// (tools/microbench): every MUFU op issues at 16 / clk / SM, FFMA2 at the same FMA rate as FFMA (124 FMA / clk / SM), and this
// body alone (bias add + gelu + pack, registers only) runs at 12.6 elements / clk / SM from 8 warps (1300 cycles per 128 x 128
// chunk); with the derivative it is FMA-pipe bound at 8.5 elements / clk / SM.  tanh.approx.f16x2 compiles to two MUFU.TANH.F16,
// so packing buys nothing.  gelu2q returns 2 gelu(x): the caller folds the
// 0.5 into the next scale.  gelu2q_both also returns the derivative of the same approximation (backward recompute).
constexpr float kGeluQ0 = 0.80015625f, kGeluQ1 = 0.034701171875f;
__device__ __forceinline__ float2 gelu2q_x2(float2 x) {  // 2 gelu(x)
  const float2 x2 = __fmul2_rn(x, x);
  const float2 u = __fmul2_rn(x, __ffma2_rn(x2, f2(kGeluQ1), f2(kGeluQ0)));
  return __ffma2_rn(x, make_float2(tanh_approx(u.x), tanh_approx(u.y)), x);
}
__device__ __forceinline__ void gelu2q_both_x2(float2 x, float2& g2, float2& dg) {  // 2 gelu(x) and gelu'(x)
  const float2 x2 = __fmul2_rn(x, x);
  const float2 u = __fmul2_rn(x, __ffma2_rn(x2, f2(kGeluQ1), f2(kGeluQ0)));
  const float2 t = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  g2 = __ffma2_rn(x, t, x);
  // g' = 0.5 (1 + t) + 0.5 x (1 - t^2) (q0 + 3 q1 x^2)
  const float2 xd = __fmul2_rn(x, __ffma2_rn(x2, f2(1.5f * kGeluQ1), f2(0.5f * kGeluQ0)));
  const float2 omt2 = __ffma2_rn(make_float2(-t.x, -t.y), t, f2(1.0f));
  dg = __ffma2_rn(xd, omt2, __ffma2_rn(t, f2(0.5f), f2(0.5f)));
}

// swish / SiLU = x sigmoid(x) = h (1 + tanh h), h = x / 2: one MUFU.TANH (bf16 tensor-core epilogues)
__device__ __forceinline__ float swish_fast(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace lnx
