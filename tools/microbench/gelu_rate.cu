// How fast can W warps of one SM run the fused-MLP GELU epilogue body (bias add, cubic-argument tanh GELU, bf16 pack) when the data
// already sits in registers?  Separates the instruction-mix limit from the TMEM / barrier overheads of the real kernel.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../linnaeus_b200/csrc -o gelu_rate gelu_rate.cu && ./gelu_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "lnx_common.cuh"
#include "lnx_tc_common.cuh"
#include "lnx_mlp_fused.cuh"

using namespace lnx;
using namespace lnx_mlp;

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ float x_dummy(const uint32_t* acc) { return __uint_as_float(acc[0] ^ acc[1]); }

// MODE 0: gelu only (forward), 32 values per batch; 1: 16 per batch; 2: gelu + derivative (backward), 32 per batch; 3: 16 per batch
template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k(uint32_t* out, int iters, long long* cycles, const float* bias_g) {
  __shared__ float bias[128];
  if (threadIdx.x < 128) bias[threadIdx.x] = bias_g[threadIdx.x];
  constexpr int NB = (MODE & 1) ? 16 : 32;
  uint32_t acc[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) acc[i] = __float_as_uint((threadIdx.x * 37 + i * 11) % 97 * 0.05f - 2.4f);
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const float* bp = bias + ((it * NB) & 96);
    uint32_t pk[NB / 2];
#pragma unroll
    for (int i = 0; i < NB / 2; i += 2) {
      const float4 bv = *reinterpret_cast<const float4*>(bp + 2 * i);
      const float2 x0 = __fadd2_rn(make_float2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1])), make_float2(bv.x, bv.y));
      const float2 x1 = __fadd2_rn(make_float2(__uint_as_float(acc[2 * i + 2]), __uint_as_float(acc[2 * i + 3])), make_float2(bv.z, bv.w));
      if (MODE < 2) {
        const float2 r0 = gelu2q_x2(x0), r1 = gelu2q_x2(x1);
        pk[i] = pack2(r0.x, r0.y);
        pk[i + 1] = pack2(r1.x, r1.y);
      } else {
        float2 g0, d0, g1, d1;
        gelu2q_both_x2(x0, g0, d0);
        gelu2q_both_x2(x1, g1, d1);
        d0 = __fmul2_rn(d0, x1);  // stands in for the dH multiply
        d1 = __fmul2_rn(d1, x0);
        pk[i] = pack2(g0.x, g0.y) ^ pack2(d0.x, d0.y);
        pk[i + 1] = pack2(g1.x, g1.y) ^ pack2(d1.x, d1.y);
      }
    }
#pragma unroll
    for (int i = 0; i < NB / 2; ++i) {
      sink ^= pk[i];
      acc[2 * i] += pk[i] & 0x00010000u;  // keeps the inputs loop-carried without changing their range much
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sink;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Lockstep variants: all warps meet at a named barrier once per "chunk" of two 32-column blocks, as the GELU warps of the fused kernel
// meet at the accumulator-ready mbarrier.  MODE 0: block after block (A0 B0 A1 B1); MODE 1: software pipelined -- the tanh half (B) of a
// block shares a scheduling region with the FMA half (A) of the next one.
__device__ __forceinline__ void half_a(const uint32_t* acc, const float* bp, float2* x, float2* u) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float4 bv = *reinterpret_cast<const float4*>(bp + 2 * i);
    x[i] = __fadd2_rn(make_float2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1])), make_float2(bv.x, bv.y));
    x[i + 1] = __fadd2_rn(make_float2(__uint_as_float(acc[2 * i + 2]), __uint_as_float(acc[2 * i + 3])), make_float2(bv.z, bv.w));
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) u[i] = __fmul2_rn(x[i], __ffma2_rn(__fmul2_rn(x[i], x[i]), f2(kGeluQ1), f2(kGeluQ0)));
}
__device__ __forceinline__ void half_b(const float2* x, const float2* u, uint32_t* pk) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 r = __ffma2_rn(x[i], make_float2(tanh_approx(u[i].x), tanh_approx(u[i].y)), x[i]);
    pk[i] = pack2(r.x, r.y);
  }
}
template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) klock(uint32_t* out, int iters, long long* cycles, const float* bias_g) {
  __shared__ float bias[128];
  if (threadIdx.x < 128) bias[threadIdx.x] = bias_g[threadIdx.x];
  uint32_t acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint((threadIdx.x * 37 + i * 11) % 97 * 0.05f - 2.4f);
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0) {
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      asm volatile("bar.sync 1, %0;" ::"r"(THREADS) : "memory");
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        float2 x[16], u[16];
        uint32_t pk[16];
        half_a(acc, bias + blk * 32, x, u);
        half_b(x, u, pk);
#pragma unroll
        for (int i = 0; i < 16; ++i) { sink ^= pk[i]; acc[2 * i] += pk[i] & 0x00010000u; }
      }
    }
  } else {
    float2 x[16], u[16];
    half_a(acc, bias, x, u);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      uint32_t pk[16];
      float2 xn[16], un[16];
      // block 0's tanh half with block 1's FMA half
      half_a(acc, bias + 32, xn, un);
      half_b(x, u, pk);
#pragma unroll
      for (int i = 0; i < 16; ++i) { sink ^= pk[i]; acc[2 * i] += pk[i] & 0x00010000u; }
      asm volatile("bar.sync 1, %0;" ::"r"(THREADS) : "memory");  // next chunk's accumulator ready
      half_a(acc, bias, x, u);
      half_b(xn, un, pk);
#pragma unroll
      for (int i = 0; i < 16; ++i) { sink ^= pk[i]; acc[2 * i + 1] += pk[i] & 0x00010000u; }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sink + __float_as_uint(x_dummy(acc));
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// TMEM variants: the accumulator block comes from tensor memory (tcgen05.ld 32x32b.x32) and the packed result goes back
// (tcgen05.st 32x32b.x16), as in the fused kernel, but no MMA runs.  MODE 0: ld, wait, math, st per block; MODE 1: the next block's
// load is issued before this block's math (two register sets); MODE 2: both blocks loaded up front, one wait.
template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) ktmem(uint32_t* out, int iters, long long* cycles, const float* bias_g) {
  __shared__ float bias[128];
  __shared__ uint32_t slot;
  if (threadIdx.x < 128) bias[threadIdx.x] = bias_g[threadIdx.x];
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t sink = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    asm volatile("bar.sync 1, %0;" ::"r"(THREADS) : "memory");
    tcgen05_fence_after();
    if (MODE == 0) {
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        uint32_t acc[32], pk[16];
        float2 x[16], u[16];
        tmem_ld32_nowait(tb + blk * 32, acc);
        tmem_ld_wait();
        half_a(acc, bias + blk * 32, x, u);
        half_b(x, u, pk);
        tmem_st16_u32(tb + blk * 32, pk);
        sink ^= pk[3];
      }
    } else if (MODE == 1) {
      uint32_t acc0[32], acc1[32], pk[16];
      float2 x[16], u[16];
      tmem_ld32_nowait(tb, acc0);
      tmem_ld_wait();
      tmem_ld32_nowait(tb + 32, acc1);
      half_a(acc0, bias, x, u);
      half_b(x, u, pk);
      tmem_st16_u32(tb, pk);
      sink ^= pk[3];
      tmem_ld_wait();
      half_a(acc1, bias + 32, x, u);
      half_b(x, u, pk);
      tmem_st16_u32(tb + 32, pk);
      sink ^= pk[3];
    } else {
      uint32_t acc0[32], acc1[32], pk[16];
      float2 x[16], u[16];
      tmem_ld32_nowait(tb, acc0);
      tmem_ld32_nowait(tb + 32, acc1);
      tmem_ld_wait();
      half_a(acc0, bias, x, u);
      half_b(x, u, pk);
      tmem_st16_u32(tb, pk);
      sink ^= pk[3];
      half_a(acc1, bias + 32, x, u);
      half_b(x, u, pk);
      tmem_st16_u32(tb + 32, pk);
      sink ^= pk[3];
    }
    tmem_st_wait();
    tcgen05_fence_before();
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sink;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<512>(slot); }
}

template <int MODE, int THREADS>
void runtmem(const char* name) {
  uint32_t* out;
  long long* cyc;
  float* bias;
  cudaMalloc(&out, 148 * 1024 * sizeof(uint32_t));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaMalloc(&bias, 128 * sizeof(float));
  cudaMemset(bias, 0, 128 * sizeof(float));
  const int iters = 4000;
  ktmem<MODE, THREADS><<<148, THREADS>>>(out, iters, cyc, bias);
  ktmem<MODE, THREADS><<<148, THREADS>>>(out, iters, cyc, bias);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double elems = (double)iters * 64 * THREADS;
  printf("%-40s warps %2d: %.2f elements/clk/SM (128x128 chunk = %.0f cycles)\n", name, THREADS / 32, elems / avg, 16384.0 / (elems / avg));
  cudaFree(out); cudaFree(cyc); cudaFree(bias);
}

template <int MODE, int THREADS>
void runlock(const char* name) {
  uint32_t* out;
  long long* cyc;
  float* bias;
  cudaMalloc(&out, 148 * 1024 * sizeof(uint32_t));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaMalloc(&bias, 128 * sizeof(float));
  cudaMemset(bias, 0, 128 * sizeof(float));
  const int iters = 4000;
  klock<MODE, THREADS><<<148, THREADS>>>(out, iters, cyc, bias);
  klock<MODE, THREADS><<<148, THREADS>>>(out, iters, cyc, bias);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double elems = (double)iters * 64 * THREADS;
  printf("%-40s warps %2d: %.2f elements/clk/SM (128x128 chunk = %.0f cycles)\n", name, THREADS / 32, elems / avg, 16384.0 / (elems / avg));
  cudaFree(out); cudaFree(cyc); cudaFree(bias);
}

template <int MODE, int THREADS>
void run(const char* name) {
  uint32_t* out;
  long long* cyc;
  float* bias;
  cudaMalloc(&out, 148 * 1024 * sizeof(uint32_t));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaMalloc(&bias, 128 * sizeof(float));
  cudaMemset(bias, 0, 128 * sizeof(float));
  const int iters = 4000;
  k<MODE, THREADS><<<148, THREADS>>>(out, iters, cyc, bias);
  k<MODE, THREADS><<<148, THREADS>>>(out, iters, cyc, bias);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const int nb = (MODE & 1) ? 16 : 32;
  const double elems = (double)iters * nb * THREADS;
  printf("%-40s warps %2d: %.2f elements/clk/SM (128x128 chunk = %.0f cycles)\n", name, THREADS / 32, elems / avg, 16384.0 / (elems / avg));
  cudaFree(out); cudaFree(cyc); cudaFree(bias);
}

int main() {
  run<0, 256>("fwd gelu, 32 per batch");
  run<0, 512>("fwd gelu, 32 per batch");
  run<0, 1024>("fwd gelu, 32 per batch");
  run<1, 256>("fwd gelu, 16 per batch");
  run<1, 512>("fwd gelu, 16 per batch");
  run<2, 256>("bwd gelu + derivative, 32 per batch");
  run<2, 512>("bwd gelu + derivative, 32 per batch");
  run<2, 1024>("bwd gelu + derivative, 32 per batch");
  run<3, 256>("bwd gelu + derivative, 16 per batch");
  run<3, 512>("bwd gelu + derivative, 16 per batch");
  runlock<0, 256>("lockstep per chunk, block after block");
  runlock<1, 256>("lockstep per chunk, software pipelined");
  runlock<0, 512>("lockstep per chunk, block after block");
  runlock<1, 512>("lockstep per chunk, software pipelined");
  runtmem<0, 256>("TMEM ld/st, block after block");
  runtmem<1, 256>("TMEM ld/st, next load in flight");
  runtmem<2, 256>("TMEM ld/st, both loads up front");
  return 0;
}
