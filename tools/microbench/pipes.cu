// Measures issue rates of the fp32 FMA forms and the MUFU ops on this GPU (sm_100a): per-SM operations per clock.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters, long long* cycles) {
  float2 a[8];
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); s[i] = threadIdx.x * 1e-3f + i; }
  const float2 w = make_float2(1.0001f, 0.9999f), b = make_float2(1e-6f, -1e-6f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) s[i] = fmaf(s[i], w.x, b.x);                          // scalar FFMA, 3 registers
        if (MODE == 1) a[i] = __ffma2_rn(a[i], w, b);                        // FFMA2
        if (MODE == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(s[i]));  // MUFU.TANH
        if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s[i]));
        if (MODE == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(s[i]));
        if (MODE == 5) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(s[i]));
        if (MODE == 6) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(s[i]));
        if (MODE == 7) {  // GELU-like mix: 1 MUFU + 2.5 packed FMA per element (two elements: 2 MUFU + 5 FFMA2)
          asm volatile("tanh.approx.f32 %0, %0;" : "+f"(s[i]));
          a[i] = __ffma2_rn(a[i], w, b);
          a[i] = __ffma2_rn(a[i], w, b);
          if (i & 1) a[i] = __ffma2_rn(a[i], w, b);
        }
        if (MODE == 8) {  // same FMA work without the MUFU
          a[i] = __ffma2_rn(a[i], w, b);
          a[i] = __ffma2_rn(a[i], w, b);
          if (i & 1) a[i] = __ffma2_rn(a[i], w, b);
        }
        if (MODE == 10) a[i] = __ffma2_rn(a[(i + 3) & 7], a[(i + 5) & 7], a[i]);      // FFMA2, three distinct register pairs
        if (MODE == 11) s[i] = fmaf(s[(i + 3) & 7], s[(i + 5) & 7], s[i]);            // FFMA, three distinct registers
        if (MODE == 12) a[i] = __ffma2_rn(w, a[(i + 5) & 7], a[i]);                   // FFMA2, one operand shared by consecutive instructions (reuse)
        if (MODE == 9) {  // MUFU + independent integer/ALU work (4 ALU ops per MUFU)
          asm volatile("tanh.approx.f32 %0, %0;" : "+f"(s[i]));
          unsigned u = __float_as_uint(a[i].x);
          u = (u << 1) ^ 0x9e3779b9u; u = (u >> 3) + 0x7f4a7c15u; u = (u << 2) ^ 0x85ebca6bu; u = (u >> 1) + 0xc2b2ae35u;
          a[i].x = __uint_as_float(u);
        }
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y + s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_instr) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int iters = 2000;
  k<MODE><<<148, 1024>>>(out, iters, cyc);
  k<MODE><<<148, 1024>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double instr = (double)iters * 32 * 1024 / 32;  // warp instructions per SM
  printf("%-28s %8.1f cycles  -> %.2f warp-instr/clk/SM = %.1f lane-ops/clk/SM\n", name, avg, instr / avg, instr / avg * 32 * ops_per_instr);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("FFMA (scalar, 3 regs)", 1);
  run<1>("FFMA2 (packed fp32x2)", 2);
  run<2>("MUFU.TANH", 1);
  run<3>("MUFU.EX2", 1);
  run<4>("MUFU.RCP", 1);
  run<5>("MUFU.LG2", 1);
  run<6>("MUFU.RSQ", 1);
  run<7>("1 TANH + 2.5 FFMA2 (mix)", 1);
  run<8>("2.5 FFMA2 alone", 1);
  run<9>("1 TANH + 8 ALU ops (mix)", 1);
  run<10>("FFMA2 3 distinct reg pairs", 2);
  run<11>("FFMA 3 distinct regs", 1);
  run<12>("FFMA2 shared 1st operand", 2);
  return 0;
}
