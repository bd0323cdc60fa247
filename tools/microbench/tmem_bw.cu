// TMEM -> register bandwidth of tcgen05.ld (32x32b.x32) per SM as a function of the number of warps issuing loads.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../linnaeus_b200/csrc -o tmem_bw tmem_bw.cu
#include <cstdio>
#include "lnx_tc_common.cuh"
using namespace lnx_tc;

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

template <int INFLIGHT>
__global__ void k(unsigned* out, long long* cycles, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[INFLIGHT][32];
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j) ld32(base + ((it * INFLIGHT + j) * 32) % 512, r[j]);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j)
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= r[j][i];
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<512>(slot); }
}

template <int INFLIGHT>
void run(int warps) {
  unsigned* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4000;
  k<INFLIGHT><<<148, warps * 32>>>(out, cyc, iters);
  k<INFLIGHT><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double bytes = (double)iters * INFLIGHT * warps * 32 * 32 * 4;
  printf("warps %2d, %d loads in flight: %.1f B/clk/SM  (%.0f cycles per x32 load per warp)\n", warps, INFLIGHT, bytes / avg, avg / (iters * INFLIGHT));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) { run<1>(w); run<2>(w); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
