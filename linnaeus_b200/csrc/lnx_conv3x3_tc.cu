// Dense 3x3 convolution (stride 1, pad 1) on NHWC bf16 as an IMPLICIT GEMM on tcgen05: no im2col buffer.
//
// Output tile = [8 rows x 16 columns] of pixels of one image (M = 128) x all output channels (N <= 256, multiple of 16).
// The reduction runs over the 9 taps; for tap (kh, kw) the A operand is the same 8 x 16 pixel window shifted by
// (kh - 1, kw - 1), fetched by ONE 4-D TMA load over the [B][H][W][C] tensor: box (64 channels, 16, 8, 1) lands in
// shared memory as [128 pixels][64 channels] = exactly a K-major, 128-byte-swizzled UMMA tile, and the hardware zero
// fill outside the image IS the conv padding (channels beyond C are zero filled too, so C <= 64 needs no repacking).
// The 9 weight tiles [N][64] stay resident in shared memory for the CTA's lifetime.
//
// Roles (192 threads, persistent CTAs): warp 0 TMA producer (4-stage A ring), warp 1 MMA issuer (2 TMEM accumulators,
// so the next tile's MMAs overlap this tile's epilogue), warps 2-5 epilogue: tcgen05.ld -> + bias (folded BatchNorm)
// -> ReLU -> bf16 -> swizzled staging -> one 4-D TMA store per warp (its 2 image rows x 16 columns x N channels).
// Used by the mFormerV0 stem (R/models/mFormerV0.py:166-190).
#include "lnx_tc_common.cuh"

using namespace lnx;
using namespace lnx_tc;

namespace {

constexpr int TH = 8, TW = 16;         // output tile (pixels)
constexpr int KC = 64;                 // channels per tap (zero padded)
constexpr int A_BYTES = 128 * KC * 2;  // 16 KB
constexpr int STAGES = 4;
constexpr int NUM_THREADS = 192;

struct ConvParams {
  int B, H, W, N;  // N = output channels
  int tiles_h, tiles_w;
  int relu;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
    conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmY,
                      const float* __restrict__ bias, const ConvParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int w_tile_bytes = p.N * KC * 2;                    // one tap: [N][64] K-major
  unsigned char* sW = base;                                 // [9][N][64]
  unsigned char* sA = sW + 9 * (size_t)w_tile_bytes;        // [STAGES][128][64]
  unsigned char* sOut = sA + STAGES * A_BYTES;              // [4 warps][N / 64 blocks][32 rows][64] staging
  const int n_blocks = (p.N + 63) / 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + 4 * (size_t)n_blocks * 4096);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;   // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint64_t* wfull = tempty + 2;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int num_tiles = p.B * tiles_per_img;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmW); prefetch_tmap(&tmY);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 4);
    }
    mbar_init(wfull, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // resident weights: 9 taps x [N rows][64 k]
      mbar_expect_tx(wfull, 9 * (uint32_t)w_tile_bytes);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(sW + (size_t)tap * w_tile_bytes, &tmW, wfull, tap * KC, 0);
      uint32_t it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int b = t / tiles_per_img, r = t - b * tiles_per_img;
        const int h0 = (r / p.tiles_w) * TH, w0 = (r % p.tiles_w) * TW;
        for (int tap = 0; tap < 9; ++tap, ++it) {
          const int s = it % STAGES;
          mbar_wait_relaxed(&empty[s], ((it / STAGES) & 1u) ^ 1u);
          mbar_expect_tx(&full[s], A_BYTES);
          tma_load_4d(sA + (size_t)s * A_BYTES, &tmX, &full[s], 0, w0 + tap % 3 - 1, h0 + tap / 3 - 1, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
      mbar_wait(wfull, 0);
      tcgen05_fence_after();
      const uint32_t aW = smem_u32(sW);
      uint32_t it = 0, tl = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tl) {
        const uint32_t as = tl & 1u;
        mbar_wait_relaxed(&tempty[as], ((tl >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem + as * 256;
        for (int tap = 0; tap < 9; ++tap, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full[s], (it / STAGES) & 1u);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(sA + (size_t)s * A_BYTES);
          const uint32_t sb = aW + tap * w_tile_bytes;
#pragma unroll
          for (int k = 0; k < KC / 16; ++k)
            umma_bf16(tacc, make_smem_desc(sa + k * 32, 0, 1024), make_smem_desc(sb + k * 32, 0, 1024), idesc, (tap > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[as]);
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter = tile rows 2q, 2q + 1 (16 pixels each)
    unsigned char* st = sOut + (size_t)(warp - 2) * n_blocks * 4096;
    const int r_sw = lane & 7;
    uint32_t tl = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tl) {
      const int b = t / tiles_per_img, r = t - b * tiles_per_img;
      const int h0 = (r / p.tiles_w) * TH, w0 = (r % p.tiles_w) * TW;
      const uint32_t as = tl & 1u;
      mbar_wait(&tfull[as], (tl >> 1) & 1u);
      tcgen05_fence_after();
      const uint32_t trow = tmem + as * 256 + ((uint32_t)(q * 32) << 16);
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous stores have read the staging blocks
      __syncwarp();
      for (int c = 0; c < p.N; c += 16) {
        float v[16];
        __syncwarp();
        tmem_ld16(trow + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          v[i] += bias ? __ldg(bias + c + i) : 0.f;
          if (p.relu) v[i] = fmaxf(v[i], 0.f);
        }
        unsigned char* blk = st + (c >> 6) * 4096;
        const int j = (c & 63) >> 3;
        *reinterpret_cast<uint4*>(blk + lane * 128 + ((j ^ r_sw) << 4)) = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
        *reinterpret_cast<uint4*>(blk + lane * 128 + (((j + 1) ^ r_sw) << 4)) =
            make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tempty[as]);
        for (int nb = 0; nb < n_blocks; ++nb) tma_store_4d(&tmY, st + nb * 4096, nb * 64, w0, h0 + 2 * q, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace

// x [B, H, W, C] bf16 (C <= 64, C % 8 == 0), w9 [N, 9 * 64] bf16 (row n: tap-major, 64 channels per tap, zero padded),
// bias float [N] (nullable), y [B, H, W, N] bf16.  N % 16 == 0, N <= 256.
extern "C" int lnx_conv3x3_s1(const void* x, const void* w9, const float* bias, void* y, int B, int H, int W, int C, int N, int relu, int dtype,
                              lnx_stream_t s) {
  LNX_REQUIRE(x && w9 && y, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && N > 0, LNX_ERR_SHAPE);
  if (dtype != LNX_BF16 || C > KC || C % 8 != 0 || N % 16 != 0 || N > 256) return LNX_ERR_UNSUPPORTED;
  LNX_REQUIRE(lnx_aligned16(x) && lnx_aligned16(w9) && lnx_aligned16(y), LNX_ERR_ALIGN);
  ConvParams p;
  p.B = B; p.H = H; p.W = W; p.N = N; p.relu = relu;
  p.tiles_h = (H + TH - 1) / TH;
  p.tiles_w = (W + TW - 1) / TW;
  CUtensorMap tmX, tmW, tmY;
  {
    const long long dims[4] = {C, W, H, B};
    const long long strides[3] = {C, (long long)W * C, (long long)H * W * C};
    const int box[4] = {KC, TW, TH, 1};
    if (!make_tmap(&tmX, x, 4, dims, strides, box)) return LNX_ERR_UNSUPPORTED;
  }
  {
    const long long dims[2] = {9 * KC, N};
    const long long strides[1] = {9 * KC};
    const int box[2] = {KC, N};
    if (!make_tmap(&tmW, w9, 2, dims, strides, box)) return LNX_ERR_UNSUPPORTED;
  }
  {
    const long long dims[4] = {N, W, H, B};
    const long long strides[3] = {N, (long long)W * N, (long long)H * W * N};
    const int box[4] = {64, TW, 2, 1};
    if (!make_tmap(&tmY, y, 4, dims, strides, box)) return LNX_ERR_UNSUPPORTED;
  }
  const int n_blocks = (N + 63) / 64;
  const size_t smem = 1024 + 9 * (size_t)N * KC * 2 + STAGES * A_BYTES + 4 * (size_t)n_blocks * 4096 + 256;
  if (smem > 232448) return LNX_ERR_UNSUPPORTED;
  static int smem_set = 0;
  if ((int)smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    smem_set = (int)smem;
  }
  const int num_tiles = B * p.tiles_h * p.tiles_w;
  conv3x3_tc_kernel<<<min(num_tiles, kNumSMs), NUM_THREADS, smem, (cudaStream_t)s>>>(tmX, tmW, tmY, bias, p);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
