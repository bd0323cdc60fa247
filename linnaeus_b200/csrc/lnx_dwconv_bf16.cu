// Depthwise 7x7 (pad 3, stride 1) on NHWC bf16 with packed fp32x2 FMAs (FFMA2, sm_100): forward / data
// gradient and weight gradient.
//
// The op is bound by the fp32 FMA pipe (49 MACs per output element against 4 bytes of traffic), and on
// Blackwell the 3-register scalar FFMA issues at half rate: the full rate needs fma.rn.f32x2.  So a lane owns a
// channel PAIR (one 32-bit shared load = both channels of one pixel, one FFMA2 = both channels of one tap) and
// keeps fp32x2 accumulators for a strip of 7 outputs per row, so every loaded pixel feeds up to 7 FFMA2.
// A warp = 16 channel pairs x two half strips (columns 0-6 / 7-13 of a tile row); the two 64-byte segments a
// warp reads are 7 pixels = 448 bytes apart, i.e. on disjoint shared-memory banks.
//
// Work item = one 14x14 output tile of one image x one 32-channel chunk.  CTAs are persistent per channel chunk
// and double buffered: a producer warp issues the TMA loads (4-D tensor map over [B][H][W][C]; negative /
// overhanging coordinates are zero filled by the hardware = the conv padding) of the NEXT 20x20 halo tile while the
// seven compute warps work on the current one.  There is no CTA-wide barrier in the steady state: buffers are
// handed over with full / empty mbarriers, so compute warps drift freely by up to one tile.  The weight-gradient kernel keeps 49 fp32x2
// partial sums per lane in registers across all tiles the CTA visits.
#include <stdlib.h>

#include "lnx_common.cuh"
#include "lnx_tc_common.cuh"

using namespace lnx;
using namespace lnx_tc;

namespace {

constexpr int TILE = 14;
constexpr int HALO = TILE + 6;  // 20
constexpr int NWARPS = 7;       // compute warp w -> tile rows 2w, 2w+1
constexpr int NCOMPUTE = NWARPS * 32;
constexpr int NTHREADS = NCOMPUTE + 32;  // + one producer warp (TMA)
constexpr int PW = 16;          // channel pairs per pixel of a chunk
constexpr int CC = 2 * PW;      // 32 channels per chunk
constexpr int HALO_BYTES = HALO * HALO * CC * 2;  // 25600
constexpr int CENTER_BYTES = TILE * TILE * CC * 2;  // 12544
constexpr int CENTER_PAD = 12800;                   // keeps every buffer 256-byte aligned

// weight element (tap, channel) in the caller's layout: 0 = [49][C] tap-major, 1 = the Conv2d weight itself [C][49], 2 = [C][49] read with the
// taps reversed (the data gradient convolves dY with the flipped filter)
__device__ __forceinline__ long long widx(int wl, int tap, int c, int C) {
  return wl == 0 ? (long long)tap * C + c : (long long)c * 49 + (wl == 2 ? 48 - tap : tap);
}

__device__ __forceinline__ float2 unpack_bf16x2(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

struct TileCoord {
  int b, h0, w0;
};
__device__ __forceinline__ TileCoord tile_coord(int t, int tiles_w, int tiles_h) {
  TileCoord c;
  const int tw = t % tiles_w;
  t /= tiles_w;
  c.w0 = tw * TILE;
  c.h0 = (t % tiles_h) * TILE;
  c.b = t / tiles_h;
  return c;
}

// ------------------------------------------------------------------ forward (and data gradient with flipped taps)
template <int NS>  // halo-tile ring depth: 2 (three CTAs per SM) or 3 (two CTAs per SM; a full tile of compute between load and use)
__global__ void __launch_bounds__(NTHREADS, NS == 2 ? 3 : 2)
    dwconv7_fwd_x2_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ w49c, const float* __restrict__ bias,
                          const bf16* __restrict__ res, bf16* __restrict__ y, int B, int H, int W, int C, int tiles_w, int tiles_h, int wl) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* tiles = smem_raw;                                              // [2][20][20][CC] bf16
  float2* wsm = reinterpret_cast<float2*>(smem_raw + NS * HALO_BYTES);           // [49][PW]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + NS * HALO_BYTES + 49 * CC * 4);  // [NS]
  uint64_t* empty = full + NS;                                                  // [NS]

  const int c0 = blockIdx.y * CC;
  const int total = B * tiles_h * tiles_w;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NWARPS);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 49 * CC; i += NTHREADS) reinterpret_cast<float*>(wsm)[i] = w49c[widx(wl, i / CC, c0 + (i % CC), C)];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == NWARPS) {
    // ===================== producer =====================
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int buf = it % NS;
        mbar_wait_relaxed(&empty[buf], (((uint32_t)it / NS) & 1u) ^ 1u);
        const TileCoord tc = tile_coord(t, tiles_w, tiles_h);
#ifdef LNX_DW_NOLOAD  // profiling ablation (tools/build_variant.sh): no halo loads, the compute warps run on whatever is in shared memory
        (void)tc;
        mbar_arrive(&full[buf]);
#else
        mbar_expect_tx(&full[buf], HALO_BYTES);
        tma_load_4d(tiles + buf * HALO_BYTES, &tmX, &full[buf], c0, tc.w0 - 3, tc.h0 - 3, tc.b);
#endif
      }
    }
    return;
  }
  const int p = lane & (PW - 1);
  const int col0 = (lane >> 4) * 7;
  const int orow0 = warp * 2;
  float2 bv = make_float2(0.f, 0.f);
  if (bias) bv = *reinterpret_cast<const float2*>(bias + c0 + 2 * p);

  int it = 0;
  for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
    const int buf = it % NS;
    const TileCoord tc = tile_coord(t, tiles_w, tiles_h);
    mbar_wait(&full[buf], ((uint32_t)it / NS) & 1u);

    float2 acc[2][7];
    if (tc.h0 + orow0 < H) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int o = 0; o < 7; ++o) acc[rr][o] = bv;
      const uint32_t* tp = reinterpret_cast<const uint32_t*>(tiles + buf * HALO_BYTES) + ((orow0 * HALO + col0) * PW + p);
#pragma unroll
      for (int kh = 0; kh < 7; ++kh) {
        float2 wk[7];
#pragma unroll
        for (int kw = 0; kw < 7; ++kw) wk[kw] = wsm[(kh * 7 + kw) * PW + p];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const uint32_t* rowp = tp + (rr + kh) * HALO * PW;
#pragma unroll
          for (int ix = 0; ix < 13; ++ix) {
            const float2 v = unpack_bf16x2(rowp[ix * PW]);
#pragma unroll
            for (int kw = 0; kw < 7; ++kw) {
              const int o = ix - kw;
              if (o >= 0 && o < 7) acc[rr][o] = ffma2(v, wk[kw], acc[rr][o]);
            }
          }
        }
      }
    }
    // this warp no longer reads tiles[buf]: hand it back to the producer before the (slow) global stores
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[buf]);
    if (tc.h0 + orow0 < H) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int hh = tc.h0 + orow0 + rr;
        if (hh < H) {
          const long long off = (((long long)tc.b * H + hh) * W + tc.w0 + col0) * C + c0 + 2 * p;
          bf16* yrow = y + off;
          if (res) {  // fused "+ residual" (the skip-connection gradient when this kernel runs as the data gradient)
#pragma unroll
            for (int o = 0; o < 7; ++o)
              if (tc.w0 + col0 + o < W) {
                const float2 rv = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(res + off + (long long)o * C)));
                acc[rr][o].x += rv.x;
                acc[rr][o].y += rv.y;
              }
          }
#pragma unroll
          for (int o = 0; o < 7; ++o)
            if (tc.w0 + col0 + o < W)
              *reinterpret_cast<__nv_bfloat162*>(yrow + (long long)o * C) = __floats2bfloat162_rn(acc[rr][o].x, acc[rr][o].y);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ forward, 4 output rows per lane
// Same lane layout (channel pair x half strip of 7 columns), but a warp owns FOUR output rows of a 28 x 14 tile and walks the
// INPUT rows: every pixel loaded from shared memory feeds all the (output row, filter row) pairs that touch it, so the
// shared-memory loads per FFMA2 drop from 1 / 3.0 to 1 / 6.3 (the 14 x 14 / 2-row kernel re-reads each input row once per filter
// row and is co-limited by the issue slots and the shared-memory pipe at 45 % of the FMA rate).  The 49 taps do not fit in
// registers next to 28 accumulators, so the filter rows are split in two passes (kh 0-3, kh 4-6) with their taps reloaded per pass.
constexpr int R4_ROWS = 28;                 // tile rows: 7 warps x 4
constexpr int R4_HALO_H = R4_ROWS + 6;      // 34
constexpr int R4_HALO_BYTES = R4_HALO_H * HALO * CC * 2;  // 34 x 20 x 32 ch x 2 B = 43 520

template <int KH0, int KH1>
__device__ __forceinline__ void dw_r4_pass(const uint32_t* __restrict__ tp, const float2* __restrict__ wsm, int p, float2 (&acc)[4][7]) {
  float2 wk[KH1 - KH0][7];
#pragma unroll
  for (int kh = KH0; kh < KH1; ++kh)
#pragma unroll
    for (int kw = 0; kw < 7; ++kw) wk[kh - KH0][kw] = wsm[(kh * 7 + kw) * PW + p];
#pragma unroll
  for (int ir = KH0; ir < KH1 + 3; ++ir) {  // input row relative to the warp's first output row: ir = rr + kh
    const uint32_t* rowp = tp + ir * HALO * PW;
#pragma unroll
    for (int ix = 0; ix < 13; ++ix) {
      const float2 v = unpack_bf16x2(rowp[ix * PW]);
#pragma unroll
      for (int kh = KH0; kh < KH1; ++kh) {
        const int rr = ir - kh;
        if (rr >= 0 && rr < 4) {
#pragma unroll
          for (int kw = 0; kw < 7; ++kw) {
            const int o = ix - kw;
            if (o >= 0 && o < 7) acc[rr][o] = ffma2(v, wk[kh - KH0][kw], acc[rr][o]);
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 2)
    dwconv7_fwd_r4_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ w49c, const float* __restrict__ bias,
                          const bf16* __restrict__ res, bf16* __restrict__ y, int B, int H, int W, int C, int tiles_w, int tiles_h, int wl) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* tiles = smem_raw;                                                   // [2][34][20][CC] bf16
  float2* wsm = reinterpret_cast<float2*>(smem_raw + 2 * R4_HALO_BYTES);             // [49][PW]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + 2 * R4_HALO_BYTES + 49 * CC * 4);
  uint64_t* empty = full + 2;

  const int c0 = blockIdx.y * CC;
  const int total = B * tiles_h * tiles_w;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NWARPS);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 49 * CC; i += NTHREADS) reinterpret_cast<float*>(wsm)[i] = w49c[widx(wl, i / CC, c0 + (i % CC), C)];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto coord = [&](int t, int& b, int& h0, int& w0) {
    const int tw = t % tiles_w;
    t /= tiles_w;
    w0 = tw * TILE;
    h0 = (t % tiles_h) * R4_ROWS;
    b = t / tiles_h;
  };
  if (warp == NWARPS) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait_relaxed(&empty[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
        int b, h0, w0;
        coord(t, b, h0, w0);
        mbar_expect_tx(&full[buf], R4_HALO_BYTES);
        tma_load_4d(tiles + buf * R4_HALO_BYTES, &tmX, &full[buf], c0, w0 - 3, h0 - 3, b);
      }
    }
    return;
  }
  const int p = lane & (PW - 1);
  const int col0 = (lane >> 4) * 7;
  const int orow0 = warp * 4;
  float2 bv = make_float2(0.f, 0.f);
  if (bias) bv = *reinterpret_cast<const float2*>(bias + c0 + 2 * p);

  int it = 0;
  for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
    const int buf = it & 1;
    int b, h0, w0;
    coord(t, b, h0, w0);
    mbar_wait(&full[buf], ((uint32_t)it >> 1) & 1u);
    const bool active = h0 + orow0 < H;
    float2 acc[4][7];
    if (active) {
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int o = 0; o < 7; ++o) acc[rr][o] = bv;
      const uint32_t* tp = reinterpret_cast<const uint32_t*>(tiles + buf * R4_HALO_BYTES) + ((orow0 * HALO + col0) * PW + p);
      dw_r4_pass<0, 4>(tp, wsm, p, acc);
      dw_r4_pass<4, 7>(tp, wsm, p, acc);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[buf]);
    if (active) {
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int hh = h0 + orow0 + rr;
        if (hh < H) {
          const long long off = (((long long)b * H + hh) * W + w0 + col0) * C + c0 + 2 * p;
          bf16* yrow = y + off;
          if (res) {
#pragma unroll
            for (int o = 0; o < 7; ++o)
              if (w0 + col0 + o < W) {
                const float2 rv = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(res + off + (long long)o * C)));
                acc[rr][o].x += rv.x;
                acc[rr][o].y += rv.y;
              }
          }
#pragma unroll
          for (int o = 0; o < 7; ++o)
            if (w0 + col0 + o < W)
              *reinterpret_cast<__nv_bfloat162*>(yrow + (long long)o * C) = __floats2bfloat162_rn(acc[rr][o].x, acc[rr][o].y);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ weight gradient
__global__ void __launch_bounds__(NTHREADS, 2)
    dwconv7_wgrad_x2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, float* __restrict__ dw49c,
                            float* __restrict__ dbias, int B, int H, int W, int C, int tiles_w, int tiles_h, int wl) {
  constexpr int STAGE = HALO_BYTES + CENTER_PAD;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* stages = smem_raw;                                        // [2]{x halo [20][20][CC], dy [14][14][CC]}
  float* red = reinterpret_cast<float*>(smem_raw + 2 * STAGE);             // [50][CC]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + 2 * STAGE + 50 * CC * 4);
  uint64_t* empty = full + 2;

  const int c0 = blockIdx.y * CC;
  const int total = B * tiles_h * tiles_w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = lane & (PW - 1);
  const int col0 = (lane >> 4) * 7;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmG);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NWARPS);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 50 * CC; i += NTHREADS) red[i] = 0.f;
  __syncthreads();

  float2 wacc[49];
  float2 bacc = make_float2(0.f, 0.f);
  if (warp == NWARPS) {
    // ===================== producer =====================
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait_relaxed(&empty[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
        const TileCoord tc = tile_coord(t, tiles_w, tiles_h);
        unsigned char* st = stages + buf * STAGE;
        mbar_expect_tx(&full[buf], HALO_BYTES + CENTER_BYTES);
        tma_load_4d(st, &tmX, &full[buf], c0, tc.w0 - 3, tc.h0 - 3, tc.b);
        tma_load_4d(st + HALO_BYTES, &tmG, &full[buf], c0, tc.w0, tc.h0, tc.b);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 49; ++k) wacc[k] = make_float2(0.f, 0.f);
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&full[buf], ((uint32_t)it >> 1) & 1u);
      const unsigned char* st = stages + buf * STAGE;
#pragma unroll 1
      for (int rr = 0; rr < 2; ++rr) {  // strips of 7 outputs: 13 tile loads feed 49 FFMA2 per filter row
        const int orow = warp * 2 + rr;
        float2 g[7];
        const uint32_t* gp = reinterpret_cast<const uint32_t*>(st + HALO_BYTES) + ((orow * TILE + col0) * PW + p);
#pragma unroll
        for (int o = 0; o < 7; ++o) {
          g[o] = unpack_bf16x2(gp[o * PW]);  // rows / columns beyond the image were zero filled
          bacc.x += g[o].x;
          bacc.y += g[o].y;
        }
        const uint32_t* tp = reinterpret_cast<const uint32_t*>(st) + ((orow * HALO + col0) * PW + p);
#pragma unroll
        for (int kh = 0; kh < 7; ++kh) {
          const uint32_t* rowp = tp + kh * HALO * PW;
#pragma unroll
          for (int ix = 0; ix < 13; ++ix) {
            const float2 v = unpack_bf16x2(rowp[ix * PW]);
#pragma unroll
            for (int kw = 0; kw < 7; ++kw) {
              const int o = ix - kw;
              if (o >= 0 && o < 7) wacc[kh * 7 + kw] = ffma2(v, g[o], wacc[kh * 7 + kw]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[buf]);
    }
#pragma unroll
    for (int k = 0; k < 49; ++k) {
      atomicAdd(&red[k * CC + 2 * p], wacc[k].x);
      atomicAdd(&red[k * CC + 2 * p + 1], wacc[k].y);
    }
    atomicAdd(&red[49 * CC + 2 * p], bacc.x);
    atomicAdd(&red[49 * CC + 2 * p + 1], bacc.y);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 49 * CC; i += NTHREADS) atomicAdd(dw49c + widx(wl, i / CC, c0 + (i % CC), C), red[i]);
  if (dbias)
    for (int i = threadIdx.x; i < CC; i += NTHREADS) atomicAdd(dbias + c0 + i, red[49 * CC + i]);
}

// [B][H][W][C] bf16, box = (CC channels, bw, bh, 1 image), no swizzle, zero fill outside
bool make_nhwc_tmap(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int bw, int bh) {
  auto enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)CC, (cuuint32_t)bw, (cuuint32_t)bh, 1u};
  cuuint32_t es[4] = {1u, 1u, 1u, 1u};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// tensor-pipe kernels of lnx_dwconv_mma.cu (LNX_ERR_UNSUPPORTED = shape not covered: use the kernels of this file)
bool lnx_dwconv7_mma_enabled();
int lnx_dwconv7_fwd_mma(const void* x, const float* w, int wl, const float* bias, const void* res, void* y, int B, int H, int W, int C, cudaStream_t st);
int lnx_dwconv7_wgrad_mma(const void* x, const void* dy, float* dw, int wl, float* dbias, int B, int H, int W, int C, cudaStream_t st);

int lnx_dwconv7_fwd_bf16(const void* x, const float* w49c, int wl, const float* bias, const void* res, void* y, int B, int H, int W, int C,
                         cudaStream_t st) {
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 7u)) return LNX_ERR_ALIGN;
  if (C % CC != 0) return LNX_ERR_SHAPE;
  if (lnx_dwconv7_mma_enabled() && lnx_aligned16(res)) {
    const int rc = lnx_dwconv7_fwd_mma(x, w49c, wl, bias, res, y, B, H, W, C, st);
    if (rc != LNX_ERR_UNSUPPORTED) return rc;
  }
  const int chunks = C / CC;
  const int tiles_w = (W + TILE - 1) / TILE;
  // LNX_DWCONV_KERNEL: 4 = four output rows per lane (28 x 14 tiles; measured faster only for large images: used when H >= 64), 2 / 3 = the
  // two-row kernel with a 2- or 3-deep halo ring
  static const int sel = getenv("LNX_DWCONV_KERNEL") ? atoi(getenv("LNX_DWCONV_KERNEL")) : 4;
  if (sel == 4 && H >= 64) {
    CUtensorMap tmX;
    if (!make_nhwc_tmap(&tmX, x, B, H, W, C, HALO, R4_HALO_H)) return LNX_ERR_UNSUPPORTED;
    const int tiles_h = (H + R4_ROWS - 1) / R4_ROWS;
    const size_t smem = 2 * (size_t)R4_HALO_BYTES + 49 * CC * 4 + 64;
    static bool attr4 = false;
    if (!attr4) {
      cudaError_t e = cudaFuncSetAttribute(dwconv7_fwd_r4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return lnx_set_cuda_error(e);
      attr4 = true;
    }
    const int total = B * tiles_h * tiles_w;
    const int gx = max(1, min(total, (kNumSMs * 2 + chunks - 1) / chunks));
    dwconv7_fwd_r4_kernel<<<dim3(gx, chunks), NTHREADS, smem, st>>>(tmX, w49c, bias, (const bf16*)res, (bf16*)y, B, H, W, C, tiles_w, tiles_h, wl);
    LNX_CHECK_LAUNCH();
    return LNX_OK;
  }
  CUtensorMap tmX;
  if (!make_nhwc_tmap(&tmX, x, B, H, W, C, HALO, HALO)) return LNX_ERR_UNSUPPORTED;
  const int tiles_h = (H + TILE - 1) / TILE;
  const int ns = sel == 3 ? 3 : 2;
  const size_t smem = (size_t)ns * HALO_BYTES + 49 * CC * 4 + 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = ns == 3 ? cudaFuncSetAttribute(dwconv7_fwd_x2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                            : cudaFuncSetAttribute(dwconv7_fwd_x2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    attr_set = true;
  }
  const int total = B * tiles_h * tiles_w;
  const int gx = max(1, min(total, (kNumSMs * (ns == 3 ? 2 : 3) + chunks - 1) / chunks));
  if (ns == 3)
    dwconv7_fwd_x2_kernel<3><<<dim3(gx, chunks), NTHREADS, smem, st>>>(tmX, w49c, bias, (const bf16*)res, (bf16*)y, B, H, W, C, tiles_w, tiles_h, wl);
  else
    dwconv7_fwd_x2_kernel<2><<<dim3(gx, chunks), NTHREADS, smem, st>>>(tmX, w49c, bias, (const bf16*)res, (bf16*)y, B, H, W, C, tiles_w, tiles_h, wl);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

int lnx_dwconv7_wgrad_bf16(const void* x, const void* dy, float* dw49c, int wl, float* dbias, int B, int H, int W, int C, cudaStream_t st) {
  if (C % CC != 0) return LNX_ERR_SHAPE;
  if (lnx_dwconv7_mma_enabled()) {
    const int rc = lnx_dwconv7_wgrad_mma(x, dy, dw49c, wl, dbias, B, H, W, C, st);
    if (rc != LNX_ERR_UNSUPPORTED) return rc;
  }
  CUtensorMap tmX, tmG;
  if (!make_nhwc_tmap(&tmX, x, B, H, W, C, HALO, HALO) || !make_nhwc_tmap(&tmG, dy, B, H, W, C, TILE, TILE)) return LNX_ERR_UNSUPPORTED;
  const int tiles_w = (W + TILE - 1) / TILE, tiles_h = (H + TILE - 1) / TILE;
  const size_t smem = 2 * (HALO_BYTES + CENTER_PAD) + 50 * CC * 4 + 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dwconv7_wgrad_x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    attr_set = true;
  }
  const int chunks = C / CC;
  const int total = B * tiles_h * tiles_w;
  const int gx = max(1, min(total, (kNumSMs * 2 + chunks - 1) / chunks));
  dwconv7_wgrad_x2_kernel<<<dim3(gx, chunks), NTHREADS, smem, st>>>(tmX, tmG, dw49c, dbias, B, H, W, C, tiles_w, tiles_h, wl);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
