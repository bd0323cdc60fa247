// Depthwise 7x7 (pad 3, stride 1) on NHWC, forward / data-gradient and weight-gradient.
//
// CTA = one 14x14 output tile x one chunk of channels of one image.  The 20x20
// input halo tile is staged once in shared memory (16-byte coalesced NHWC loads,
// zero filled outside the image); a lane owns CPL channels (1 or 2), a warp owns
// two tile rows, and each thread walks full 14-wide output rows so that every value
// read from shared memory feeds up to 7 FMAs (the kernel is CUDA-core-FMA bound,
// not HBM bound: 49 MACs per output element at 4 bytes of traffic).
#include "lnx_common.cuh"

using namespace lnx;

namespace {

constexpr int TILE = 14;
constexpr int HALO = TILE + 6;  // 20
constexpr int NWARPS = 7;       // warp w -> tile rows 2w, 2w+1

template <typename T, int CPL>
struct Chan;  // CPL channels of one pixel as fp32
template <>
struct Chan<float, 1> {
  static __device__ __forceinline__ void ld(const float* p, float* v) { v[0] = p[0]; }
};
template <>
struct Chan<float, 2> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  }
};
template <>
struct Chan<bf16, 1> {
  static __device__ __forceinline__ void ld(const bf16* p, float* v) { v[0] = __bfloat162float(p[0]); }
};
template <>
struct Chan<bf16, 2> {
  static __device__ __forceinline__ void ld(const bf16* p, float* v) {
    const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(p);
    v[0] = __bfloat162float(t.x); v[1] = __bfloat162float(t.y);
  }
};

template <typename T>
__device__ __forceinline__ void load_tile(T* tile, const T* __restrict__ src, int b, int h0, int w0, int c0, int H, int W, int C,
                                          int CC, int rows, int cols, int roff, int coff) {
  // tile[rows][cols][CC]; source pixel (h0 - roff + r, w0 - coff + c)
  constexpr int V = Vec16<T>::N;
  const int vec_per_px = CC / V;
  const int total = rows * cols * vec_per_px;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int v = i % vec_per_px;
    const int px = i / vec_per_px;
    const int r = px / cols, c = px % cols;
    const int h = h0 - roff + r, w = w0 - coff + c;
    Vec16<T> val;
    if (h >= 0 && h < H && w >= 0 && w < W) {
      val = ld16(src + (((long long)b * H + h) * W + w) * C + c0 + v * V);
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) val.set(j, 0.f);
    }
    st16(tile + (long long)px * CC + v * V, val);
  }
}

// weight element (tap, channel) in the caller's layout (see include/linnaeus_b200.h: LNX_DW_W_*)
__device__ __forceinline__ long long widx(int wl, int tap, int c, int C) {
  return wl == 0 ? (long long)tap * C + c : (long long)c * 49 + (wl == 2 ? 48 - tap : tap);
}

// ------------------------------------------------------------------ forward
template <typename T, int CPL>
__global__ void __launch_bounds__(NWARPS * 32) dwconv7_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w49c,
                                                                   const float* __restrict__ bias, const T* __restrict__ res,
                                                                   T* __restrict__ y, int B, int H, int W, int C, int tiles_w,
                                                                   int tiles_h, int wl) {
  constexpr int CC = 32 * CPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);                                     // [20][20][CC]
  float* wsm = reinterpret_cast<float*>(smem_raw + sizeof(T) * HALO * HALO * CC);  // [49][CC]

  const int c0 = blockIdx.y * CC;
  int t = blockIdx.x;
  const int tw = t % tiles_w; t /= tiles_w;
  const int th = t % tiles_h;
  const int b = t / tiles_h;
  const int h0 = th * TILE, w0 = tw * TILE;

  for (int i = threadIdx.x; i < 49 * CC; i += blockDim.x) wsm[i] = w49c[widx(wl, i / CC, c0 + (i % CC), C)];
  load_tile<T>(tile, x, b, h0, w0, c0, H, W, C, CC, HALO, HALO, 3, 3);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cl = lane * CPL;
  float bv[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) bv[q] = bias ? bias[c0 + cl + q] : 0.f;

  // Each thread owns 2 tile rows x 14 columns of its CPL channels.  kh is the outer loop so the 7 taps of a
  // filter row are fetched once and reused for both output rows: 47 shared loads per 196*CPL FMAs.
  const int orow0 = warp * 2;
  if (h0 + orow0 < H) {
    float acc[2][TILE][CPL];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int o = 0; o < TILE; ++o)
#pragma unroll
        for (int q = 0; q < CPL; ++q) acc[rr][o][q] = bv[q];
#pragma unroll
    for (int kh = 0; kh < 7; ++kh) {
      float wk[7][CPL];
#pragma unroll
      for (int kw = 0; kw < 7; ++kw)
#pragma unroll
        for (int q = 0; q < CPL; ++q) wk[kw][q] = wsm[(kh * 7 + kw) * CC + cl + q];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const T* rowp = tile + ((orow0 + rr + kh) * HALO) * CC + cl;
#pragma unroll
        for (int ix = 0; ix < HALO; ++ix) {
          float v[CPL];
          Chan<T, CPL>::ld(rowp + ix * CC, v);
#pragma unroll
          for (int kw = 0; kw < 7; ++kw) {
            const int o = ix - kw;
            if (o >= 0 && o < TILE) {
#pragma unroll
              for (int q = 0; q < CPL; ++q) acc[rr][o][q] = fmaf(v[q], wk[kw][q], acc[rr][o][q]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int hh = h0 + orow0 + rr;
      if (hh >= H) continue;
#pragma unroll
      for (int o = 0; o < TILE; ++o) {
        const int ww = w0 + o;
        if (ww < W) {
          const long long off = (((long long)b * H + hh) * W + ww) * C + c0 + cl;
          T* dst = y + off;
          if (res) {
#pragma unroll
            for (int q = 0; q < CPL; ++q) acc[rr][o][q] += to_f32(res[off + q]);
          }
          if constexpr (CPL == 2) {
            if constexpr (sizeof(T) == 2) {
              *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(acc[rr][o][0], acc[rr][o][1]);
            } else {
              *reinterpret_cast<float2*>(dst) = make_float2(acc[rr][o][0], acc[rr][o][1]);
            }
          } else {
            dst[0] = from_f32<T>(acc[rr][o][0]);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ weight gradient
// grid.x CTAs stride over (image, tile) pairs of one channel chunk, keeping the 49 x CPL
// partial sums of their lanes in registers; one shared + global atomic pass at the end.
template <typename T, int CPL>
__global__ void __launch_bounds__(NWARPS * 32) dwconv7_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                     float* __restrict__ dw49c, float* __restrict__ dbias, int B, int H,
                                                                     int W, int C, int tiles_w, int tiles_h, int wl) {
  constexpr int CC = 32 * CPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);                       // [20][20][CC]
  T* gt = tile + HALO * HALO * CC;                                // [14][14][CC]
  float* red = reinterpret_cast<float*>(gt + TILE * TILE * CC);   // [50][CC]

  const int c0 = blockIdx.y * CC;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cl = lane * CPL;
  float wacc[49][CPL];
  float bacc[CPL];
#pragma unroll
  for (int k = 0; k < 49; ++k)
#pragma unroll
    for (int q = 0; q < CPL; ++q) wacc[k][q] = 0.f;
#pragma unroll
  for (int q = 0; q < CPL; ++q) bacc[q] = 0.f;
  for (int i = threadIdx.x; i < 50 * CC; i += blockDim.x) red[i] = 0.f;

  const int total_tiles = B * tiles_h * tiles_w;
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int tw = t % tiles_w;
    const int th = (t / tiles_w) % tiles_h;
    const int b = t / (tiles_w * tiles_h);
    const int h0 = th * TILE, w0 = tw * TILE;
    __syncthreads();
    load_tile<T>(tile, x, b, h0, w0, c0, H, W, C, CC, HALO, HALO, 3, 3);
    load_tile<T>(gt, dy, b, h0, w0, c0, H, W, C, CC, TILE, TILE, 0, 0);  // zero outside the image
    __syncthreads();
#pragma unroll 1
    for (int rr = 0; rr < 2; ++rr) {  // full 14-wide rows: 20 tile loads feed 98*CPL FMAs per filter row
      const int orow = warp * 2 + rr;
      if (h0 + orow >= H) continue;
      float g[TILE][CPL];
#pragma unroll
      for (int o = 0; o < TILE; ++o) {
        Chan<T, CPL>::ld(gt + (orow * TILE + o) * CC + cl, g[o]);
#pragma unroll
        for (int q = 0; q < CPL; ++q) bacc[q] += g[o][q];
      }
#pragma unroll
      for (int kh = 0; kh < 7; ++kh) {
        const T* rowp = tile + ((orow + kh) * HALO) * CC + cl;
#pragma unroll
        for (int ix = 0; ix < HALO; ++ix) {
          float v[CPL];
          Chan<T, CPL>::ld(rowp + ix * CC, v);
#pragma unroll
          for (int kw = 0; kw < 7; ++kw) {
            const int o = ix - kw;
            if (o >= 0 && o < TILE) {
#pragma unroll
              for (int q = 0; q < CPL; ++q) wacc[kh * 7 + kw][q] = fmaf(v[q], g[o][q], wacc[kh * 7 + kw][q]);
            }
          }
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 49; ++k)
#pragma unroll
    for (int q = 0; q < CPL; ++q) atomicAdd(&red[k * CC + cl + q], wacc[k][q]);
#pragma unroll
  for (int q = 0; q < CPL; ++q) atomicAdd(&red[49 * CC + cl + q], bacc[q]);
  __syncthreads();
  for (int i = threadIdx.x; i < 49 * CC; i += blockDim.x) atomicAdd(dw49c + widx(wl, i / CC, c0 + (i % CC), C), red[i]);
  if (dbias)
    for (int i = threadIdx.x; i < CC; i += blockDim.x) atomicAdd(dbias + c0 + i, red[49 * CC + i]);
}

template <typename T, int CPL>
int fwd_launch(const void* x, const float* w49c, int wl, const float* bias, const void* res, void* y, int B, int H, int W, int C, cudaStream_t st) {
  constexpr int CC = 32 * CPL;
  const int tiles_w = (W + TILE - 1) / TILE, tiles_h = (H + TILE - 1) / TILE;
  const size_t smem = sizeof(T) * HALO * HALO * CC + sizeof(float) * 49 * CC;
  auto kern = dwconv7_fwd_kernel<T, CPL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return lnx_set_cuda_error(e);
  dim3 grid(B * tiles_h * tiles_w, C / CC);
  kern<<<grid, NWARPS * 32, smem, st>>>((const T*)x, w49c, bias, (const T*)res, (T*)y, B, H, W, C, tiles_w, tiles_h, wl);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

template <typename T, int CPL>
int wgrad_launch(const void* x, const void* dy, float* dw, int wl, float* db, int B, int H, int W, int C, cudaStream_t st) {
  constexpr int CC = 32 * CPL;
  const int tiles_w = (W + TILE - 1) / TILE, tiles_h = (H + TILE - 1) / TILE;
  const size_t smem = sizeof(T) * (HALO * HALO + TILE * TILE) * CC + sizeof(float) * 50 * CC;
  auto kern = dwconv7_wgrad_kernel<T, CPL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return lnx_set_cuda_error(e);
  const int chunks = C / CC;
  const int total_tiles = B * tiles_h * tiles_w;
  int gx = max(1, min(total_tiles, (kNumSMs * 2 + chunks - 1) / chunks));
  dim3 grid(gx, chunks);
  kern<<<grid, NWARPS * 32, smem, st>>>((const T*)x, (const T*)dy, dw, db, B, H, W, C, tiles_w, tiles_h, wl);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

}  // namespace

// bf16: packed-fp32x2 (FFMA2) kernels of lnx_dwconv_bf16.cu
int lnx_dwconv7_fwd_bf16(const void* x, const float* w49c, int wl, const float* bias, const void* res, void* y, int B, int H, int W, int C,
                         cudaStream_t st);
int lnx_dwconv7_wgrad_bf16(const void* x, const void* dy, float* dw49c, int wl, float* dbias, int B, int H, int W, int C, cudaStream_t st);

extern "C" int lnx_dwconv7_fwd(const void* x, const float* w, int w_layout, const float* bias, const void* residual, void* y, int B, int H, int W,
                               int C, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(x && w && y, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 32 == 0 && w_layout >= 0 && w_layout <= 2, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(x) && lnx_aligned16(y), LNX_ERR_ALIGN);
  cudaStream_t st = (cudaStream_t)s;
  const bool pair = (C % 64 == 0);
  if (dtype == LNX_F32)
    return pair ? fwd_launch<float, 2>(x, w, w_layout, bias, residual, y, B, H, W, C, st)
                : fwd_launch<float, 1>(x, w, w_layout, bias, residual, y, B, H, W, C, st);
  if (dtype == LNX_BF16) return lnx_dwconv7_fwd_bf16(x, w, w_layout, bias, residual, y, B, H, W, C, st);
  return LNX_ERR_DTYPE;
}

extern "C" int lnx_dwconv7_wgrad(const void* x, const void* dy, float* dw, int w_layout, float* dbias, int B, int H, int W, int C, int dtype,
                                 lnx_stream_t s) {
  LNX_REQUIRE(x && dy && dw, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 32 == 0 && (w_layout == 0 || w_layout == 1), LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(x) && lnx_aligned16(dy), LNX_ERR_ALIGN);
  cudaStream_t st = (cudaStream_t)s;
  const bool pair = (C % 64 == 0);
  if (dtype == LNX_F32)
    return pair ? wgrad_launch<float, 2>(x, dy, dw, w_layout, dbias, B, H, W, C, st) : wgrad_launch<float, 1>(x, dy, dw, w_layout, dbias, B, H, W, C, st);
  if (dtype == LNX_BF16) return lnx_dwconv7_wgrad_bf16(x, dy, dw, w_layout, dbias, B, H, W, C, st);
  return LNX_ERR_DTYPE;
}
