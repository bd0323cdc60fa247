// Depthwise 7x7 (pad 3, stride 1) on NHWC bf16 on the TENSOR pipe: forward / data gradient and weight gradient as banded
// (Toeplitz) matrix products on mma.sync.m16n8k16 (bf16 x bf16 -> fp32), replacing the FFMA2 kernels of lnx_dwconv_bf16.cu
// (which stay as the fallback for wide images and as the A/B reference: LNX_DWCONV_MMA=0).
//
// Why the register-fragment MMA and not tcgen05: a depthwise conv has no contraction over channels, so every channel is its
// own small matrix problem whose operands are channel-PLANAR (two neighbouring pixels of one channel in one 32-bit register),
// while the activations are NHWC (two neighbouring channels of one pixel per word).  tcgen05 reads its operands from shared
// memory in canonical layouts that no TMA box can produce from NHWC without a 2-byte transposition pass through registers;
// mma.sync takes the operands FROM registers, where that transposition is one PRMT per register (two words of two pixels ->
// the pixel pair of channel 2q and of channel 2q+1), and its legacy-pipe rate (~1/4 of tcgen05) is still 4x what the op needs
// once it is off the fp32 FMA pipe: the kernels become shared-memory / HBM bound.
//
// Forward, one channel c, one filter row ky, 16 output rows x 8 output columns per MMA:
//     out[y][8j + n] += sum_k in_h[y + ky][8j + xo(k)] * Wband[k][n],   Wband[k][n] = w[ky][xo(k) - n] (0 outside 0..6)
// in_h = the zero-padded (halo) input, xo = the K = 16 window of input columns the 8 outputs touch (8 + 6 = 14 <= 16).
// The MMA rows are mapped m = g -> tile row 2g, m = g + 8 -> tile row 2g + 1, so the A registers of filter row ky + 1
// (rows "g + 8") are the registers of filter row ky + 2 (rows "g"): a thread loads 8 input rows for 7 MMAs.  The K index is
// permuted (k = 2t, 2t+1, 2t+8, 2t+9 -> columns t, t+4, t+8, t+12) so that the four lanes of a quad read four CONSECUTIVE
// pixels: with the 16-byte-chunk XOR swizzle below, every 8-lane phase of the 128-bit shared loads is conflict free.
// 7 MMAs of 2048 MACs do 896 useful ones (2.3x inflation, not the 9x of a full-row Toeplitz matrix).
//
// Weight gradient, one channel, TWO padded input rows yi, yi + 1 and 16 output columns per MMA:
//     A[m = kx][k = x] = in_h[yi][x + kx]  (rows 8..15: the same for row yi + 1),   B[k = x][n] = dy[yi - n][x]  (n = 7: dy[yi + 1][x])
//     D[kx][n]      -> dW[ky = n][kx]      (n <= 6),        D[8 + kx][n] -> dW[ky = n + 1][kx] (n <= 5),  D[8 + kx][7] -> dW[0][kx]
// i.e. all 49 taps of both input rows (98 of 128 outputs useful); row m = 7 of A is all ones, so D[7][0] + D[7][7] = the bias gradient.
// Accumulators stay in registers across every tile a persistent CTA visits.
#include <stdlib.h>

#include "lnx_common.cuh"

using namespace lnx;

namespace {

constexpr int CC = 32;            // channels per CTA (one 64-byte pixel row of the shared tile)
constexpr int TROWS = 16;         // forward: output rows per tile
constexpr int IN_ROWS = TROWS + 6;
constexpr int BTAB_ENTRIES = CC * 7 * 11;
constexpr int BTAB_BYTES = BTAB_ENTRIES * 8;  // 19712 = 154 x 128
constexpr int FWD_SLACK_PX = 16;
constexpr int WG_ROWS = 8;        // weight gradient: padded input rows per tile (4 row pairs)
constexpr int WG_GROWS = WG_ROWS + 6;
constexpr int WG_SLACK_PX = 16;
constexpr int WG_RED_BYTES = 50 * CC * 4;  // 6400 = 50 x 128
constexpr int WG_WARPS = 8;

__device__ __forceinline__ long long widx(int wl, int tap, int c, int C) {
  return wl == 0 ? (long long)tap * C + c : (long long)c * 49 + (wl == 2 ? 48 - tap : tap);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// byte offset of the 16-byte chunk cq (8 channels) of pixel p in a [pixel][32 channels] bf16 tile.  The chunk index is XORed with
// bits 1-2 of the pixel index: eight lanes reading the same chunk of pixels that are distinct mod 8 hit eight distinct 16-byte bank groups.
__device__ __forceinline__ uint32_t swz(int p, int cq) { return (uint32_t)p * 64u + (uint32_t)((cq ^ ((p >> 1) & 3)) << 4); }
// words of two pixels (channel pair q of each) -> {pixel a, pixel b} of the even channel / of the odd channel
__device__ __forceinline__ uint32_t pair_lo(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x5410); }
__device__ __forceinline__ uint32_t pair_hi(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x7632); }

__device__ __forceinline__ void planar8(const uint4& a, const uint4& b, uint32_t (&out)[8]) {
  out[0] = pair_lo(a.x, b.x);
  out[1] = pair_hi(a.x, b.x);
  out[2] = pair_lo(a.y, b.y);
  out[3] = pair_hi(a.y, b.y);
  out[4] = pair_lo(a.z, b.z);
  out[5] = pair_hi(a.z, b.z);
  out[6] = pair_lo(a.w, b.w);
  out[7] = pair_hi(a.w, b.w);
}

// ------------------------------------------------------------------ forward (and data gradient with flipped taps)
// grid = (B * tiles_h * tiles_w, C / 32); CTA = one 16-row x WT-column output tile of one image x 32 channels; two CTAs per SM
// cover each other's load / store phases.  smem: band table [32][7][11] x 8 B | halo tile [22][RP] pixels x 64 B, swizzled.
template <int NW>
__global__ void __launch_bounds__(NW * 32, 2)
    dwconv7_fwd_mma_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, const bf16* __restrict__ res,
                           bf16* __restrict__ y, int B, int H, int W, int C, int tiles_w, int tiles_h, int WT, int RP, int nxb, int wl) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint2* btab = reinterpret_cast<uint2*>(smem);
  unsigned char* tile = smem + BTAB_BYTES;
  const uint32_t tile_s = smem_u32(tile);

  const int c0 = blockIdx.y * CC;
  int tt = blockIdx.x;
  const int x0 = (tt % tiles_w) * WT;
  tt /= tiles_w;
  const int y0 = (tt % tiles_h) * TROWS;
  const int b = tt / tiles_h;

  // halo tile: image rows y0-3 .. y0+18, columns x0-3 .. x0+RP-4; everything outside the image is zero filled (= the conv padding)
  const int npix = IN_ROWS * RP;
  for (int idx = threadIdx.x; idx < npix * 4; idx += NW * 32) {
    const int cq = idx & 3, p = idx >> 2;
    const int r = p / RP, xh = p - r * RP;
    const int yi = y0 + r - 3, xi = x0 + xh - 3;
    const bool ok = yi >= 0 && yi < H && xi >= 0 && xi < W;
    const bf16* src = ok ? x + ((((long long)b * H + yi) * W + xi) * C + c0 + cq * 8) : x;
    cp_async16(tile_s + swz(p, cq), src, ok ? 16 : 0);
  }
  // the K window of the last column block reaches a few pixels past the row end (into the next row / this slack): those
  // products meet zero band entries, the data only has to be finite
  for (int idx = threadIdx.x; idx < FWD_SLACK_PX * 4; idx += NW * 32) reinterpret_cast<uint4*>(tile + (size_t)npix * 64)[idx] = make_uint4(0, 0, 0, 0);
  // band table: entry (c, ky, d + 7), d = t - g in -7..3: b0 = {w[d], w[d+4]}, b1 = {w[d+8], w[d+12]} (bf16, zero outside 0..6)
  for (int i = threadIdx.x; i < BTAB_ENTRIES; i += NW * 32) {
    const int e = i % 11, ky = (i / 11) % 7, c = i / 77;
    const int d = e - 7;
    uint32_t v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int kx = d + 4 * q;
      v[q] = (kx >= 0 && kx <= 6) ? (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(w[widx(wl, ky * 7 + kx, c0 + c, C)])) : 0u;
    }
    btab[i] = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
  }
  cp_async_wait_all();
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int e = t - g + 7;
  const int xlim = min(W, x0 + WT);
  for (int u = warp; u < nxb * 4; u += NW) {
    const int j = u % nxb, cg = u / nxb;
    float acc[8][4];
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      const float bv = bias ? bias[c0 + cg * 8 + ch] : 0.f;
      acc[ch][0] = acc[ch][1] = acc[ch][2] = acc[ch][3] = bv;
    }
    const uint2* bt = btab + (cg * 8) * 77 + e;
    uint32_t plo[8], phi[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int p0 = (2 * g + r) * RP + 8 * j + t;
      const uint4 q0 = lds128(tile_s + swz(p0, cg));
      const uint4 q1 = lds128(tile_s + swz(p0 + 4, cg));
      const uint4 q2 = lds128(tile_s + swz(p0 + 8, cg));
      const uint4 q3 = lds128(tile_s + swz(p0 + 12, cg));
      uint32_t clo[8], chi[8];
      planar8(q0, q1, clo);
      planar8(q2, q3, chi);
      if (r > 0) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint2 bb = bt[ch * 77 + (r - 1) * 11];
          mma16816(acc[ch], plo[ch], clo[ch], phi[ch], chi[ch], bb.x, bb.y);
        }
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        plo[ch] = clo[ch];
        phi[ch] = chi[ch];
      }
    }
    // thread (g, t): rows 2g, 2g+1 x columns 8j+2t, 8j+2t+1 x 8 channels = four 16-byte stores
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int yy = y0 + 2 * g + hh;
      if (yy >= H) continue;
#pragma unroll
      for (int xx = 0; xx < 2; ++xx) {
        const int xc = x0 + 8 * j + 2 * t + xx;
        if (xc >= xlim) continue;
        const long long off = (((long long)b * H + yy) * W + xc) * C + c0 + cg * 8;
        float v[8];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) v[ch] = acc[ch][2 * hh + xx];
        if (res) {  // fused "+ residual" (the skip-connection gradient when this kernel runs as the data gradient)
          const uint4 rv = __ldg(reinterpret_cast<const uint4*>(res + off));
          const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            v[2 * q] += __uint_as_float(rw[q] << 16);
            v[2 * q + 1] += __uint_as_float(rw[q] & 0xffff0000u);
          }
        }
        uint4 o;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
        o.x = *reinterpret_cast<uint32_t*>(&h0);
        o.y = *reinterpret_cast<uint32_t*>(&h1);
        o.z = *reinterpret_cast<uint32_t*>(&h2);
        o.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(y + off) = o;
      }
    }
  }
}

// ------------------------------------------------------------------ weight gradient
// grid = (gx, C / 32), persistent over tiles = (image, band of 8 padded input rows); 8 warps: warp & 3 = channel group of 8,
// warp >> 2 = which two of the band's four row pairs.  smem: red [50][32] f32 | input band [8][RP] (+ slack) | dy band [14][RPG].
__global__ void __launch_bounds__(WG_WARPS * 32, 2)
    dwconv7_wgrad_mma_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw, float* __restrict__ dbias, int B, int H,
                             int W, int C, int tiles_h, int RP, int RPG, int nxc, int wl) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* red = reinterpret_cast<float*>(smem);
  unsigned char* xin = smem + WG_RED_BYTES;
  const int in_px = WG_ROWS * RP;
  unsigned char* gin = xin + (size_t)(in_px + WG_SLACK_PX) * 64;
  const uint32_t xin_s = smem_u32(xin), gin_s = smem_u32(gin);
  constexpr int NT = WG_WARPS * 32;

  const int c0 = blockIdx.y * CC;
  const int total = B * tiles_h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int cg = warp & 3, half = warp >> 2;

  for (int i = threadIdx.x; i < 50 * CC; i += NT) red[i] = 0.f;
  for (int i = threadIdx.x; i < WG_SLACK_PX * 4; i += NT) reinterpret_cast<uint4*>(xin + (size_t)in_px * 64)[i] = make_uint4(0, 0, 0, 0);

  float acc[8][4];
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) acc[ch][0] = acc[ch][1] = acc[ch][2] = acc[ch][3] = 0.f;

  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int b = tile / tiles_h;
    const int yi0 = (tile - b * tiles_h) * WG_ROWS;  // first padded input row of the band
    __syncthreads();                                 // everyone is done with the previous band
    for (int idx = threadIdx.x; idx < in_px * 4; idx += NT) {
      const int cq = idx & 3, p = idx >> 2;
      const int r = p / RP, xh = p - r * RP;
      const int yi = yi0 + r - 3, xi = xh - 3;
      const bool ok = yi >= 0 && yi < H && xi >= 0 && xi < W;
      const bf16* src = ok ? x + ((((long long)b * H + yi) * W + xi) * C + c0 + cq * 8) : x;
      cp_async16(xin_s + swz(p, cq), src, ok ? 16 : 0);
    }
    for (int idx = threadIdx.x; idx < WG_GROWS * RPG * 4; idx += NT) {
      const int cq = idx & 3, p = idx >> 2;
      const int r = p / RPG, xx = p - r * RPG;
      const int yd = yi0 - 6 + r;
      const bool ok = yd >= 0 && yd < H && xx < W;
      const bf16* src = ok ? dy + ((((long long)b * H + yd) * W + xx) * C + c0 + cq * 8) : dy;
      cp_async16(gin_s + swz(p, cq), src, ok ? 16 : 0);
    }
    cp_async_wait_all();
    __syncthreads();

#pragma unroll 1
    for (int rr = 0; rr < 2; ++rr) {
      const int rp = half * 2 + rr;
      const int rowb = (g < 7) ? (2 * rp + 6 - g) : (2 * rp + 7);
#pragma unroll 1
      for (int xc = 0; xc < nxc; ++xc) {
        const int pa = (2 * rp) * RP + 16 * xc + t + g;
        const int pb = rowb * RPG + 16 * xc + t;
        uint32_t a0[8], a1[8], a2[8], a3[8], b0[8], b1[8];
        {
          const uint4 q0 = lds128(xin_s + swz(pa, cg)), q1 = lds128(xin_s + swz(pa + 4, cg));
          const uint4 q2 = lds128(xin_s + swz(pa + 8, cg)), q3 = lds128(xin_s + swz(pa + 12, cg));
          planar8(q0, q1, a0);
          planar8(q2, q3, a2);
        }
        {
          const uint4 q0 = lds128(xin_s + swz(pa + RP, cg)), q1 = lds128(xin_s + swz(pa + RP + 4, cg));
          const uint4 q2 = lds128(xin_s + swz(pa + RP + 8, cg)), q3 = lds128(xin_s + swz(pa + RP + 12, cg));
          planar8(q0, q1, a1);
          planar8(q2, q3, a3);
        }
        {
          const uint4 q0 = lds128(gin_s + swz(pb, cg)), q1 = lds128(gin_s + swz(pb + 4, cg));
          const uint4 q2 = lds128(gin_s + swz(pb + 8, cg)), q3 = lds128(gin_s + swz(pb + 12, cg));
          planar8(q0, q1, b0);
          planar8(q2, q3, b1);
        }
        if (g == 7) {  // row m = 7 of A: ones -> column sums of dy (bias gradient)
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) a0[ch] = a2[ch] = 0x3F803F80u;
        }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) mma16816(acc[ch], a0[ch], a1[ch], a2[ch], a3[ch], b0[ch], b1[ch]);
      }
    }
  }

  // c0, c1 = D[g][2t], D[g][2t+1];  c2, c3 = D[g+8][2t], D[g+8][2t+1]
  __syncthreads();
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    const int c = cg * 8 + ch;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int n = 2 * t + q;
      if (g < 7) {
        if (n <= 6) atomicAdd(&red[(n * 7 + g) * CC + c], acc[ch][q]);           // dW[ky = n][kx = g] from row yi
        if (n <= 5) atomicAdd(&red[((n + 1) * 7 + g) * CC + c], acc[ch][2 + q]);  // dW[ky = n + 1][kx = g] from row yi + 1
        if (n == 7) atomicAdd(&red[g * CC + c], acc[ch][2 + q]);                  // dW[0][kx = g] from row yi + 1
      } else if (n == 0 || n == 7) {
        atomicAdd(&red[49 * CC + c], acc[ch][q]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 49 * CC; i += NT) atomicAdd(dw + widx(wl, i / CC, c0 + (i % CC), C), red[i]);
  if (dbias)
    for (int i = threadIdx.x; i < CC; i += NT) atomicAdd(dbias + c0 + i, red[49 * CC + i]);
}

int round_up_mod(int v, int m, int r) {  // smallest value >= v that is r mod m
  int o = v - ((v - r) % m + m) % m;
  return o < v ? o + m : o;
}

}  // namespace

bool lnx_dwconv7_mma_enabled() {
  static const int en = getenv("LNX_DWCONV_MMA") ? atoi(getenv("LNX_DWCONV_MMA")) : 1;
  return en != 0;
}

// -> LNX_OK, or LNX_ERR_UNSUPPORTED when the caller should use the FFMA2 kernels
int lnx_dwconv7_fwd_mma(const void* x, const float* w, int wl, const float* bias, const void* res, void* y, int B, int H, int W, int C,
                        cudaStream_t st) {
  if (C % CC != 0) return LNX_ERR_SHAPE;
  const int tiles_w = (W + 55) / 56;
  const int WT = tiles_w == 1 ? W : ((W + tiles_w - 1) / tiles_w + 7) / 8 * 8;
  const int RP = round_up_mod(WT + 6, 4, 2);  // 2 mod 4: two tile rows apart = 4 pixels mod 8 (bank-conflict-free quads)
  const int nxb = (WT + 7) / 8;
  const int tiles_h = (H + TROWS - 1) / TROWS;
  const size_t smem = BTAB_BYTES + (size_t)(IN_ROWS * RP + FWD_SLACK_PX) * 64;
  if (smem > 113 * 1024) return LNX_ERR_UNSUPPORTED;
  const bool seven = (nxb * 4) % 7 == 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dwconv7_fwd_mma_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv7_fwd_mma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    attr_set = true;
  }
  const dim3 grid(B * tiles_h * tiles_w, C / CC);
  if (seven)
    dwconv7_fwd_mma_kernel<7><<<grid, 7 * 32, smem, st>>>((const bf16*)x, w, bias, (const bf16*)res, (bf16*)y, B, H, W, C, tiles_w, tiles_h, WT, RP, nxb, wl);
  else
    dwconv7_fwd_mma_kernel<8><<<grid, 8 * 32, smem, st>>>((const bf16*)x, w, bias, (const bf16*)res, (bf16*)y, B, H, W, C, tiles_w, tiles_h, WT, RP, nxb, wl);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

int lnx_dwconv7_wgrad_mma(const void* x, const void* dy, float* dw, int wl, float* dbias, int B, int H, int W, int C, cudaStream_t st) {
  if (C % CC != 0) return LNX_ERR_SHAPE;
  if (W > 64) return LNX_ERR_UNSUPPORTED;
  const int nxc = (W + 15) / 16;
  const int RP = W + 6;
  const int RPG = round_up_mod(16 * nxc, 8, 4);  // 4 mod 8: neighbouring dy rows = the other four pixels mod 8
  const int tiles_h = (H + 6 + WG_ROWS - 1) / WG_ROWS;
  const size_t smem = WG_RED_BYTES + (size_t)(WG_ROWS * RP + WG_SLACK_PX) * 64 + (size_t)WG_GROWS * RPG * 64;
  if (smem > 113 * 1024) return LNX_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dwconv7_wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    attr_set = true;
  }
  const int chunks = C / CC;
  const int total = B * tiles_h;
  const int gx = max(1, min(total, (kNumSMs * 2 + chunks - 1) / chunks));
  dwconv7_wgrad_mma_kernel<<<dim3(gx, chunks), WG_WARPS * 32, smem, st>>>((const bf16*)x, (const bf16*)dy, dw, dbias, B, H, W, C, tiles_h, RP, RPG, nxc, wl);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
