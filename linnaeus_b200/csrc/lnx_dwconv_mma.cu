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
// i.e. all 49 taps of both input rows (98 of 128 outputs useful); row m = 7 of A is all ones, so D[7][2] + D[7][3] (dy rows yi - 2, yi - 3 = the
// image rows under the two padded input rows, which are stepped over the image rows only) = the bias gradient.
// Accumulators stay in registers across every tile a persistent CTA visits.
#include <stdlib.h>

#include "lnx_common.cuh"
#include "lnx_tc_common.cuh"

using namespace lnx;
using namespace lnx_tc;

namespace {

constexpr int CC = 32;            // channels per work item (one 64-byte pixel row of the shared tile)
constexpr int TROWS = 16;         // forward: output rows per tile
constexpr int IN_ROWS = TROWS + 6;
constexpr int BTAB_ENTRIES = CC * 7 * 11;
constexpr int BTAB_BYTES = BTAB_ENTRIES * 8;  // 19712
constexpr int SLACK_BYTES = 1024;             // 16 pixels behind a tile: the K windows of the last column block end there
// weight gradient: padded input rows per band = 8 (4 row pairs, one per warp quarter) or, when two stages of the wider band still fit in
// shared memory (W <= 28), 16 (two row pairs per warp: half the per-item overhead, 22 / 16 instead of 14 / 8 dy rows fetched per band)
constexpr int WG_RED_BYTES = 50 * CC * 4;  // 6400
constexpr int WG_WARPS = 16;

__device__ __forceinline__ long long widx(int wl, int tap, int c, int C) {
  return wl == 0 ? (long long)tap * C + c : (long long)c * 49 + (wl == 2 ? 48 - tap : tap);
}
template <int OFF>
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(OFF) : "memory");
  return v;
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// Shared tiles are [pixel][32 channels] bf16 written by TMA with the 64-byte swizzle: the 16-byte chunk index of a pixel is XORed with
// address bits 7-8 = bits 1-2 of the pixel index (tiles are 512-byte aligned).  Eight lanes reading the same chunk of pixels that are
// distinct mod 8 therefore hit eight distinct 16-byte bank groups.
__device__ __forceinline__ uint32_t swz(int p, int cq) { return (uint32_t)p * 64u + (uint32_t)((cq ^ ((p >> 1) & 3)) << 4); }
// the four pixels p, p+4, p+8, p+12 of a K window: p+4 / p+12 flip bit 1 of the chunk index (= address bit 5), p+8 keeps it
struct Quad {
  uint4 q0, q1, q2, q3;
};
__device__ __forceinline__ Quad lds_quad(uint32_t a) {
  Quad q;
  q.q0 = lds128<0>(a);
  q.q1 = lds128<256>(a ^ 32u);
  q.q2 = lds128<512>(a);
  q.q3 = lds128<768>(a ^ 32u);
  return q;
}
// words of two pixels (channel pair q of each) -> {pixel a, pixel b} of the even channel / of the odd channel
__device__ __forceinline__ void planar8(const uint4& a, const uint4& b, uint32_t (&out)[8]) {
  out[0] = __byte_perm(a.x, b.x, 0x5410);
  out[1] = __byte_perm(a.x, b.x, 0x7632);
  out[2] = __byte_perm(a.y, b.y, 0x5410);
  out[3] = __byte_perm(a.y, b.y, 0x7632);
  out[4] = __byte_perm(a.z, b.z, 0x5410);
  out[5] = __byte_perm(a.z, b.z, 0x7632);
  out[6] = __byte_perm(a.w, b.w, 0x5410);
  out[7] = __byte_perm(a.w, b.w, 0x7632);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// work item -> (chunk, image, tile row, tile column), advanced without divisions
struct Item {
  int chunk, b, th, tw;
};
__device__ __forceinline__ Item item_of(int item, int tiles_w, int tiles_h, int B) {
  Item c;
  c.tw = item % tiles_w;
  item /= tiles_w;
  c.th = item % tiles_h;
  item /= tiles_h;
  c.b = item % B;
  c.chunk = item / B;
  return c;
}
__device__ __forceinline__ void item_next(Item& c, int tiles_w, int tiles_h, int B) {
  if (++c.tw == tiles_w) {
    c.tw = 0;
    if (++c.th == tiles_h) {
      c.th = 0;
      if (++c.b == B) {
        c.b = 0;
        ++c.chunk;
      }
    }
  }
}

// ------------------------------------------------------------------ forward (and data gradient with flipped taps)
// One persistent CTA per SM, 16 warps.  Work items = (32-channel chunk, image, 16-row x WT-column output tile, WT <= 32), chunk-major, a
// contiguous range per CTA.  Per item: the halo tile [22][RP] pixels arrives by one 4-D TMA load (hardware zero fill = the conv padding)
// issued one item ahead into the other buffer; every warp computes one (column block, channel group) unit and writes / adds its bf16
// outputs into a swizzled staging tile [16][WT] x 64 B; the staging tile leaves by ONE 4-D TMA store (whole 64-byte pixel rows, clipped
// at the image edge by the hardware) issued by thread 0 once every warp has arrived on the item's "done" barrier.  For the data gradient
// the skip-connection gradient is TMA-loaded INTO the staging tile first and the epilogue adds to it in place.  (Per-thread 16-byte
// stores straight to global cost 40 % of the kernel: half-written 32-byte sectors.)  The band table [32][7][11] x 8 B + bias of the
// chunk is rebuilt only when the chunk changes (at most twice per CTA).
constexpr int FWD_WARPS = 16;

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__global__ void __launch_bounds__(FWD_WARPS * 32, 1)
    dwconv7_fwd_mma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                           const float* __restrict__ w, const float* __restrict__ bias, int has_res, int B, int W, int C, int tiles_w, int tiles_h, int WT,
                           int RP, int nxb, int wl, int tile_bytes, int out_bytes, int dbg) {
  constexpr int NW = FWD_WARPS;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* tiles = smem;                               // [2][tile_bytes]
  unsigned char* outs = smem + 2 * (size_t)tile_bytes;       // [2][out_bytes]
  unsigned char* rest = outs + 2 * (size_t)out_bytes;
  uint2* btab = reinterpret_cast<uint2*>(rest);
  float* bias_s = reinterpret_cast<float*>(rest + BTAB_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(rest + BTAB_BYTES + CC * 4);
  uint64_t* empty = full + 2;    // = "done": every warp has read the halo tile and written its outputs
  uint64_t* outfree = full + 4;  // the staging tile may be written (its previous TMA store has been read out; the residual has landed)
  const uint32_t tiles_s = smem_u32(tiles), outs_s = smem_u32(outs);
  const uint32_t load_bytes = (uint32_t)(IN_ROWS * RP * 64), res_bytes = (uint32_t)(TROWS * WT * 64);

  const long long n_items = (long long)B * tiles_w * tiles_h * (C / CC);
  const int it_begin = (int)(n_items * blockIdx.x / gridDim.x), it_end = (int)(n_items * (blockIdx.x + 1) / gridDim.x);
  const int n_my = it_end - it_begin;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmY);
    if (has_res) prefetch_tmap(&tmR);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NW);
      mbar_init(&outfree[i], 1);
    }
    mbar_fence_init();
  }
  // the slack behind each halo buffer is never written by TMA: it only has to hold finite values (it meets zero band entries)
  for (int i = threadIdx.x; i < 2 * (SLACK_BYTES / 16); i += NW * 32)
    reinterpret_cast<uint4*>(tiles + (size_t)(i / (SLACK_BYTES / 16)) * tile_bytes + load_bytes)[i % (SLACK_BYTES / 16)] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  Item cur = item_of(it_begin, tiles_w, tiles_h, B);
  Item prev = cur;
  auto claim_out = [&](int n, const Item& c) {  // staging tile n & 1 may be written for the CTA's n-th item (after its skip gradient has landed)
    if (has_res) {
      mbar_expect_tx(&outfree[n & 1], res_bytes);
      tma_load_4d(outs + (size_t)(n & 1) * out_bytes, &tmR, &outfree[n & 1], c.chunk * CC, c.tw * WT, c.th * TROWS, c.b);
    } else {
      mbar_arrive(&outfree[n & 1]);
    }
  };
  if (threadIdx.x == 0 && n_my > 0) {
    mbar_expect_tx(&full[0], load_bytes);
    tma_load_4d(tiles, &tmX, &full[0], cur.chunk * CC, cur.tw * WT - 3, cur.th * TROWS - 3, cur.b);
    claim_out(0, cur);
    if (n_my > 1) {
      Item c1 = cur;
      item_next(c1, tiles_w, tiles_h, B);
      claim_out(1, c1);
    }
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int e = t - g + 7;
  const int hRP = RP >> 1;
  const int P = 2 * g * RP + t;
  int cur_chunk = -1;
  int duty = 0;  // the warp whose lane 0 does the per-item TMA chores: it rotates, so no warp lags behind the others item after item
  for (int it = 0; it < n_my; ++it) {
    Item nxt = cur;
    item_next(nxt, tiles_w, tiles_h, B);
    const int buf = it & 1;
    if (warp == duty && lane == 0) {
      if (it >= 1) {  // item it-1 is complete in its staging tile: send it off
        mbar_wait(&empty[buf ^ 1], (((uint32_t)(it - 1)) >> 1) & 1u);
        if (!(dbg & 2)) tma_store_4d(&tmY, outs + (size_t)(buf ^ 1) * out_bytes, prev.chunk * CC, prev.tw * WT, prev.th * TROWS, prev.b);
        bulk_commit_group();
      }
      if (it + 1 < n_my) {  // halo tile of the next item into the buffer item it-1 has just released
        mbar_expect_tx(&full[buf ^ 1], load_bytes);
        tma_load_4d(tiles + (size_t)(buf ^ 1) * tile_bytes, &tmX, &full[buf ^ 1], nxt.chunk * CC, nxt.tw * WT - 3, nxt.th * TROWS - 3, nxt.b);
      }
      if (it >= 1) {
        bulk_wait_group_read<0>();  // the store just issued has been read out of its staging tile: hand the tile to item it+1
        if (it + 1 < n_my) claim_out(it + 1, nxt);
      }
    }
    duty = duty + 1 == NW ? 0 : duty + 1;
    const int c0 = cur.chunk * CC;
    if (cur.chunk != cur_chunk) {  // uniform over the CTA
      cur_chunk = cur.chunk;
      __syncthreads();  // every warp is done with the previous chunk's table
      // band table: entry (c, ky, d + 7), d = t - g in -7..3: b0 = {w[d], w[d+4]}, b1 = {w[d+8], w[d+12]} (bf16, zero outside 0..6)
      for (int i = threadIdx.x; i < BTAB_ENTRIES; i += NW * 32) {
        const int ee = i % 11, ky = (i / 11) % 7, c = i / 77;
        const int d = ee - 7;
        uint32_t v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int kx = d + 4 * q;
          v[q] = (kx >= 0 && kx <= 6) ? (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(w[widx(wl, ky * 7 + kx, c0 + c, C)])) : 0u;
        }
        btab[i] = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
      }
      if (threadIdx.x < CC) bias_s[threadIdx.x] = bias ? bias[c0 + threadIdx.x] : 0.f;
      __syncthreads();
    }
    prev = cur;
    cur = nxt;
    const uint32_t par = ((uint32_t)it >> 1) & 1u;
    mbar_wait(&full[buf], par);
    const uint32_t tile_s = tiles_s + (uint32_t)buf * (uint32_t)tile_bytes;
    const uint32_t out_s = outs_s + (uint32_t)buf * (uint32_t)out_bytes;

    // column blocks of this item: the last column tile of an image may be narrower than WT (W = 56: tiles of 32 + 24 columns)
    const int nxb_it = min(nxb, (W - prev.tw * WT + 7) >> 3);
    for (int u = warp; u < nxb_it * 4; u += NW) {
      const int j = u % nxb_it, cg = u / nxb_it;
      float acc[8][4];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const float bv = bias_s[cg * 8 + ch];
        acc[ch][0] = acc[ch][1] = acc[ch][2] = acc[ch][3] = bv;
      }
      const uint2* bt = btab + (cg * 8) * 77 + e;
      const int p = P + 8 * j;
      uint32_t lin = tile_s + (uint32_t)p * 64u;
      int s = p >> 1;
      uint32_t plo[8], phi[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const Quad q = lds_quad(lin + (uint32_t)(((cg ^ s) & 3) << 4));
        lin += (uint32_t)RP * 64u;
        s += hRP;
        uint32_t clo[8], chi[8];
        planar8(q.q0, q.q1, clo);
        planar8(q.q2, q.q3, chi);
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
      // thread (g, t): rows 2g, 2g+1 x columns 8j+2t, 8j+2t+1 x 8 channels = four 16-byte chunks of the staging tile
      mbar_wait(&outfree[buf], par);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
        for (int xx = 0; xx < 2; ++xx) {
          const int xl = 8 * j + 2 * t + xx;
          if (xl < WT) {
            const uint32_t addr = out_s + swz((2 * g + hh) * WT + xl, cg);
            float v[8];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) v[ch] = acc[ch][2 * hh + xx];
            if (has_res) {  // fused "+ residual" (the skip-connection gradient when this kernel runs as the data gradient)
              const uint4 rv = lds128<0>(addr);
              const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[2 * q] += __uint_as_float(rw[q] << 16);
                v[2 * q + 1] += __uint_as_float(rw[q] & 0xffff0000u);
              }
            }
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pack_bf16x2(v[0], v[1])), "r"(pack_bf16x2(v[2], v[3])),
                         "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7]))
                         : "memory");
          }
        }
      }
    }
    fence_proxy_async_smem();  // this thread's staging writes become visible to the TMA store
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[buf]);
  }
  if (threadIdx.x == 0 && n_my > 0) {
    const int lb = (n_my - 1) & 1;
    mbar_wait(&empty[lb], (((uint32_t)(n_my - 1)) >> 1) & 1u);
    if (!(dbg & 2)) tma_store_4d(&tmY, outs + (size_t)lb * out_bytes, prev.chunk * CC, prev.tw * WT, prev.th * TROWS, prev.b);
    bulk_commit_group();
    bulk_wait_group_all();
  }
}

// ------------------------------------------------------------------ weight gradient
// One persistent CTA per SM, 16 warps: warp & 3 = channel group of 8, warp >> 2 = row pair of the band.  Work items = (chunk, image,
// band of 8 padded input rows), chunk-major; the input band [8][RP] and the dy band [14][RPG] of the next item arrive by two TMA loads
// while this one is computed.  The accumulators live in registers until the chunk changes or the CTA runs out of items.
__device__ __forceinline__ void wgrad_flush(float (&acc)[8][4], float* red, float* __restrict__ dw, float* __restrict__ dbias, int c0, int C, int wl,
                                            int g, int t, int cg) {
  constexpr int NT = WG_WARPS * 32;
  for (int i = threadIdx.x; i < 50 * CC; i += NT) red[i] = 0.f;
  __syncthreads();
  // c0, c1 = D[g][2t], D[g][2t+1];  c2, c3 = D[g+8][2t], D[g+8][2t+1]
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
      } else if (n == 2 || n == 3) {  // row of ones x dy rows yi - 3, yi - 2 = the two IMAGE rows of this row pair: every dy row exactly once
        atomicAdd(&red[49 * CC + c], acc[ch][q]);
      }
    }
    acc[ch][0] = acc[ch][1] = acc[ch][2] = acc[ch][3] = 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 49 * CC; i += NT) atomicAdd(dw + widx(wl, i / CC, c0 + (i % CC), C), red[i]);
  if (dbias)
    for (int i = threadIdx.x; i < CC; i += NT) atomicAdd(dbias + c0 + i, red[49 * CC + i]);
  __syncthreads();
}

__global__ void __launch_bounds__(WG_WARPS * 32, 1)
    dwconv7_wgrad_mma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, float* __restrict__ dw,
                             float* __restrict__ dbias, int B, int H, int W, int C, int tiles_h, int RP, int RPG, int nxc, int wl, int in_bytes,
                             int g_bytes, int band_rows) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int stage_bytes = in_bytes + g_bytes;
  float* red = reinterpret_cast<float*>(smem + 2 * (size_t)stage_bytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * (size_t)stage_bytes + WG_RED_BYTES);
  uint64_t* empty = full + 2;
  const uint32_t smem_s = smem_u32(smem);
  const uint32_t in_load = (uint32_t)(band_rows * RP * 64), g_load = (uint32_t)((band_rows + 6) * RPG * 64);
  constexpr int NT = WG_WARPS * 32;

  const long long n_items = (long long)B * tiles_h * (C / CC);
  const int it_begin = (int)(n_items * blockIdx.x / gridDim.x), it_end = (int)(n_items * (blockIdx.x + 1) / gridDim.x);
  const int n_my = it_end - it_begin;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmG);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], WG_WARPS);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 2 * (SLACK_BYTES / 16); i += NT)
    reinterpret_cast<uint4*>(smem + (size_t)(i / (SLACK_BYTES / 16)) * stage_bytes + in_load)[i % (SLACK_BYTES / 16)] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  auto issue = [&](int n, const Item& c) {  // c.th = band index, c.tw unused
    const int y0 = c.th * band_rows;  // first IMAGE row of the input band (= padded row y0 + 3: the all-zero padding rows above / below the
                                      // image are never visited); the dy band starts six padded rows = three image rows higher
    const int buf = n & 1;
    unsigned char* st = smem + (size_t)buf * stage_bytes;
    mbar_wait(&empty[buf], (((uint32_t)n >> 1) & 1u) ^ 1u);
    mbar_expect_tx(&full[buf], in_load + g_load);
    tma_load_4d(st, &tmX, &full[buf], c.chunk * CC, -3, y0, c.b);
    tma_load_4d(st + in_bytes, &tmG, &full[buf], c.chunk * CC, 0, y0 - 3, c.b);
  };
  Item cur = item_of(it_begin, 1, tiles_h, B);
  if (threadIdx.x == 0 && n_my > 0) issue(0, cur);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int cg = warp & 3, rp = warp >> 2;
  // per-lane byte offsets inside a stage of the warp's first row pair (the swizzle term does not change when the column advances by 16
  // pixels; the second row pair of a 16-row band lies 8 tile rows further down)
  const int pa0 = (2 * rp) * RP + t + g, pa1 = pa0 + RP;
  const int pb = ((g < 7) ? (2 * rp + 6 - g) : (2 * rp + 7)) * RPG + t;
  const int nrp = band_rows / 8;

  float acc[8][4];
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) acc[ch][0] = acc[ch][1] = acc[ch][2] = acc[ch][3] = 0.f;

  int cur_chunk = -1;
  for (int it = 0; it < n_my; ++it) {
    Item nxt = cur;
    item_next(nxt, 1, tiles_h, B);
    if (threadIdx.x == 0 && it + 1 < n_my) issue(it + 1, nxt);
    if (cur.chunk != cur_chunk) {  // uniform over the CTA
      if (cur_chunk >= 0) wgrad_flush(acc, red, dw, dbias, cur_chunk * CC, C, wl, g, t, cg);
      cur_chunk = cur.chunk;
    }
    cur = nxt;
    const int buf = it & 1;
    mbar_wait(&full[buf], ((uint32_t)it >> 1) & 1u);
    const uint32_t st = smem_s + (uint32_t)buf * (uint32_t)stage_bytes;
#pragma unroll 1
    for (int k = 0; k < nrp; ++k) {
      const uint32_t oa0 = st + swz(pa0 + 8 * k * RP, cg), oa1 = st + swz(pa1 + 8 * k * RP, cg);
      const uint32_t ob = st + (uint32_t)in_bytes + swz(pb + 8 * k * RPG, cg);
#pragma unroll 1
      for (int xc = 0; xc < nxc; ++xc) {
        const uint32_t xo = (uint32_t)xc * 1024u;
        uint32_t a0[8], a1[8], a2[8], a3[8], b0[8], b1[8];
        {
          const Quad q = lds_quad(oa0 + xo);
          planar8(q.q0, q.q1, a0);
          planar8(q.q2, q.q3, a2);
        }
        {
          const Quad q = lds_quad(oa1 + xo);
          planar8(q.q0, q.q1, a1);
          planar8(q.q2, q.q3, a3);
        }
        {
          const Quad q = lds_quad(ob + xo);
          planar8(q.q0, q.q1, b0);
          planar8(q.q2, q.q3, b1);
        }
        if (g == 7) {  // row m = 7 of A: ones -> column sums of dy (bias gradient)
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) a0[ch] = a2[ch] = 0x3F803F80u;
        }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) mma16816(acc[ch], a0[ch], a1[ch], a2[ch], a3[ch], b0[ch], b1[ch]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[buf]);
  }
  if (cur_chunk >= 0) wgrad_flush(acc, red, dw, dbias, cur_chunk * CC, C, wl, g, t, cg);
}

int round_up_mod(int v, int m, int r) {  // smallest value >= v that is r mod m
  int o = v - ((v - r) % m + m) % m;
  return o < v ? o + m : o;
}

// [B][H][W][C] bf16, box = (32 channels, bw, bh, 1 image), 64-byte swizzle, zero fill outside
bool make_nhwc_sw64_tmap(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int bw, int bh) {
  auto enc = get_encode_fn();
  if (!enc) return false;
  // L2 promotion of the 64-byte pixel rows (a 32-channel chunk of a 192- / 384-byte pixel): 64 B measured best (weight gradient 0.162 ->
  // 0.147 ms, data gradient 0.175 -> 0.165 ms at 256x56x56x96; none / 128 B / 256 B are equal).  LNX_DW_PROMO = 0 / 1 / 2 / 3 overrides.
  static const int promo = getenv("LNX_DW_PROMO") ? atoi(getenv("LNX_DW_PROMO")) : 1;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)CC, (cuuint32_t)bw, (cuuint32_t)bh, 1u};
  cuuint32_t es[4] = {1u, 1u, 1u, 1u};
  const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                    : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                    : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                 : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr size_t kMaxSmem = 227 * 1024;

}  // namespace

static int g_dwconv_impl = -1;  // -1: LNX_DWCONV_MMA or the default (tensor pipe)

bool lnx_dwconv7_mma_enabled() {
  static const int env = getenv("LNX_DWCONV_MMA") ? atoi(getenv("LNX_DWCONV_MMA")) : 1;
  return (g_dwconv_impl < 0 ? env : g_dwconv_impl) != 0;
}

extern "C" int lnx_dwconv7_set_impl(int impl) {
  const int prev = g_dwconv_impl;
  g_dwconv_impl = impl < 0 ? -1 : (impl ? 1 : 0);
  return prev;
}

// -> LNX_OK, or LNX_ERR_UNSUPPORTED when the caller should use the FFMA2 kernels
int lnx_dwconv7_fwd_mma(const void* x, const float* w, int wl, const float* bias, const void* res, void* y, int B, int H, int W, int C,
                        cudaStream_t st) {
  if (C % CC != 0) return LNX_ERR_SHAPE;
  const int tiles_w = (W + 31) / 32;
  const int WT = tiles_w == 1 ? W : 32;  // full 32-column tiles + one narrower last tile: no column block is computed for nothing
  const int RP = round_up_mod(WT + 6, 4, 2);  // 2 mod 4: two tile rows apart = 4 pixels mod 8 (bank-conflict-free quads)
  const int nxb = (WT + 7) / 8;
  const int tiles_h = (H + TROWS - 1) / TROWS;
  const int tile_bytes = (IN_ROWS * RP * 64 + SLACK_BYTES + 1023) / 1024 * 1024;
  const int out_bytes = (TROWS * WT * 64 + 1023) / 1024 * 1024;
  const size_t smem = 2 * (size_t)tile_bytes + 2 * (size_t)out_bytes + BTAB_BYTES + CC * 4 + 64;
  if (smem > kMaxSmem) return LNX_ERR_UNSUPPORTED;
  CUtensorMap tmX, tmY, tmR;
  if (!make_nhwc_sw64_tmap(&tmX, x, B, H, W, C, RP, IN_ROWS) || !make_nhwc_sw64_tmap(&tmY, y, B, H, W, C, WT, TROWS)) return LNX_ERR_UNSUPPORTED;
  if (!make_nhwc_sw64_tmap(&tmR, res ? res : y, B, H, W, C, WT, TROWS)) return LNX_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dwconv7_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    attr_set = true;
  }
  static const int dbg = getenv("LNX_DW_DBG") ? atoi(getenv("LNX_DW_DBG")) : 0;  // profiling ablation: 2 = no output stores
  const long long n_items = (long long)B * tiles_h * tiles_w * (C / CC);
  const int grid = (int)(n_items < kNumSMs ? n_items : kNumSMs);
  dwconv7_fwd_mma_kernel<<<grid, FWD_WARPS * 32, smem, st>>>(tmX, tmY, tmR, w, bias, res ? 1 : 0, B, W, C, tiles_w, tiles_h, WT, RP, nxb, wl, tile_bytes,
                                                             out_bytes, dbg);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

int lnx_dwconv7_wgrad_mma(const void* x, const void* dy, float* dw, int wl, float* dbias, int B, int H, int W, int C, cudaStream_t st) {
  if (C % CC != 0) return LNX_ERR_SHAPE;
  if (W > 64) return LNX_ERR_UNSUPPORTED;
  const int nxc = (W + 15) / 16;
  const int RP = W + 6;
  const int RPG = round_up_mod(16 * nxc, 8, 4);  // 4 mod 8: neighbouring dy rows = the other four pixels mod 8
  // bands of 16 input rows when two stages of them fit in shared memory and they waste no more rows than bands of 8
  int band_rows = ((H + 15) / 16 * 16 <= (H + 7) / 8 * 8) ? 16 : 8, in_bytes = 0, g_bytes = 0;
  size_t smem = 0;
  for (;; band_rows = 8) {
    in_bytes = (band_rows * RP * 64 + SLACK_BYTES + 511) / 512 * 512;
    g_bytes = ((band_rows + 6) * RPG * 64 + 511) / 512 * 512;
    smem = 2 * (size_t)(in_bytes + g_bytes) + WG_RED_BYTES + 64;
    if (smem <= kMaxSmem || band_rows == 8) break;
  }
  if (smem > kMaxSmem) return LNX_ERR_UNSUPPORTED;
  const int tiles_h = (H + band_rows - 1) / band_rows;  // bands cover the image rows only
  CUtensorMap tmX, tmG;
  if (!make_nhwc_sw64_tmap(&tmX, x, B, H, W, C, RP, band_rows) || !make_nhwc_sw64_tmap(&tmG, dy, B, H, W, C, RPG, band_rows + 6)) return LNX_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dwconv7_wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    attr_set = true;
  }
  const long long n_items = (long long)B * tiles_h * (C / CC);
  const int grid = (int)(n_items < kNumSMs ? n_items : kNumSMs);
  dwconv7_wgrad_mma_kernel<<<grid, WG_WARPS * 32, smem, st>>>(tmX, tmG, dw, dbias, B, H, W, C, tiles_h, RP, RPG, nxc, wl, in_bytes, g_bytes, band_rows);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
