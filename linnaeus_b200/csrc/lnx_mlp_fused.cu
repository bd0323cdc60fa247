// Fused ConvNeXt pointwise pair (convnext.py:79-86) on tcgen05: the 4C-wide hidden tensor never leaves the SM.
//
//   y = residual + row_scale * gamma * (gelu(x W1^T + b1) W2^T + b2)          x, y, residual: [M, C] bf16
//
// One persistent CTA per SM walks 128-row tiles.  Per tile and per chunk of HC hidden units:
//   MMA1  pre[128, HC]  = x_tile W1[chunk]^T            A, B from shared memory, accumulator in TMEM
//   GELU  h = gelu(pre + b1) -> bf16, written back INTO the accumulator's own TMEM columns (tcgen05.st)
//   MMA2  y[128, C]    += h W2[:, chunk]^T              A from TMEM (the bf16 hidden tile), B from shared memory
// and once per tile the output warps apply b2, gamma, DropPath scale and the residual and store with TMA.
// C = 96 keeps W1 and W2 resident in shared memory for the whole kernel (147 KB), so the only DRAM / L2 traffic
// is x, the residual and y; wider blocks stream the weight chunks through two TMA rings (L2 resident).
//
// Warp roles: 0 TMA producer (x tiles, W1), 1 MMA issuer, 2-5 output warps (one per TMEM lane quarter),
// 6..6+NGW-1 GELU warps (NGW/4 per lane quarter, each a private column slice), last warp W2 producer (streamed).
// The MMA warp runs MMA1 one chunk ahead of MMA2, so the tensor pipe works on chunk g+1 while the GELU warps
// are busy with chunk g; tcgen05.mma executes in issue order, which is what makes the in-place hidden tile safe.
#include <stdlib.h>

#include "lnx_mlp_fused.cuh"

using namespace lnx;
using namespace lnx_tc;
using namespace lnx_mlp;

namespace {

// Ablation switches for profiling builds (tools/build_variant.sh -DLNX_DBG=n; the shipped library is built with 0, results are wrong
// otherwise): 1 bias from a constant, 4 no GELU math (raw accumulator bits packed), 16 no MMA1, 32 no MMA2.
#ifndef LNX_DBG
#define LNX_DBG 0
#endif

struct FusedArgs {
  const float* b1;
  const float* b2;
  const float* gamma;      // nullable (no layer scale)
  const float* row_scale;  // nullable (DropPath inactive)
  int rows_per_group;
  int M;
  int has_res;
  int num_tiles;
};

template <int C_, int HC_, bool RESIDENT_, int NGW_>
struct Cfg {
  static constexpr int C = C_, HC = HC_, NGW = NGW_;
  static constexpr bool RESIDENT = RESIDENT_;
  static constexpr int H = 4 * C;
  static constexpr int NC = H / HC;             // hidden chunks per tile
  static constexpr int KB64 = C / 64;           // 64-wide (128B swizzle) k-blocks of x / W1
  static constexpr int KREM = C % 64;           // 0 or 32: one more 32-wide (64B swizzle) k-block
  static constexpr int NPRE = (HC == 128) ? 2 : 4;
  static constexpr int NY = (NPRE * HC + 2 * C <= 512) ? 2 : 1;
  static constexpr int NX = RESIDENT ? 2 : 1;
  static constexpr int S1 = 3, S2 = 2;          // streamed weight rings
  static constexpr int X_BYTES = BM * C * 2;
  static constexpr int STG_BYTES = BM * C * 2;  // 4 output warps x (C / 32) blocks of [32 rows][32 cols]
  static constexpr int W1C_BYTES = HC * C * 2;  // one hidden chunk of W1: KB64 blocks [HC][64] (+ [HC][32])
  static constexpr int W2C_BYTES = C * HC * 2;  // one hidden chunk of W2: HC / 64 blocks [C][64]
  static constexpr int W1_REGION = RESIDENT ? H * C * 2 : S1 * W1C_BYTES;
  static constexpr int W2_REGION = RESIDENT ? C * H * 2 : S2 * W2C_BYTES;
  static constexpr int W1_KB_STRIDE = (RESIDENT ? H : HC) * 128;  // bytes between 64-wide k-blocks of W1
  static constexpr int VEC_BYTES = (H + 2 * C) * 4;              // b1 | gamma | b2 * gamma
  static constexpr int NBAR = 2 + 2 + 1 + 4 * 4 + 4 + 4 + 2 + 2 + 4;
  static constexpr int SMEM = W1_REGION + W2_REGION + NX * X_BYTES + STG_BYTES + VEC_BYTES + NBAR * 8 + 16 + 1024;
  static constexpr int NWARPS = 6 + NGW + (RESIDENT ? 0 : 1);
  static constexpr int NTHREADS = 32 * NWARPS;
  static constexpr int CPS = HC / (NGW / 4);    // hidden columns per GELU warp and chunk
  static_assert(C % 32 == 0 && (KREM == 0 || KREM == 32), "C must be a multiple of 32");
  static_assert(H % HC == 0 && HC % 64 == 0 && CPS % 32 == 0, "hidden chunking");
  static_assert(NPRE * HC + NY * C <= 512, "TMEM budget");
  static_assert(SMEM <= MAX_SMEM, "shared memory budget");
  static_assert(C <= 256, "single MMA2 accumulator");
};

template <class CF>
__global__ void __launch_bounds__(CF::NTHREADS, 1)
    mlp_fused_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXr, const __grid_constant__ CUtensorMap tmW1,
                         const __grid_constant__ CUtensorMap tmW1r, const __grid_constant__ CUtensorMap tmW2,
                         const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmY, const FusedArgs a) {
  constexpr int C = CF::C, HC = CF::HC, H = CF::H, NC = CF::NC, KB64 = CF::KB64, KREM = CF::KREM, NPRE = CF::NPRE, NY = CF::NY, NX = CF::NX;
  constexpr bool RESIDENT = CF::RESIDENT;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* w1s = base;
  unsigned char* w2s = w1s + CF::W1_REGION;
  unsigned char* xs = w2s + CF::W2_REGION;
  unsigned char* stg = xs + NX * CF::X_BYTES;
  float* b1s = reinterpret_cast<float*>(stg + CF::STG_BYTES);
  float* g2s = b1s + H;   // 0.5 * gamma (gamma = 1 when absent)
  float* bbs = g2s + C;   // b2 * gamma
  uint64_t* bars = reinterpret_cast<uint64_t*>(bbs + C);
  uint64_t* x_full = bars;           // [2]
  uint64_t* x_empty = x_full + 2;    // [2]
  uint64_t* w_full = x_empty + 2;    // [1]   resident weights landed
  uint64_t* w1_full = w_full + 1;    // [4]
  uint64_t* w1_empty = w1_full + 4;  // [4]
  uint64_t* w2_full = w1_empty + 4;  // [4]
  uint64_t* w2_empty = w2_full + 4;  // [4]
  uint64_t* pre_full = w2_empty + 4; // [4]   MMA1 of a chunk complete
  uint64_t* h_full = pre_full + 4;   // [4]   hidden tile written back to TMEM by every GELU warp
  uint64_t* y_full = h_full + 4;     // [2]
  uint64_t* y_empty = y_full + 2;    // [2]
  uint64_t* res_full = y_empty + 2;  // [4]   per output warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = ((int)blockIdx.x < a.num_tiles) ? (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t total_chunks = (uint32_t)my_tiles * NC;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2); prefetch_tmap(&tmY);
    if (KREM) { prefetch_tmap(&tmXr); prefetch_tmap(&tmW1r); }
    if (a.has_res) prefetch_tmap(&tmRes);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], 4);
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&w1_full[i], 1);
      mbar_init(&w1_empty[i], 1);
      mbar_init(&w2_full[i], 1);
      mbar_init(&w2_empty[i], 1);
      mbar_init(&pre_full[i], 1);
      mbar_init(&h_full[i], CF::NGW);
      mbar_init(&res_full[i], 1);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < H; i += CF::NTHREADS) b1s[i] = a.b1 ? a.b1[i] : 0.f;
  for (int i = threadIdx.x; i < C; i += CF::NTHREADS) {
    const float gm = a.gamma ? a.gamma[i] : 1.f;
    g2s[i] = 0.5f * gm;  // the hidden tile holds 2 gelu(pre)
    bbs[i] = (a.b2 ? a.b2[i] : 0.f) * gm;
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t Y_COL0 = NPRE * HC;

  if (warp == 0) {
    // ===================== producer: resident weights once, then x tiles (+ the W1 ring when streamed) =====================
    if (lane == 0) {
      if (RESIDENT) {
        mbar_expect_tx(w_full, (uint32_t)(CF::W1_REGION + CF::W2_REGION));
        for (int j = 0; j < NC; ++j) {
          for (int kb = 0; kb < KB64; ++kb) tma_load_2d(w1s + kb * CF::W1_KB_STRIDE + j * HC * 128, &tmW1, w_full, kb * 64, j * HC);
          if (KREM) tma_load_2d(w1s + KB64 * CF::W1_KB_STRIDE + j * HC * 64, &tmW1r, w_full, KB64 * 64, j * HC);
        }
        for (int kb = 0; kb < H / 64; ++kb) tma_load_2d(w2s + kb * (C * 128), &tmW2, w_full, kb * 64, 0);
      }
      uint32_t g = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int t = (int)blockIdx.x + tl * (int)gridDim.x;
        const int m0 = t * BM;
        const int s = tl % NX;
        mbar_wait_relaxed(&x_empty[s], (((uint32_t)tl / NX) & 1u) ^ 1u);
        mbar_expect_tx(&x_full[s], CF::X_BYTES);
        unsigned char* xb = xs + s * CF::X_BYTES;
        for (int kb = 0; kb < KB64; ++kb) tma_load_2d(xb + kb * (BM * 128), &tmX, &x_full[s], kb * 64, m0);
        if (KREM) tma_load_2d(xb + KB64 * (BM * 128), &tmXr, &x_full[s], KB64 * 64, m0);
        if (!RESIDENT) {
          for (int j = 0; j < NC; ++j, ++g) {
            const int st = g % CF::S1;
            mbar_wait_relaxed(&w1_empty[st], ((g / CF::S1) & 1u) ^ 1u);
            mbar_expect_tx(&w1_full[st], CF::W1C_BYTES);
            unsigned char* wb = w1s + st * CF::W1C_BYTES;
            for (int kb = 0; kb < KB64; ++kb) tma_load_2d(wb + kb * CF::W1_KB_STRIDE, &tmW1, &w1_full[st], kb * 64, j * HC);
            if (KREM) tma_load_2d(wb + KB64 * CF::W1_KB_STRIDE, &tmW1r, &w1_full[st], KB64 * 64, j * HC);
          }
        }
      }
    }
  } else if (!RESIDENT && warp == CF::NWARPS - 1) {
    // ===================== W2 chunk producer (streamed) =====================
    if (lane == 0) {
      for (uint32_t g = 0; g < total_chunks; ++g) {
        const int j = g % NC;
        const int st = g % CF::S2;
        mbar_wait_relaxed(&w2_empty[st], ((g / CF::S2) & 1u) ^ 1u);
        mbar_expect_tx(&w2_full[st], CF::W2C_BYTES);
        unsigned char* wb = w2s + st * CF::W2C_BYTES;
        for (int kb = 0; kb < HC / 64; ++kb) tma_load_2d(wb + kb * (C * 128), &tmW2, &w2_full[st], j * HC + kb * 64, 0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_bf16(BM, HC, 0, 0);
      const uint32_t idesc2 = make_idesc_bf16(BM, C, 0, 0);
      if (RESIDENT && total_chunks > 0) {
        mbar_wait_relaxed(w_full, 0);
        tcgen05_fence_after();
      }
      for (uint32_t g = 0; g <= total_chunks; ++g) {
        if (g < total_chunks) {
          // ---- MMA1 of chunk g: pre = x_tile W1[chunk]^T
          const uint32_t tl = g / NC, j = g % NC;
          const uint32_t s = tl % NX;
          if (j == 0) {
            mbar_wait_relaxed(&x_full[s], (tl / NX) & 1u);
            tcgen05_fence_after();
          }
          uint32_t w1c;
          if (RESIDENT) {
            w1c = smem_u32(w1s) + j * HC * 128;
          } else {
            const uint32_t st = g % CF::S1;
            mbar_wait_relaxed(&w1_full[st], (g / CF::S1) & 1u);
            tcgen05_fence_after();
            w1c = smem_u32(w1s) + st * CF::W1C_BYTES;
          }
          const uint32_t xa = smem_u32(xs) + s * CF::X_BYTES;
          const uint32_t tpre = tmem_base + (g % NPRE) * HC;
          if (!(LNX_DBG & 16))
#pragma unroll
          for (int kb = 0; kb < KB64; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tpre, desc_k128(xa + kb * (BM * 128) + k * 32), desc_k128(w1c + kb * CF::W1_KB_STRIDE + k * 32), idesc1,
                        (kb > 0 || k > 0) ? 1u : 0u);
          if (KREM && !(LNX_DBG & 16)) {
            // the 32-wide remainder block: [rows][64 B], 64-byte swizzle; chunk j starts j * HC * 64 bytes into the resident block
            const uint32_t w1r = RESIDENT ? smem_u32(w1s) + KB64 * CF::W1_KB_STRIDE + j * HC * 64 : w1c + KB64 * CF::W1_KB_STRIDE;
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16(tpre, desc_k64(xa + KB64 * (BM * 128) + k * 32), desc_k64(w1r + k * 32), idesc1, (KB64 > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&pre_full[g % NPRE]);
          if (!RESIDENT) umma_commit(&w1_empty[g % CF::S1]);
          if (j == NC - 1) umma_commit(&x_empty[s]);
        }
        if (g >= 1) {
          // ---- MMA2 of chunk g - 1: y += h W2[:, chunk]^T, h read from TMEM
          const uint32_t gp = g - 1;
          const uint32_t tl = gp / NC, j = gp % NC;
          const uint32_t ys = tl % NY;
          if (j == 0) {
            mbar_wait_relaxed(&y_empty[ys], ((tl / NY) & 1u) ^ 1u);
            tcgen05_fence_after();
          }
          mbar_wait(&h_full[gp % NPRE], (gp / NPRE) & 1u);  // on the critical path: no sleep between polls
          tcgen05_fence_after();
          uint32_t w2c;
          if (RESIDENT) {
            w2c = smem_u32(w2s) + j * (HC / 64) * (C * 128);
          } else {
            const uint32_t st = gp % CF::S2;
            mbar_wait_relaxed(&w2_full[st], (gp / CF::S2) & 1u);
            tcgen05_fence_after();
            w2c = smem_u32(w2s) + st * CF::W2C_BYTES;
          }
          const uint32_t th = tmem_base + (gp % NPRE) * HC;
          const uint32_t ty = tmem_base + Y_COL0 + ys * C;
          if (!(LNX_DBG & 32))
#pragma unroll
          for (int kk = 0; kk < HC / 16; ++kk)
            umma_bf16_ts(ty, th + 32 * (kk >> 1) + 8 * (kk & 1), desc_k128(w2c + (kk >> 2) * (C * 128) + (kk & 3) * 32), idesc2,
                         (j > 0 || kk > 0) ? 1u : 0u);
          if (!RESIDENT) umma_commit(&w2_empty[gp % CF::S2]);
          if (j == NC - 1) umma_commit(&y_full[ys]);
        }
      }
    }
  } else if (warp < 6) {
    // ===================== output warps: y accumulator -> b2, gamma, DropPath scale, + residual -> TMA store =====================
    const int q = warp & 3;                // TMEM lane quarter
    constexpr int NB = C / 32;             // [32 rows][32 cols] blocks, 64-byte swizzle, 2 KB each
    unsigned char* st = stg + q * (NB * 2048);
    uint64_t* my_res = &res_full[q];
    const uint32_t sw = (uint32_t)((lane >> 1) & 3);
    auto issue_res = [&](int tl) {
      const int t = (int)blockIdx.x + tl * (int)gridDim.x;
      mbar_expect_tx(my_res, NB * 2048);
      for (int cb = 0; cb < NB; ++cb) tma_load_2d(st + cb * 2048, &tmRes, my_res, cb * 32, t * BM + q * 32);
    };
    if (a.has_res && lane == 0 && my_tiles > 0) issue_res(0);
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int t = (int)blockIdx.x + tl * (int)gridDim.x;
      const int row0 = t * BM + q * 32;
      const int m = row0 + lane;
      const uint32_t ys = tl % NY;
      float rs = 1.f;
      if (a.row_scale && m < a.M) rs = a.row_scale[m / a.rows_per_group];
      mbar_wait(&y_full[ys], ((uint32_t)tl / NY) & 1u);
      tcgen05_fence_after();
      if (a.has_res) mbar_wait(my_res, (uint32_t)tl & 1u);
      const uint32_t trow = tmem_base + Y_COL0 + ys * C + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int cb = 0; cb < NB; ++cb) {
        uint32_t acc[32];
        tmem_ld32_nowait(trow + cb * 32, acc);
        tmem_ld_wait();
        unsigned char* blk = st + cb * 2048;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint32_t soff = (uint32_t)(lane * 64) + ((((uint32_t)jj) ^ sw) << 4);
          float vv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = cb * 32 + jj * 8 + k;
            vv[k] = fmaf(__uint_as_float(acc[jj * 8 + k]), g2s[c], bbs[c]) * rs;
          }
          if (a.has_res) {
            const uint4 raw = *reinterpret_cast<const uint4*>(blk + soff);
            vv[0] += __uint_as_float(raw.x << 16); vv[1] += __uint_as_float(raw.x & 0xffff0000u);
            vv[2] += __uint_as_float(raw.y << 16); vv[3] += __uint_as_float(raw.y & 0xffff0000u);
            vv[4] += __uint_as_float(raw.z << 16); vv[5] += __uint_as_float(raw.z & 0xffff0000u);
            vv[6] += __uint_as_float(raw.w << 16); vv[7] += __uint_as_float(raw.w & 0xffff0000u);
          }
          *reinterpret_cast<uint4*>(blk + soff) =
              make_uint4(pack_bf16x2(vv[0], vv[1]), pack_bf16x2(vv[2], vv[3]), pack_bf16x2(vv[4], vv[5]), pack_bf16x2(vv[6], vv[7]));
        }
      }
      // the accumulator has been read: hand it back to the MMA warp before the stores
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&y_empty[ys]);
        if (row0 < a.M)
          for (int cb = 0; cb < NB; ++cb) tma_store_2d(&tmY, st + cb * 2048, cb * 32, row0);
        tma_store_commit();
        tma_store_wait_read0();  // staging may be overwritten: by the next residual tile or the next results
        if (a.has_res && tl + 1 < my_tiles) issue_res(tl + 1);
      }
      __syncwarp();
    }
    if (lane == 0) tma_store_wait_all();
  } else if (warp < 6 + CF::NGW) {
    // ===================== GELU warps: pre (fp32, TMEM) -> h = gelu(pre + b1) (bf16 pairs, same TMEM columns) =====================
    const int q = warp & 3;
    const int slice = (warp - 6) >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    for (uint32_t g = 0; g < total_chunks; ++g) {
      const uint32_t b = g % NPRE, j = g % NC;
      mbar_wait(&pre_full[b], (g / NPRE) & 1u);
      tcgen05_fence_after();
      const uint32_t tbuf = tmem_base + b * HC + lane_off;
      // the TMEM read of the next 32-column block is in flight while this one is computed (its ~100-cycle latency is otherwise exposed)
      constexpr int NBLK = CF::CPS / 32;
      uint32_t acc[NBLK > 1 ? 2 : 1][32];
      tmem_ld32_nowait(tbuf + slice * CF::CPS, acc[0]);
#pragma unroll
      for (int cb = 0; cb < NBLK; ++cb) {
        const int col = slice * CF::CPS + cb * 32;
        tmem_ld_wait();
        if (cb + 1 < NBLK) tmem_ld32_nowait(tbuf + col + 32, acc[(cb + 1) & 1]);
        const uint32_t* ac = acc[cb & 1];
        const float* bp = b1s + j * HC + col;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          if (LNX_DBG & 4) {
            pk[i] = ac[2 * i] ^ ac[2 * i + 1];
            pk[i + 1] = ac[2 * i + 2] ^ ac[2 * i + 3];
            continue;
          }
          const float4 bv = (LNX_DBG & 1) ? make_float4(0.1f, 0.2f, 0.3f, 0.4f) : *reinterpret_cast<const float4*>(bp + 2 * i);
          // h holds 2 gelu(pre): the output warps fold the 0.5 into gamma
          const float2 r0 = gelu2q_x2(__fadd2_rn(make_float2(__uint_as_float(ac[2 * i]), __uint_as_float(ac[2 * i + 1])), make_float2(bv.x, bv.y)));
          const float2 r1 = gelu2q_x2(__fadd2_rn(make_float2(__uint_as_float(ac[2 * i + 2]), __uint_as_float(ac[2 * i + 3])), make_float2(bv.z, bv.w)));
          pk[i] = pack_bf16x2(r0.x, r0.y);
          pk[i + 1] = pack_bf16x2(r1.x, r1.y);
        }
        tmem_st16_u32(tbuf + col, pk);  // k = col + 2 i, col + 2 i + 1 -> column col + i (inside this warp's own slice)
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_full[b]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <class CF>
int launch_fwd(const void* x, const void* w1, const void* w2, const void* residual, void* y, const FusedArgs& a, cudaStream_t st) {
  constexpr int C = CF::C, H = CF::H;
  CUtensorMap tmX, tmXr, tmW1, tmW1r, tmW2, tmRes, tmY;
  bool ok = tmap_k(&tmX, x, C, a.M, C, 64, BM) && tmap_k(&tmW1, w1, C, H, C, 64, CF::HC) && tmap_k(&tmW2, w2, H, C, H, 64, C) &&
            tmap_k(&tmY, y, C, a.M, C, 32, 32);
  if (CF::KREM) ok = ok && tmap_k(&tmXr, x, C, a.M, C, 32, BM) && tmap_k(&tmW1r, w1, C, H, C, 32, CF::HC);
  else tmXr = tmX, tmW1r = tmW1;
  if (residual) ok = ok && tmap_k(&tmRes, residual, C, a.M, C, 32, 32);
  else tmRes = tmY;
  if (!ok) return LNX_ERR_UNSUPPORTED;
  auto kern = mlp_fused_fwd_kernel<CF>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    attr_set = true;
  }
  const int grid = min(a.num_tiles, kNumSMs);
  kern<<<grid, CF::NTHREADS, CF::SMEM, st>>>(tmX, tmXr, tmW1, tmW1r, tmW2, tmRes, tmY, a);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

}  // namespace

// x [M, C], w1 [H, C], w2 [C, H], residual / y [M, C]: bf16 row-major; b1 [H], b2 [C], gamma [C], row_scale: float32.
extern "C" int lnx_mlp_fused_fwd(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const float* gamma,
                                 const float* row_scale, int rows_per_group, const void* residual, void* y, int64_t M, int C, int H,
                                 lnx_stream_t s) {
  LNX_REQUIRE(x && w1 && w2 && y, LNX_ERR_NULL);
  LNX_REQUIRE(M > 0 && M < (1ll << 31) - BM, LNX_ERR_SHAPE);
  if (H != 4 * C || (C != 96 && C != 192)) return LNX_ERR_UNSUPPORTED;
  if (!lnx_aligned16(x) || !lnx_aligned16(w1) || !lnx_aligned16(w2) || !lnx_aligned16(y) || !lnx_aligned16(residual)) return LNX_ERR_ALIGN;
  if (row_scale && rows_per_group <= 0) return LNX_ERR_SHAPE;
  FusedArgs a;
  a.b1 = b1; a.b2 = b2; a.gamma = gamma; a.row_scale = row_scale;
  a.rows_per_group = rows_per_group > 0 ? rows_per_group : 1;
  a.M = (int)M;
  a.has_res = residual != nullptr;
  a.num_tiles = (int)((M + BM - 1) / BM);
  cudaStream_t st = (cudaStream_t)s;
  static const int ngw = getenv("LNX_MLP_NGW") ? atoi(getenv("LNX_MLP_NGW")) : 8;  // GELU warps: 8 measured 0.154 ms, 16 0.162 ms (MUFU bound either way)
  if (C == 96) return ngw == 8 ? launch_fwd<Cfg<96, 128, true, 8>>(x, w1, w2, residual, y, a, st) : launch_fwd<Cfg<96, 128, true, 16>>(x, w1, w2, residual, y, a, st);
  return launch_fwd<Cfg<192, 64, false, 8>>(x, w1, w2, residual, y, a, st);
}
