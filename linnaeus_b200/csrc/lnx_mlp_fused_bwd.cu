// Backward data path of the fused ConvNeXt pointwise pair (autograd of convnext.py:79-86) on tcgen05, with the
// pre-activation RECOMPUTED from x instead of loaded (the forward kernel saves nothing but its input):
//
//   pre  = x W1^T + b1                      MMA1   (A = x tile, B = W1 chunk, both K-major from shared memory)
//   dH   = dY W2e                           MMA2   (A = dY tile K-major, B = the same resident W2 tile read MN-major)
//   h    = gelu(pre),  dPre = dH * gelu'(pre)        one MUFU.TANH serves both
//   dX   = dPre W1                          MMA3   (A = bf16 dPre written back into dH's TMEM columns, B = the resident
//                                                   W1 tile read MN-major: no transposed weight copy anywhere)
// Outputs: dX [M, C], and h / dPre [M, 4C] for the two weight-gradient GEMMs (lnx_wgrad).  W2e = W2 * gamma[:, None]
// (layer scale folded by the caller).  C = 96: W1 and W2e stay resident in shared memory (147 KB) for the whole kernel.
//
// One persistent CTA per SM, 128-row tiles, hidden chunks of 64.  Warp 0 TMA producer, warp 1 MMA issuer (front =
// MMA1 + MMA2 of chunk i, issued two chunks ahead of back = MMA3 of chunk i - 2; three TMEM stages), warps 2-9
// epilogue (two per TMEM lane quarter, 32 hidden columns each): TMEM -> registers -> {TMEM (dPre for MMA3),
// swizzled staging -> TMA store (h, dPre)}; the dX tile of a finished row tile leaves through the same staging
// one chunk later, so the epilogue warps never wait for the tensor pipe at a tile boundary.
#include "lnx_mlp_fused.cuh"

using namespace lnx;
using namespace lnx_tc;
using namespace lnx_mlp;

namespace {

struct BwdArgs {
  const float* b1;
  int M;
  int num_tiles;
};

template <int C_>
struct BCfg {
  static constexpr int C = C_;
  static constexpr int H = 4 * C;
  static constexpr int HC = 64;
  static constexpr int NC = H / HC;
  static constexpr int KB64 = C / 64, KREM = C % 64;
  static constexpr int NST = 3;                       // TMEM stages of {pre 64 cols, dH 64 cols}
  static constexpr int DX_COL = NST * 2 * HC;          // dX accumulator: C columns
  static constexpr int NEW = 8;                        // epilogue warps
  static constexpr int W_BYTES = H * C * 2;            // W1 and W2e: 73 728 each at C = 96
  static constexpr int T_BYTES = BM * C * 2;           // one x or dY tile
  static constexpr int STG_BYTES = NEW * 2 * 2048;     // per warp two [32 rows][32 cols] blocks (64-byte swizzle)
  static constexpr int NBAR = 16;
  static constexpr int SMEM = 2 * W_BYTES + 2 * T_BYTES + STG_BYTES + H * 4 + NBAR * 8 + 16 + 1024;
  static constexpr int NTHREADS = 32 * (2 + NEW);
  static_assert(KB64 == 1 && KREM == 32, "resident-weight backward is laid out for C = 96");
  static_assert(DX_COL + C <= 512, "TMEM budget");
  static_assert(SMEM <= MAX_SMEM, "shared memory budget");
};

template <class CF>
__global__ void __launch_bounds__(CF::NTHREADS, 1)
    mlp_fused_bwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXr, const __grid_constant__ CUtensorMap tmDy,
                         const __grid_constant__ CUtensorMap tmDyr, const __grid_constant__ CUtensorMap tmW1,
                         const __grid_constant__ CUtensorMap tmW1r, const __grid_constant__ CUtensorMap tmW2,
                         const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmDpre,
                         const __grid_constant__ CUtensorMap tmDx, const BwdArgs a) {
  constexpr int C = CF::C, H = CF::H, HC = CF::HC, NC = CF::NC, NST = CF::NST;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* w1s = base;                         // [H][64] 128B swizzle | [H][32] 64B swizzle
  unsigned char* w2s = w1s + CF::W_BYTES;            // H / 64 blocks of [C][64] 128B swizzle
  unsigned char* xs = w2s + CF::W_BYTES;             // [128][64] | [128][32]
  unsigned char* dys = xs + CF::T_BYTES;
  unsigned char* stg = dys + CF::T_BYTES;
  float* b1s = reinterpret_cast<float*>(stg + CF::STG_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(b1s + H);
  uint64_t* xy_full = bars;          // [1]
  uint64_t* xy_empty = bars + 1;     // [1]
  uint64_t* w_full = bars + 2;       // [1]
  uint64_t* st_full = bars + 3;      // [NST]  MMA1 + MMA2 of a chunk complete
  uint64_t* ep_done = bars + 6;      // [NST]  dPre written back to TMEM by every epilogue warp
  uint64_t* dx_full = bars + 9;      // [1]
  uint64_t* dx_empty = bars + 10;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + CF::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = ((int)blockIdx.x < a.num_tiles) ? (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t total = (uint32_t)my_tiles * NC;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmXr); prefetch_tmap(&tmDy); prefetch_tmap(&tmDyr);
    prefetch_tmap(&tmW1); prefetch_tmap(&tmW1r); prefetch_tmap(&tmW2);
    prefetch_tmap(&tmH); prefetch_tmap(&tmDpre); prefetch_tmap(&tmDx);
    mbar_init(xy_full, 1);
    mbar_init(xy_empty, 1);
    mbar_init(w_full, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&st_full[i], 1);
      mbar_init(&ep_done[i], CF::NEW);
    }
    mbar_init(dx_full, 1);
    mbar_init(dx_empty, CF::NEW);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < H; i += CF::NTHREADS) b1s[i] = a.b1 ? a.b1[i] : 0.f;
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, 2u * CF::W_BYTES);
      for (int r = 0; r < H; r += 128) {
        tma_load_2d(w1s + r * 128, &tmW1, w_full, 0, r);
        tma_load_2d(w1s + H * 128 + r * 64, &tmW1r, w_full, 64, r);
      }
      for (int kb = 0; kb < H / 64; ++kb) tma_load_2d(w2s + kb * (C * 128), &tmW2, w_full, kb * 64, 0);
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int m0 = ((int)blockIdx.x + tl * (int)gridDim.x) * BM;
        mbar_wait_relaxed(xy_empty, ((uint32_t)tl & 1u) ^ 1u);
        mbar_expect_tx(xy_full, 2u * CF::T_BYTES);
        tma_load_2d(xs, &tmX, xy_full, 0, m0);
        tma_load_2d(xs + BM * 128, &tmXr, xy_full, 64, m0);
        tma_load_2d(dys, &tmDy, xy_full, 0, m0);
        tma_load_2d(dys + BM * 128, &tmDyr, xy_full, 64, m0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && total > 0) {
      const uint32_t id_pre = make_idesc_bf16(BM, HC, 0, 0);
      const uint32_t id_dh = make_idesc_bf16(BM, HC, 0, 1);
      const uint32_t id_dx64 = make_idesc_bf16(BM, 64, 0, 1);
      const uint32_t id_dx32 = make_idesc_bf16(BM, 32, 0, 1);
      const uint32_t xa = smem_u32(xs), xr = xa + BM * 128;
      const uint32_t da = smem_u32(dys), dr = da + BM * 128;
      const uint32_t w1 = smem_u32(w1s), w1r = w1 + H * 128, w2 = smem_u32(w2s);
      const uint32_t tdx = tmem_base + CF::DX_COL;
      constexpr uint32_t LA = NST - 1;
      mbar_wait_relaxed(w_full, 0);
      tcgen05_fence_after();
      for (uint32_t i = 0; i < total + LA; ++i) {
        if (i < total) {
          // ---- front: pre = x W1[chunk]^T and dH = dY W2e[:, chunk]
          const uint32_t tl = i / NC, j = i % NC, s = i % NST;
          if (j == 0) {
            mbar_wait_relaxed(xy_full, tl & 1u);
            tcgen05_fence_after();
          }
          const uint32_t tpre = tmem_base + s * (2 * HC), tdh = tpre + HC;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tpre, desc_k128(xa + k * 32), desc_k128(w1 + j * HC * 128 + k * 32), id_pre, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_bf16(tpre, desc_k64(xr + k * 32), desc_k64(w1r + j * HC * 64 + k * 32), id_pre, 1u);
          const uint32_t w2b = w2 + j * (C * 128);  // [C rows = k][64 hidden = n]
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tdh, desc_k128(da + k * 32), desc_mn128(w2b + k * 2048, C * 128), id_dh, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_bf16(tdh, desc_k64(dr + k * 32), desc_mn128(w2b + (4 + k) * 2048, C * 128), id_dh, 1u);
          umma_commit(&st_full[s]);
          if (j == NC - 1) umma_commit(xy_empty);
        }
        if (i >= LA) {
          // ---- back: dX += dPre W1[chunk, :], dPre read from TMEM
          const uint32_t ib = i - LA;
          const uint32_t tl = ib / NC, j = ib % NC, s = ib % NST;
          mbar_wait(&ep_done[s], (ib / NST) & 1u);
          tcgen05_fence_after();
          if (j == 0) {
            mbar_wait(dx_empty, (tl & 1u) ^ 1u);
            tcgen05_fence_after();
          }
          const uint32_t tdp = tmem_base + s * (2 * HC) + HC;
#pragma unroll
          for (int kk = 0; kk < HC / 16; ++kk) {
            const uint32_t ta = tdp + 32 * (kk >> 1) + 8 * (kk & 1);
            const uint32_t row = j * HC + 16 * kk;  // hidden rows of W1 = the k dimension
            const uint32_t acc = (j > 0 || kk > 0) ? 1u : 0u;
            umma_bf16_ts(tdx, ta, desc_mn128(w1 + row * 128, H * 128), id_dx64, acc);
            umma_bf16_ts(tdx + 64, ta, desc_mn64(w1r + row * 64, H * 64), id_dx32, acc);
          }
          if (j == NC - 1) umma_commit(dx_full);
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int e = warp - 2;
    const int q = warp & 3;       // TMEM lane quarter
    const int hf = e >> 2;        // which 32 of the chunk's 64 hidden columns
    unsigned char* stA = stg + e * 4096;  // h block, later dX blocks
    unsigned char* stB = stA + 2048;      // dPre block
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t sw = (uint32_t)((lane >> 1) & 3);

    auto store_dx = [&](int tl) {
      // dX tile of row tile tl: warps with hf = 0 take columns 0-63, hf = 1 columns 64-95
      const int row0 = ((int)blockIdx.x + tl * (int)gridDim.x) * BM + q * 32;
      mbar_wait(dx_full, (uint32_t)tl & 1u);
      tcgen05_fence_after();
      if (lane == 0) tma_store_wait_read0();
      __syncwarp();
      const int nblk = hf == 0 ? 2 : 1;
      for (int b = 0; b < nblk; ++b) {
        const int cb = hf == 0 ? b : 2;
        uint32_t acc[32];
        tmem_ld32_nowait(tmem_base + CF::DX_COL + cb * 32 + lane_off, acc);
        tmem_ld_wait();
        unsigned char* blk = b == 0 ? stA : stB;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint32_t soff = (uint32_t)(lane * 64) + ((((uint32_t)jj) ^ sw) << 4);
          *reinterpret_cast<uint4*>(blk + soff) =
              make_uint4(pack_bf16x2(__uint_as_float(acc[jj * 8 + 0]), __uint_as_float(acc[jj * 8 + 1])),
                         pack_bf16x2(__uint_as_float(acc[jj * 8 + 2]), __uint_as_float(acc[jj * 8 + 3])),
                         pack_bf16x2(__uint_as_float(acc[jj * 8 + 4]), __uint_as_float(acc[jj * 8 + 5])),
                         pack_bf16x2(__uint_as_float(acc[jj * 8 + 6]), __uint_as_float(acc[jj * 8 + 7])));
        }
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(dx_empty);
        if (row0 < a.M) {
          if (hf == 0) {
            tma_store_2d(&tmDx, stA, 0, row0);
            tma_store_2d(&tmDx, stB, 32, row0);
          } else {
            tma_store_2d(&tmDx, stA, 64, row0);
          }
        }
        tma_store_commit();
      }
      __syncwarp();
    };

    for (uint32_t i = 0; i < total; ++i) {
      const uint32_t tl = i / NC, j = i % NC, s = i % NST;
      const int row0 = ((int)blockIdx.x + (int)tl * (int)gridDim.x) * BM + q * 32;
      mbar_wait(&st_full[s], (i / NST) & 1u);
      tcgen05_fence_after();
      const uint32_t tpre = tmem_base + s * (2 * HC) + 32 * hf + lane_off;
      const uint32_t tdh = tpre + HC;
      uint32_t ap[32], ad[32];
      tmem_ld32_nowait(tpre, ap);
      tmem_ld32_nowait(tdh, ad);
      tmem_ld_wait();
      const float* bp = b1s + j * HC + 32 * hf;
      uint32_t hp[16], dp[16];
#pragma unroll
      for (int p = 0; p < 16; p += 2) {
        const float4 bv = *reinterpret_cast<const float4*>(bp + 2 * p);
        float2 g0, d0, g1, d1;
        gelu2q_both_x2(__fadd2_rn(make_float2(__uint_as_float(ap[2 * p]), __uint_as_float(ap[2 * p + 1])), make_float2(bv.x, bv.y)), g0, d0);
        gelu2q_both_x2(__fadd2_rn(make_float2(__uint_as_float(ap[2 * p + 2]), __uint_as_float(ap[2 * p + 3])), make_float2(bv.z, bv.w)), g1, d1);
        g0 = __fmul2_rn(g0, f2(0.5f));
        g1 = __fmul2_rn(g1, f2(0.5f));
        d0 = __fmul2_rn(d0, make_float2(__uint_as_float(ad[2 * p]), __uint_as_float(ad[2 * p + 1])));
        d1 = __fmul2_rn(d1, make_float2(__uint_as_float(ad[2 * p + 2]), __uint_as_float(ad[2 * p + 3])));
        hp[p] = pack_bf16x2(g0.x, g0.y);
        hp[p + 1] = pack_bf16x2(g1.x, g1.y);
        dp[p] = pack_bf16x2(d0.x, d0.y);
        dp[p + 1] = pack_bf16x2(d1.x, d1.y);
      }
      // dPre (bf16 pairs) back into this warp's own dH columns: the A operand of the dX MMA
      tmem_st16_u32(tdh, dp);
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ep_done[s]);
        tma_store_wait_read0();  // the stores of the previous chunk have read the staging blocks
      }
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const uint32_t soff = (uint32_t)(lane * 64) + ((((uint32_t)jj) ^ sw) << 4);
        *reinterpret_cast<uint4*>(stA + soff) = make_uint4(hp[jj * 4], hp[jj * 4 + 1], hp[jj * 4 + 2], hp[jj * 4 + 3]);
        *reinterpret_cast<uint4*>(stB + soff) = make_uint4(dp[jj * 4], dp[jj * 4 + 1], dp[jj * 4 + 2], dp[jj * 4 + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (row0 < a.M) {
          const int col0 = (int)j * HC + 32 * hf;
          tma_store_2d(&tmH, stA, col0, row0);
          tma_store_2d(&tmDpre, stB, col0, row0);
        }
        tma_store_commit();
      }
      __syncwarp();
      if (j == 0 && tl > 0) store_dx((int)tl - 1);
    }
    if (my_tiles > 0) store_dx(my_tiles - 1);
    if (lane == 0) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// x, dy, dx [M, C]; w1 [H, C]; w2e [C, H] (= W2 * gamma[:, None]); h, dpre [M, H]: bf16 row-major.  b1 [H] float32.
extern "C" int lnx_mlp_fused_bwd(const void* x, const void* dy, const void* w1, const float* b1, const void* w2e, void* h, void* dpre,
                                 void* dx, int64_t M, int C, int H, lnx_stream_t s) {
  LNX_REQUIRE(x && dy && w1 && w2e && h && dpre && dx, LNX_ERR_NULL);
  LNX_REQUIRE(M > 0 && M < (1ll << 31) - BM, LNX_ERR_SHAPE);
  if (H != 4 * C || C != 96) return LNX_ERR_UNSUPPORTED;
  if (!lnx_aligned16(x) || !lnx_aligned16(dy) || !lnx_aligned16(w1) || !lnx_aligned16(w2e) || !lnx_aligned16(h) || !lnx_aligned16(dpre) ||
      !lnx_aligned16(dx))
    return LNX_ERR_ALIGN;
  using CF = BCfg<96>;
  BwdArgs a;
  a.b1 = b1;
  a.M = (int)M;
  a.num_tiles = (int)((M + BM - 1) / BM);
  CUtensorMap tmX, tmXr, tmDy, tmDyr, tmW1, tmW1r, tmW2, tmH, tmDpre, tmDx;
  const bool ok = tmap_k(&tmX, x, C, M, C, 64, BM) && tmap_k(&tmXr, x, C, M, C, 32, BM) && tmap_k(&tmDy, dy, C, M, C, 64, BM) &&
                  tmap_k(&tmDyr, dy, C, M, C, 32, BM) && tmap_k(&tmW1, w1, C, H, C, 64, 128) && tmap_k(&tmW1r, w1, C, H, C, 32, 128) &&
                  tmap_k(&tmW2, w2e, H, C, H, 64, C) && tmap_k(&tmH, h, H, M, H, 32, 32) && tmap_k(&tmDpre, dpre, H, M, H, 32, 32) &&
                  tmap_k(&tmDx, dx, C, M, C, 32, 32);
  if (!ok) return LNX_ERR_UNSUPPORTED;
  auto kern = mlp_fused_bwd_kernel<CF>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    attr_set = true;
  }
  const int grid = min(a.num_tiles, kNumSMs);
  kern<<<grid, CF::NTHREADS, CF::SMEM, (cudaStream_t)s>>>(tmX, tmXr, tmDy, tmDyr, tmW1, tmW1r, tmW2, tmH, tmDpre, tmDx, a);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
