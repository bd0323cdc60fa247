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
  int store_hd;  // 0: only dX leaves the kernel (the weight gradients come from mlp_fused_wgrad_kernel)
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
      if (a.store_hd) {
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
      }
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
  LNX_REQUIRE(x && dy && w1 && w2e && dx && ((h && dpre) || (!h && !dpre)), LNX_ERR_NULL);
  LNX_REQUIRE(M > 0 && M < (1ll << 31) - BM, LNX_ERR_SHAPE);
  if (H != 4 * C || C != 96) return LNX_ERR_UNSUPPORTED;
  if (!lnx_aligned16(x) || !lnx_aligned16(dy) || !lnx_aligned16(w1) || !lnx_aligned16(w2e) || !lnx_aligned16(h) || !lnx_aligned16(dpre) ||
      !lnx_aligned16(dx))
    return LNX_ERR_ALIGN;  // (NULL h / dpre count as aligned)
  using CF = BCfg<96>;
  BwdArgs a;
  a.b1 = b1;
  a.M = (int)M;
  a.num_tiles = (int)((M + BM - 1) / BM);
  a.store_hd = h != nullptr;
  CUtensorMap tmX, tmXr, tmDy, tmDyr, tmW1, tmW1r, tmW2, tmH, tmDpre, tmDx;
  bool ok = tmap_k(&tmX, x, C, M, C, 64, BM) && tmap_k(&tmXr, x, C, M, C, 32, BM) && tmap_k(&tmDy, dy, C, M, C, 64, BM) &&
            tmap_k(&tmDyr, dy, C, M, C, 32, BM) && tmap_k(&tmW1, w1, C, H, C, 64, 128) && tmap_k(&tmW1r, w1, C, H, C, 32, 128) &&
            tmap_k(&tmW2, w2e, H, C, H, 64, C) && tmap_k(&tmDx, dx, C, M, C, 32, 32);
  if (h) ok = ok && tmap_k(&tmH, h, H, M, H, 32, 32) && tmap_k(&tmDpre, dpre, H, M, H, 32, 32);
  else tmH = tmDx, tmDpre = tmDx;
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

// =====================================================================================================================
// Weight-gradient kernel of the fused pointwise pair: dW1, db1 and the un-scaled dW2, db2 accumulated ON CHIP.
//
// The two [4C, C] fp32 weight gradients (2 x 147 KB at C = 96) do not fit into one SM's 256 KB of TMEM next to the working
// accumulators, so the hidden dimension is split three ways: CTA (p, r) owns hidden units [128 p, 128 p + 128) and walks the
// row tiles r, r + G, r + 2G, ...  Per tile it recomputes its third of the pre-activation and of dH, turns them into h and dPre
// (one tanh serves both), stages both as bf16 MN-major A operands in shared memory and accumulates
//     dW1[p]   (128 x C) += dPre^T x        dW2^T[p] (128 x C) += h^T dY
// in TMEM across ALL its tiles; one flush with fp32 atomics at the end.  The x / dY tiles already in shared memory are the B
// operands (read MN-major).  Nothing 4C-wide ever touches HBM: traffic = x + dY (read by the three CTAs of a row tile, the
// second and third time from L2).  db1 = colsum(dPre) rides on the dW1 MMAs as 16 extra columns against a tile of ones; db2 =
// colsum(dY) needs no hidden tensor and stays with lnx_colsum.
// =====================================================================================================================
namespace {

struct WArgs {
  const float* b1;
  float* dw1;      // [H, C]  +=
  float* db1;      // [H]     +=
  float* dw2;      // [C, H]  +=  (un-scaled: the caller applies the layer scale)
  int M, num_tiles, H;
};

template <int C_>
struct WCfg {
  static constexpr int C = C_;
  static constexpr int HP = 128;                       // hidden units per CTA
  static constexpr int HC = 64;                        // sub-chunk of one MMA front
  static constexpr int NEW = 8;                        // epilogue warps
  static constexpr int W1_BYTES = HP * C * 2;          // [128 hid][64] + [128 hid][32]
  static constexpr int W2_BYTES = C * HP * 2;          // two blocks of [C][64 hid]
  static constexpr int T_BYTES = BM * C * 2;
  static constexpr int STG_BYTES = BM * HP * 2;        // [128 rows][128 hid] bf16 = two [128][64] blocks, per tensor
  static constexpr int ONES_BYTES = BM * 64;           // [128 rows][32] bf16 of 1.0: B operand whose MMA column sums dPre (= db1)
  static constexpr int SMEM = W1_BYTES + W2_BYTES + 4 * T_BYTES + 2 * STG_BYTES + ONES_BYTES + HP * 4 + 40 * 8 + 16 + 1024;
  static constexpr int NTHREADS = 32 * (2 + NEW);
  static constexpr int ACC1 = 256, ACC2 = 384;         // TMEM columns: dW1 (96) + db1 (16) at 256, dW2^T (96) at 384
  static_assert(C == 96, "laid out for C = 96");
  static_assert(SMEM <= MAX_SMEM, "shared memory budget");
};

template <class CF>
__global__ void __launch_bounds__(CF::NTHREADS, 1)
    mlp_fused_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXr, const __grid_constant__ CUtensorMap tmDy,
                           const __grid_constant__ CUtensorMap tmDyr, const __grid_constant__ CUtensorMap tmW1,
                           const __grid_constant__ CUtensorMap tmW1r, const __grid_constant__ CUtensorMap tmW2, const WArgs a) {
  constexpr int C = CF::C, HP = CF::HP, HC = CF::HC;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* w1s = base;                          // [128 hid][64 c] 128B swizzle | [128 hid][32 c] 64B swizzle
  unsigned char* w2s = w1s + CF::W1_BYTES;            // [C][64 hid] x 2, 128B swizzle
  unsigned char* xs = w2s + CF::W2_BYTES;             // 2 x {[128 rows][64] | [128 rows][32]}
  unsigned char* dys = xs + 2 * CF::T_BYTES;
  unsigned char* hst = dys + 2 * CF::T_BYTES;         // h   staging: [128 rows][64 hid] x 2 blocks, 128B swizzle
  unsigned char* pst = hst + CF::STG_BYTES;           // dPre staging
  unsigned char* ones = pst + CF::STG_BYTES;          // all 1.0: every layout of an all-ones tile is the same tile
  float* b1s = reinterpret_cast<float*>(ones + CF::ONES_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(b1s + HP);
  uint64_t* xy_full = bars;        // [2]
  uint64_t* xy_empty = bars + 2;   // [2]  MMA commit + column-sum warp
  uint64_t* w_full = bars + 4;     // [1]
  uint64_t* st_full = bars + 5;    // [2]  pre + dH of a sub-chunk complete
  uint64_t* st_empty = bars + 7;   // [2]  epilogue has read them
  uint64_t* stg_full = bars + 9;   // [1]  h / dPre of a whole tile staged (16 arrivals)
  uint64_t* stg_free = bars + 10;  // [1]  weight-gradient MMAs of the tile have read the staging
  uint64_t* acc_done = bars + 11;  // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x % 3;                       // hidden third
  const int r = blockIdx.x / 3, G = gridDim.x / 3;    // row-tile walker
  const int my_tiles = (r < a.num_tiles) ? (a.num_tiles - r + G - 1) / G : 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmXr); prefetch_tmap(&tmDy); prefetch_tmap(&tmDyr);
    prefetch_tmap(&tmW1); prefetch_tmap(&tmW1r); prefetch_tmap(&tmW2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&xy_full[i], 1);
      mbar_init(&xy_empty[i], 1);
      mbar_init(&st_full[i], 1);
      mbar_init(&st_empty[i], CF::NEW);
    }
    mbar_init(w_full, 1);
    mbar_init(stg_full, 2 * CF::NEW);
    mbar_init(stg_free, 1);
    mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < HP; i += CF::NTHREADS) b1s[i] = a.b1 ? a.b1[p * HP + i] : 0.f;
  for (int i = threadIdx.x; i < CF::ONES_BYTES / 4; i += CF::NTHREADS) reinterpret_cast<uint32_t*>(ones)[i] = 0x3f803f80u;  // bf16 1.0 pairs
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0 && my_tiles > 0) {
      mbar_expect_tx(w_full, (uint32_t)(CF::W1_BYTES + CF::W2_BYTES));
      tma_load_2d(w1s, &tmW1, w_full, 0, p * HP);
      tma_load_2d(w1s + HP * 128, &tmW1r, w_full, 64, p * HP);
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(w2s + kb * (C * 128), &tmW2, w_full, p * HP + kb * 64, 0);
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int m0 = (r + tl * G) * BM;
        const int s = tl & 1;
        mbar_wait_relaxed(&xy_empty[s], (((uint32_t)tl >> 1) & 1u) ^ 1u);
        mbar_expect_tx(&xy_full[s], 2u * CF::T_BYTES);
        tma_load_2d(xs + s * CF::T_BYTES, &tmX, &xy_full[s], 0, m0);
        tma_load_2d(xs + s * CF::T_BYTES + BM * 128, &tmXr, &xy_full[s], 64, m0);
        tma_load_2d(dys + s * CF::T_BYTES, &tmDy, &xy_full[s], 0, m0);
        tma_load_2d(dys + s * CF::T_BYTES + BM * 128, &tmDyr, &xy_full[s], 64, m0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && my_tiles > 0) {
      const uint32_t id_pre = make_idesc_bf16(BM, HC, 0, 0);   // A = x (K-major), B = W1 rows (K-major)
      const uint32_t id_dh = make_idesc_bf16(BM, HC, 0, 1);    // A = dY (K-major), B = W2e tile read MN-major
      const uint32_t id_w64 = make_idesc_bf16(HP, 64, 1, 1);   // A = staged h / dPre (MN-major), B = x / dY tile (MN-major), c 0..63
      const uint32_t id_w32 = make_idesc_bf16(HP, 32, 1, 1);   //                                                           c 64..95
      const uint32_t id_w16 = make_idesc_bf16(HP, 16, 1, 1);   // B = ones: 16 identical columns = colsum over the rows of dPre (db1)
      const uint32_t on1 = smem_u32(ones);
      const uint32_t w1 = smem_u32(w1s), w1r = w1 + HP * 128, w2 = smem_u32(w2s);
      const uint32_t hs = smem_u32(hst), ps = smem_u32(pst);
      mbar_wait_relaxed(w_full, 0);
      tcgen05_fence_after();
      auto front = [&](uint32_t tl, uint32_t s) {
        const uint32_t sub = 2 * tl + s;  // global sub-chunk counter: TMEM stage = s
        mbar_wait_relaxed(&st_empty[s], ((sub >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t xa = smem_u32(xs) + (tl & 1) * CF::T_BYTES, xr = xa + BM * 128;
        const uint32_t da = smem_u32(dys) + (tl & 1) * CF::T_BYTES, dr = da + BM * 128;
        const uint32_t tpre = tmem_base + s * (2 * HC), tdh = tpre + HC;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tpre, desc_k128(xa + k * 32), desc_k128(w1 + s * HC * 128 + k * 32), id_pre, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_bf16(tpre, desc_k64(xr + k * 32), desc_k64(w1r + s * HC * 64 + k * 32), id_pre, 1u);
        const uint32_t w2b = w2 + s * (C * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tdh, desc_k128(da + k * 32), desc_mn128(w2b + k * 2048, C * 128), id_dh, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_bf16(tdh, desc_k64(dr + k * 32), desc_mn128(w2b + (4 + k) * 2048, C * 128), id_dh, 1u);
        umma_commit(&st_full[s]);
      };
      auto back = [&](uint32_t tl) {
        mbar_wait(stg_full, tl & 1u);
        tcgen05_fence_after();
        const uint32_t xa = smem_u32(xs) + (tl & 1) * CF::T_BYTES, xr = xa + BM * 128;
        const uint32_t da = smem_u32(dys) + (tl & 1) * CF::T_BYTES, dr = da + BM * 128;
        const uint32_t acc = tl > 0 ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < BM / 16; ++k) {  // reduction over the 128 rows of the tile
          const uint32_t on = (acc || k > 0) ? 1u : 0u;
          umma_bf16(tmem_base + CF::ACC1, desc_mn128(ps + k * 2048, BM * 128), desc_mn128(xa + k * 2048, 0), id_w64, on);
          umma_bf16(tmem_base + CF::ACC1 + 64, desc_mn128(ps + k * 2048, BM * 128), desc_mn64(xr + k * 1024, 0), id_w32, on);
          umma_bf16(tmem_base + CF::ACC1 + 96, desc_mn128(ps + k * 2048, BM * 128), desc_mn64(on1 + k * 1024, 0), id_w16, on);
          umma_bf16(tmem_base + CF::ACC2, desc_mn128(hs + k * 2048, BM * 128), desc_mn128(da + k * 2048, 0), id_w64, on);
          umma_bf16(tmem_base + CF::ACC2 + 64, desc_mn128(hs + k * 2048, BM * 128), desc_mn64(dr + k * 1024, 0), id_w32, on);
        }
        umma_commit(stg_free);
        umma_commit(&xy_empty[tl & 1]);
      };
      for (uint32_t tl = 0; tl < (uint32_t)my_tiles; ++tl) {
        mbar_wait_relaxed(&xy_full[tl & 1], (tl >> 1) & 1u);
        tcgen05_fence_after();
        front(tl, 0);
        if (tl > 0) back(tl - 1);
        front(tl, 1);
      }
      back((uint32_t)my_tiles - 1);
      umma_commit(acc_done);
    }
  } else if (warp < 2 + CF::NEW) {
    // ===================== epilogue warps =====================
    const int e = warp - 2;
    const int q = warp & 3;
    const int hf = e >> 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;  // row of the tile = k index of the weight-gradient MMAs
    for (int tl = 0; tl < my_tiles; ++tl) {
      for (int s = 0; s < 2; ++s) {
        const uint32_t sub = 2 * (uint32_t)tl + s;
        mbar_wait(&st_full[s], (sub >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t tpre = tmem_base + s * (2 * HC) + 32 * hf + lane_off;
        uint32_t ap[32], ad[32];
        tmem_ld32_nowait(tpre, ap);
        tmem_ld32_nowait(tpre + HC, ad);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&st_empty[s]);  // the accumulators are in registers: the next front may overwrite them
        const float* bp = b1s + s * HC + 32 * hf;
        uint32_t hp[16], dp[16];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float4 bv = *reinterpret_cast<const float4*>(bp + 2 * i);
          float2 g0, d0, g1, d1;
          gelu2q_both_x2(__fadd2_rn(make_float2(__uint_as_float(ap[2 * i]), __uint_as_float(ap[2 * i + 1])), make_float2(bv.x, bv.y)), g0, d0);
          gelu2q_both_x2(__fadd2_rn(make_float2(__uint_as_float(ap[2 * i + 2]), __uint_as_float(ap[2 * i + 3])), make_float2(bv.z, bv.w)), g1, d1);
          g0 = __fmul2_rn(g0, f2(0.5f));
          g1 = __fmul2_rn(g1, f2(0.5f));
          d0 = __fmul2_rn(d0, make_float2(__uint_as_float(ad[2 * i]), __uint_as_float(ad[2 * i + 1])));
          d1 = __fmul2_rn(d1, make_float2(__uint_as_float(ad[2 * i + 2]), __uint_as_float(ad[2 * i + 3])));
          hp[i] = pack_bf16x2(g0.x, g0.y);
          hp[i + 1] = pack_bf16x2(g1.x, g1.y);
          dp[i] = pack_bf16x2(d0.x, d0.y);
          dp[i + 1] = pack_bf16x2(d1.x, d1.y);
        }
        // rows beyond M were zero filled on load: dPre = 0 there (dH = 0), but h = gelu(b1) is not -> mask it
        if ((r + tl * G) * BM + row >= a.M) {
#pragma unroll
          for (int i = 0; i < 16; ++i) hp[i] = 0u;
        }
        if (s == 0 && tl > 0) mbar_wait(stg_free, ((uint32_t)tl - 1) & 1u);  // the previous tile's weight-gradient MMAs have read the staging
        // MN-major A operand: [k = row][m = hidden], block s holds hidden 64 s .. 64 s + 63 as 128-byte rows, 16-byte chunks swizzled by row
        unsigned char* hb = hst + s * (BM * 128) + row * 128;
        unsigned char* pb = pst + s * (BM * 128) + row * 128;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint32_t ch = ((uint32_t)(4 * hf + jj) ^ (uint32_t)(row & 7)) << 4;
          *reinterpret_cast<uint4*>(hb + ch) = make_uint4(hp[jj * 4], hp[jj * 4 + 1], hp[jj * 4 + 2], hp[jj * 4 + 3]);
          *reinterpret_cast<uint4*>(pb + ch) = make_uint4(dp[jj * 4], dp[jj * 4 + 1], dp[jj * 4 + 2], dp[jj * 4 + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(stg_full);
      }
    }
    // ---- flush the two TMEM accumulators with atomics
    if (my_tiles > 0) {
      mbar_wait(acc_done, 0);
      tcgen05_fence_after();
      if (e < 4) {  // one warp per lane quarter: lane = hidden row 32 q + lane of this CTA's third
        const int hid = p * HP + q * 32 + lane;
        if (a.db1) {
          uint32_t vb[32];
          tmem_ld32_nowait(tmem_base + CF::ACC1 + 96 + lane_off, vb);  // 16 identical db1 columns (+ 16 unused ones)
          tmem_ld_wait();
          atomicAdd(a.db1 + hid, __uint_as_float(vb[0]));
        }
        for (int cb = 0; cb < 3; ++cb) {
          uint32_t v1[32], v2[32];
          tmem_ld32_nowait(tmem_base + CF::ACC1 + cb * 32 + lane_off, v1);
          tmem_ld32_nowait(tmem_base + CF::ACC2 + cb * 32 + lane_off, v2);
          tmem_ld_wait();
          float* d1p = a.dw1 + (long long)hid * C + cb * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            atomicAdd(reinterpret_cast<float4*>(d1p + i), make_float4(__uint_as_float(v1[i]), __uint_as_float(v1[i + 1]), __uint_as_float(v1[i + 2]),
                                                                      __uint_as_float(v1[i + 3])));
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(a.dw2 + (long long)(cb * 32 + i) * a.H + hid, __uint_as_float(v2[i]));
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// x, dy [M, C]; w1 [H, C]; w2e [C, H]: bf16.  dw1 [H, C], db1 [H], dw2_raw [C, H]: float32, += (atomics).
extern "C" int lnx_mlp_fused_wgrad(const void* x, const void* dy, const void* w1, const float* b1, const void* w2e, float* dw1, float* db1,
                                   float* dw2_raw, int64_t M, int C, int H, lnx_stream_t s) {
  LNX_REQUIRE(x && dy && w1 && w2e && dw1 && dw2_raw, LNX_ERR_NULL);
  LNX_REQUIRE(M > 0 && M < (1ll << 31) - BM, LNX_ERR_SHAPE);
  if (H != 4 * C || C != 96) return LNX_ERR_UNSUPPORTED;
  if (!lnx_aligned16(x) || !lnx_aligned16(dy) || !lnx_aligned16(w1) || !lnx_aligned16(w2e) || !lnx_aligned16(dw1)) return LNX_ERR_ALIGN;
  using CF = WCfg<96>;
  WArgs a;
  a.b1 = b1; a.dw1 = dw1; a.db1 = db1; a.dw2 = dw2_raw;
  a.M = (int)M;
  a.num_tiles = (int)((M + BM - 1) / BM);
  a.H = H;
  CUtensorMap tmX, tmXr, tmDy, tmDyr, tmW1, tmW1r, tmW2;
  const bool ok = tmap_k(&tmX, x, C, M, C, 64, BM) && tmap_k(&tmXr, x, C, M, C, 32, BM) && tmap_k(&tmDy, dy, C, M, C, 64, BM) &&
                  tmap_k(&tmDyr, dy, C, M, C, 32, BM) && tmap_k(&tmW1, w1, C, H, C, 64, 128) && tmap_k(&tmW1r, w1, C, H, C, 32, 128) &&
                  tmap_k(&tmW2, w2e, H, C, H, 64, C);
  if (!ok) return LNX_ERR_UNSUPPORTED;
  auto kern = mlp_fused_wgrad_kernel<CF>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    attr_set = true;
  }
  const int groups = min(a.num_tiles, kNumSMs / 3);  // 49 row-tile walkers x 3 hidden thirds = 147 CTAs
  kern<<<3 * groups, CF::NTHREADS, CF::SMEM, (cudaStream_t)s>>>(tmX, tmXr, tmDy, tmDyr, tmW1, tmW1r, tmW2, a);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
