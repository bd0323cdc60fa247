// Weight-gradient GEMM of a Linear layer on tcgen05:  dW[N, K] += dY[M, N]^T X[M, K]  (reduction over the M
// tokens), with the bias gradient db[N] += colsum(dY) produced by the tensor core in the same pass.
//
// Both operands are consumed MN-major straight from their row-major activations (no transposed copies).  The
// output is tiny and the reduction huge, so the kernel is split-K over all SMs with one fp32 vector-atomic pass at
// the end (the kernel supports MT x NT accumulator tiles per CTA; the host currently picks 1 x 1, see lnx_wgrad).
// db rides on the same MMAs: every B stage is followed in shared memory by a constant 64-column block of ones, and the
// CTAs of the first column group issue their MMAs with N = block_n + 16, so accumulator column block_n is the column
// sum of dY -- no epilogue arithmetic, no separate reduction kernel over dY, +16/block_n tensor time.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue (TMEM lane quarters).
#include <stdio.h>
#include <stdlib.h>

#include "lnx_gemm.cuh"
#include "lnx_tc_common.cuh"

using namespace lnx;
using namespace lnx_tc;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int MAX_SMEM = 232448;
constexpr int BLK_BYTES = BLOCK_K * 128;  // one [32 k][64 mn] swizzled block
constexpr int ONES_BYTES = BLOCK_K * 128;  // [32 k][64 n] bf16 MN-major block of ones appended to every B stage
constexpr int DB_COLS = 16;           // extra accumulator columns (column block_n = colsum(dY))
constexpr int FLUSH_BLK_BYTES = 32 * 128;  // [32 rows][32 float32] staging block of the epilogue flush
constexpr int FLUSH_BUFS = 2;              // per epilogue warp; 4 warps x 2 x 4 KB = 32 KB <= the smallest operand ring (4 in flight measured no faster)

struct WgParams {
  int M, N, K;          // reduction length, dW rows, dW columns
  int mt, nt, block_n;  // accumulator tiles per CTA
  int tile_cols;        // TMEM columns per accumulator tile (block_n, + 16 with db)
  int stages;
  int kb_per_split;
  int groups_n;         // column groups of nt * block_n
  int tmem_cols;
  int has_db;
};

template <int COLS>
__device__ __forceinline__ void alloc_cols(uint32_t* slot) { tmem_alloc<COLS>(slot); }

__global__ void __launch_bounds__(NUM_THREADS) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmW, float* __restrict__ dw,
                                                               float* __restrict__ db, const WgParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = p.mt * BLOCK_M * BLOCK_K * 2;
  const int b_tile_bytes = p.block_n * BLOCK_K * 2;
  const int b_bytes = p.nt * b_tile_bytes;          // bytes TMA delivers per stage
  const int b_stride = b_bytes + ONES_BYTES;        // stage pitch: the ones block sits right behind the B tile
  unsigned char* smem_a = base;
  unsigned char* smem_b = smem_a + (size_t)p.stages * a_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * b_stride);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gn = blockIdx.x % p.groups_n, gm = blockIdx.x / p.groups_n;
  const int m_base = gm * p.mt * BLOCK_M;          // first dW row of this CTA
  const int n_base = gn * p.nt * p.block_n;        // first dW column
  const int num_kb_total = (p.M + BLOCK_K - 1) / BLOCK_K;
  const int kb_begin = blockIdx.y * p.kb_per_split;
  const int num_kb = max(0, min(num_kb_total, kb_begin + p.kb_per_split) - kb_begin);
  const bool do_db = p.has_db && gn == 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_fence_init();
  }
  for (int st = 0; st < p.stages; ++st)
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NUM_THREADS)
      reinterpret_cast<uint32_t*>(smem_b + (size_t)st * b_stride + b_bytes)[i] = 0x3F803F80u;  // bf16 1.0 pairs
  fence_proxy_async_smem();
  if (warp == 1) {
    if (p.tmem_cols == 128) alloc_cols<128>(tmem_slot);
    else if (p.tmem_cols == 256) alloc_cols<256>(tmem_slot);
    else alloc_cols<512>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t stage_bytes = (uint32_t)(a_bytes + b_bytes);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % p.stages;
        mbar_wait(&empty_bar[s], ((uint32_t)(i / p.stages) & 1u) ^ 1u);
        mbar_expect_tx(&full_bar[s], stage_bytes);
        const int k0 = (kb_begin + i) * BLOCK_K;
        unsigned char* sa = smem_a + (size_t)s * a_bytes;
        unsigned char* sb = smem_b + (size_t)s * b_stride;
        for (int j = 0; j < 2 * p.mt; ++j) tma_load_2d(sa + j * BLK_BYTES, &tmA, &full_bar[s], m_base + 64 * j, k0);
        for (int j = 0; j < p.nt * (p.block_n / 64); ++j) tma_load_2d(sb + j * BLK_BYTES, &tmB, &full_bar[s], n_base + 64 * j, k0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BLOCK_M, p.block_n + (do_db ? DB_COLS : 0), 1, 1);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % p.stages;
        mbar_wait(&full_bar[s], (uint32_t)(i / p.stages) & 1u);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem_a + (size_t)s * a_bytes);
        const uint32_t sb = smem_u32(smem_b + (size_t)s * b_stride);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          const uint32_t accf = (i > 0 || k > 0) ? 1u : 0u;
          for (int mt = 0; mt < p.mt; ++mt) {
            const uint64_t ad = make_smem_desc(sa + mt * 2 * BLK_BYTES + k * 2048, BLK_BYTES, 1024);
            for (int nt = 0; nt < p.nt; ++nt) {
              const uint64_t bd = make_smem_desc(sb + nt * b_tile_bytes + k * 2048, BLK_BYTES, 1024);
              umma_bf16(tmem_base + (mt * p.nt + nt) * p.tile_cols, ad, bd, idesc, accf);
            }
          }
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tfull_bar);
    }
  } else if (num_kb > 0) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tcgen05_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    // Split-K partials leave as TMA reduce-adds: each warp stages [32 rows][32 columns] float32 blocks (128-byte rows, 128B swizzle)
    // in the operand ring -- every load has been consumed once tfull_bar fires -- and the L2 adds whole rows; two blocks in flight.
    // (Per-thread float4 atomics were 13-22 us of every launch: 28 MB of 16-byte L2 operations from 148 CTAs at once.)
    unsigned char* stg = base + (size_t)(warp - 2) * (FLUSH_BUFS * FLUSH_BLK_BYTES);
    int nb = 0;
    for (int mt = 0; mt < p.mt; ++mt) {
      const int m = m_base + mt * BLOCK_M + row;
      const int m_blk = m_base + mt * BLOCK_M + q * 32;  // warp-uniform: first dW row of this warp's block
      for (int nt = 0; nt < p.nt; ++nt) {
        const int n0 = n_base + nt * p.block_n;
        for (int c = 0; c < p.block_n; c += 32) {
          if (n0 + c >= p.K || m_blk >= p.N) break;  // warp-uniform
          uint32_t v[32];
          __syncwarp();
          tmem_ld16_nowait(trow + (mt * p.nt + nt) * p.tile_cols + c, v);  // block_n is a multiple of 64
          tmem_ld16_nowait(trow + (mt * p.nt + nt) * p.tile_cols + c + 16, v + 16);
          tmem_ld_wait();
          unsigned char* blk = stg + (nb % FLUSH_BUFS) * FLUSH_BLK_BYTES;
          if (nb >= FLUSH_BUFS) {
            if (lane == 0) bulk_wait_group_read<FLUSH_BUFS - 1>();  // the bulk op that last read this buffer is done with it
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(blk + sw128_chunk(lane, j)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tmW, blk, n0 + c, m_blk);
            bulk_commit_group();
          }
          ++nb;
        }
      }
      if (do_db) {
        float v[16];
        __syncwarp();
        tmem_ld16(trow + (mt * p.nt + p.nt - 1) * p.tile_cols + p.block_n, v);
        if (m < p.N) atomicAdd(db + m, v[0]);
      }
    }
    if (lane == 0) bulk_wait_group_all();
    __syncwarp();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (p.tmem_cols == 128) tmem_dealloc<128>(tmem_base);
    else if (p.tmem_cols == 256) tmem_dealloc<256>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

bool tmap_mn(CUtensorMap* tm, const void* ptr, long long inner, long long outer, long long ld) {
  const long long dims[2] = {inner, outer};
  const long long strides[1] = {ld};
  const int box[2] = {64, BLOCK_K};
  return make_tmap(tm, ptr, 2, dims, strides, box);
}

}  // namespace

// dy [M, N] (pitch ldy), x [M, K] (pitch ldx), both bf16 row-major; dw float [N, K] +=, db float [N] += (nullable)
extern "C" int lnx_wgrad(const void* dy, int64_t ldy, const void* x, int64_t ldx, float* dw, float* db, int64_t M, int N, int K, int dtype,
                         lnx_stream_t s) {
  LNX_REQUIRE(dy && x && dw, LNX_ERR_NULL);
  LNX_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31), LNX_ERR_SHAPE);
  if (dtype != LNX_BF16) return LNX_ERR_DTYPE;
  if (N % 8 != 0 || K % 4 != 0 || ldy % 8 != 0 || ldx % 8 != 0) return LNX_ERR_UNSUPPORTED;
  if (!lnx_aligned16(dy) || !lnx_aligned16(x) || !lnx_aligned16(dw)) return LNX_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)s;

  // Accumulator tiles per CTA (mt x nt tiles of [128 x block_n], TMEM budget 512 columns) -- measured on B200
  // (tools/prof_wgrad2.py sweeps, B = 256 shapes):
  //   K >= 321 (two 192-column tiles, one A tile feeds both)   1 x 2 x 192, one CTA / SM   1536x384: 0.118 -> 0.082 ms
  //   K  = 192                                                  2 x 1 x 192, one CTA / SM    768x192: 0.118 -> 0.099 ms
  //   K  =  96 (HBM bound: every byte of dY and X read once)    3 x 1 x 128, one CTA / SM     384x96: 0.190 -> 0.163 ms
  // Anything else keeps one tile per CTA with two co-resident CTAs per SM.  Bigger footprints lose again: the split-K
  // partials leave through fp32 atomics and that tail grows with the accumulator size per CTA.
  WgParams p;
  p.M = (int)M; p.N = N; p.K = K; p.has_db = db ? 1 : 0;
  const int tiles_m = (N + BLOCK_M - 1) / BLOCK_M;
  const int kpad = ((K + 63) / 64) * 64;
  p.mt = 1; p.nt = 1;
  int ctas_per_sm = 2;
  if (kpad <= 192) p.block_n = kpad;
  else p.block_n = (kpad % 192 == 0) ? 192 : 128;  // <= 192: the db variant issues N = block_n + 16 <= 256
  if (kpad >= 2 * p.block_n) {
    p.nt = 2;
    ctas_per_sm = 1;
  } else if (p.block_n == 192 && tiles_m >= 2) {
    p.mt = 2;
    ctas_per_sm = 1;
  } else if (p.block_n == 128 && tiles_m >= 3) {
    p.mt = 3;
    ctas_per_sm = 1;
  }
  if (const char* e = getenv("LNX_WGRAD_CFG")) {  // experiments: "mt,nt,bn,ctas_per_sm"
    int a, b2, c, d;
    if (sscanf(e, "%d,%d,%d,%d", &a, &b2, &c, &d) == 4) {
      p.mt = a; p.nt = b2; p.block_n = c; ctas_per_sm = d;
    }
  }
  const int groups_m = (tiles_m + p.mt - 1) / p.mt;
  p.groups_n = (kpad + p.nt * p.block_n - 1) / (p.nt * p.block_n);
  p.tile_cols = p.block_n + (db ? DB_COLS : 0);
  const int acc = p.mt * p.nt * p.tile_cols;
  p.tmem_cols = acc <= 128 ? 128 : (acc <= 256 ? 256 : 512);
  const int stage = p.mt * 2 * BLK_BYTES + p.nt * (p.block_n / 64) * BLK_BYTES + ONES_BYTES;
  p.stages = max(2, min(6, ((ctas_per_sm == 2 ? 110 : 220) * 1024 - 4096) / stage));  // <= 110 KB: two CTAs per SM
  const int num_kb = (int)((M + BLOCK_K - 1) / BLOCK_K);
  const int ctas_xy = groups_m * p.groups_n;
  int splits = max(1, min(num_kb / 8, ctas_per_sm == 2 ? (2 * kNumSMs + ctas_xy - 1) / ctas_xy : kNumSMs / ctas_xy));
  p.kb_per_split = (num_kb + splits - 1) / splits;
  splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
  const size_t smem = (size_t)p.stages * stage + 4096;

  CUtensorMap tmA, tmB, tmW;
  if (!tmap_mn(&tmA, dy, N, M, ldy) || !tmap_mn(&tmB, x, K, M, ldx)) return LNX_ERR_UNSUPPORTED;
  if (!make_tmap_f32(&tmW, dw, K, N, K, 32, 32)) return LNX_ERR_UNSUPPORTED;
  static int smem_set = 0;
  if ((int)smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    smem_set = (int)smem;
  }
  wgrad_tc_kernel<<<dim3(ctas_xy, splits), NUM_THREADS, smem, st>>>(tmA, tmB, tmW, dw, db, p);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
