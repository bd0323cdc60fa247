// Persistent, warp-specialised bf16 GEMM (tcgen05 + TMEM + TMA) with a TMA-fed, TMA-stored epilogue.
//
// One CTA per SM walks the output tiles (n fastest).  Roles (64 + 128*G threads, G = 2 or 4 epilogue groups):
//   warp 0      TMA producer: A/B k-blocks into a 128B-swizzled smem ring (full/empty mbarriers)
//   warp 1      MMA issuer: tcgen05.mma cta_group::1, M = 128, N = block_n, accumulators double
//               buffered in TMEM (2 x 256 columns) so the next tile's MMAs overlap this tile's epilogue
//   groups      4 warps each (one TMEM lane quarter per warp); group g owns the 64-column blocks
//               g, g+G, ... of every tile.  Per block: the element-wise input tile (residual or the saved
//               pre-activation) arrives by TMA (prefetched one block ahead), the accumulator block is read
//               with four back-to-back tcgen05.ld, bias / GELU / act' / layer-scale / DropPath mask / residual
//               are applied in registers, results go to a swizzled staging tile and leave with one TMA store
//               (full 128-byte lines, clipped at the M/N edges by the tensor map).
// G = 4 when K is small (the HBM-bound conv-stage GEMMs: the epilogue is the whole cost), G = 2 with a deeper
// operand ring when K is large.  Optional fused column sums of the output (bias gradients) are accumulated per
// CTA in shared memory across its tiles and flushed with one atomic per column.
#include "lnx_gemm.cuh"
#include "lnx_tc_common.cuh"

using namespace lnx;
using namespace lnx_tc;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int MAX_THREADS = 64 + 128 * 4;
constexpr int STAGING_BYTES = 16384;  // [128 rows][64 bf16]
constexpr int TMEM_STAGE_COLS = 256;
constexpr int MAX_SMEM = 232448;

struct Tc2Params {
  int M, N, K;
  int block_n, stages, groups;
  int a_trans, b_trans;
  int tiles_m, tiles_n;
  int has_aux;   // second output (pre-activation) staged in buffer B
  int in_kind;   // 0 none, 1 residual, 2 act_grad_in  -> TMA-loaded into buffer B
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void pack_store_sw(unsigned char* tile, int r, int j, const float* v) {
  uint4 raw;
  __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) p2[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
  *reinterpret_cast<uint4*>(tile + sw128_chunk(r, j)) = raw;
}

__global__ void __launch_bounds__(MAX_THREADS, 1) gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux,
                                                                  const __grid_constant__ CUtensorMap tmIn, const GemmArgs g, const Tc2Params p,
                                                                  float* __restrict__ colsum_out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int b_stage_bytes = p.block_n * BLOCK_K * 2;
  unsigned char* smem_a = base;
  unsigned char* smem_b = smem_a + (size_t)p.stages * A_STAGE_BYTES;
  unsigned char* staging = smem_b + (size_t)p.stages * b_stage_bytes;  // [groups][C tile, B tile][16 KB]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + (size_t)p.groups * 2 * STAGING_BYTES);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;        // [2]
  uint64_t* in_bar = tempty_bar + 2;           // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_bar + 4);
  float* s_vec = reinterpret_cast<float*>(tmem_slot + 4);   // [4 groups][bias, col_scale][256] of the current tile
  float* s_colsum = s_vec + 4 * 2 * 256;                    // [N] when colsum_out

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int nthreads = blockDim.x;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmC);
    if (p.has_aux) prefetch_tmap(&tmAux);
    if (p.in_kind) prefetch_tmap(&tmIn);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4 * p.groups);  // one arrival per epilogue warp
    }
    for (int s = 0; s < 4; ++s) mbar_init(&in_bar[s], 1);
    mbar_fence_init();
  }
  if (colsum_out)
    for (int i = threadIdx.x; i < p.N; i += nthreads) s_colsum[i] = 0.f;
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t stage_bytes = A_STAGE_BYTES + (uint32_t)b_stage_bytes;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / p.tiles_n) * BLOCK_M, n0 = (t % p.tiles_n) * p.block_n;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], stage_bytes);
          const int k0 = kb * BLOCK_K;
          unsigned char* sa = smem_a + (size_t)s * A_STAGE_BYTES;
          unsigned char* sb = smem_b + (size_t)s * b_stage_bytes;
          if (!p.a_trans) {
            tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[s], m0, k0);
            tma_load_2d(sa + 8192, &tmA, &full_bar[s], m0 + 64, k0);
          }
          if (!p.b_trans) {
            tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
          } else {
            for (int j = 0; j < p.block_n / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], n0 + 64 * j, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BLOCK_M, p.block_n, p.a_trans, p.b_trans);
      uint32_t it = 0, tl = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tl) {
        const uint32_t as = tl & 1u, aph = (tl >> 1) & 1u;
        mbar_wait_relaxed(&tempty_bar[as], aph ^ 1u);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + as * TMEM_STAGE_COLS;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1u;
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem_a + (size_t)s * A_STAGE_BYTES);
          const uint32_t sb = smem_u32(smem_b + (size_t)s * b_stage_bytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t ad = p.a_trans ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 0, 1024);
            const uint64_t bd = p.b_trans ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 0, 1024);
            umma_bf16(tacc, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[as]);
      }
    }
  } else {
    // ===================== epilogue =====================
    const int e = warp - 2;
    const int grp = e >> 2;                 // 0 .. groups-1
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;            // tile row
    const int gtid = (e & 3) * 32 + lane;   // 0..127 inside the group
    const int G = p.groups;
    unsigned char* st_c = staging + (size_t)grp * 2 * STAGING_BYTES;
    unsigned char* st_b = st_c + STAGING_BYTES;  // aux output OR element-wise input tile
    float* s_bias = s_vec + grp * 512;
    float* s_scale = s_bias + 256;
    const int n_blocks = (p.block_n + 63) / 64;
    const bool has_items = grp < n_blocks;
    uint32_t tl = 0, in_cnt = 0;

    // prefetch the element-wise input tile of this group's first block
    if (p.in_kind && has_items && gtid == 0 && (int)blockIdx.x < num_tiles) {
      const int t0 = blockIdx.x;
      mbar_expect_tx(&in_bar[grp], STAGING_BYTES);
      tma_load_2d(st_b, &tmIn, &in_bar[grp], (t0 % p.tiles_n) * p.block_n + grp * 64, (t0 / p.tiles_n) * BLOCK_M);
    }

    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tl) {
      const int m0 = (t / p.tiles_n) * BLOCK_M, n0 = (t % p.tiles_n) * p.block_n;
      const uint32_t as = tl & 1u, aph = (tl >> 1) & 1u;
      const int m = m0 + r;
      const bool row_ok = m < p.M;
      if (has_items) {
        // bias / layer-scale of this tile's columns -> shared (read back as broadcasts).  The previous tile's
        // readers of s_bias/s_scale are past their last named barrier, which precedes this point in program order
        // for the writers too (same threads).
        for (int i = gtid; i < p.block_n; i += 128) {
          const int n = n0 + i;
          s_bias[i] = (g.bias && n < p.N) ? g.bias[n] : 0.f;
          s_scale[i] = (g.col_scale && n < p.N) ? g.col_scale[n] : 1.f;
        }
      }
      const float rs = (g.row_scale && row_ok) ? g.row_scale[m / g.rows_per_group] : 1.f;
      mbar_wait(&tfull_bar[as], aph);
      tcgen05_fence_after();
      const uint32_t trow = tmem_base + as * TMEM_STAGE_COLS + ((uint32_t)(q * 32) << 16);
      for (int cb = grp; cb < n_blocks; cb += G) {
        const int nb0 = n0 + cb * 64;
        const int ncols = min(64, p.block_n - cb * 64);
        // the previous TMA store out of this group's staging tiles must have finished reading them
        if (gtid == 0 && !p.in_kind) tma_store_wait_read();
        named_bar_sync(1 + grp, 128);
        // accumulator block -> registers (four loads in flight, one wait)
        uint32_t acc[64];
        __syncwarp();
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
          if (c4 * 16 < ncols) tmem_ld16_nowait(trow + cb * 64 + c4 * 16, acc + c4 * 16);
        tmem_ld_wait();
        if (p.in_kind) {
          mbar_wait(&in_bar[grp], in_cnt & 1u);
          ++in_cnt;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j * 8 < ncols) {
            float vv[8];
            const int cl = cb * 64 + j * 8;  // column inside the tile
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] = __uint_as_float(acc[j * 8 + i]) + s_bias[cl + i];
            if (p.has_aux) pack_store_sw(st_b, r, j, vv);
            if (p.in_kind == 2) {
              const uint4 raw = *reinterpret_cast<const uint4*>(st_b + sw128_chunk(r, j));
              const bf16* u = reinterpret_cast<const bf16*>(&raw);
              if (g.act == LNX_ACT_GELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) vv[i] *= gelu_grad_fast(__bfloat162float(u[i]));
              } else if (g.act == LNX_ACT_RELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) vv[i] = __bfloat162float(u[i]) > 0.f ? vv[i] : 0.f;
              }
            } else if (g.act == LNX_ACT_GELU) {
#pragma unroll
              for (int i = 0; i < 8; ++i) vv[i] = gelu_fast(vv[i]);
            } else if (g.act == LNX_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 8; ++i) vv[i] = fmaxf(vv[i], 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] *= s_scale[cl + i] * rs;
            if (p.in_kind == 1) {
              const uint4 raw = *reinterpret_cast<const uint4*>(st_b + sw128_chunk(r, j));
              const bf16* u = reinterpret_cast<const bf16*>(&raw);
#pragma unroll
              for (int i = 0; i < 8; ++i) vv[i] += __bfloat162float(u[i]);
            }
            if (!row_ok) {
#pragma unroll
              for (int i = 0; i < 8; ++i) vv[i] = 0.f;  // keeps the fused column sums exact at the M edge
            }
            pack_store_sw(st_c, r, j, vv);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + grp, 128);
        if (colsum_out) {
          // column sums of the bf16 values just staged: thread = (column, half of the rows)
          const int col = gtid & 63, half = gtid >> 6;
          if (col < ncols && nb0 + col < p.N) {
            float a = 0.f;
            const int j = col >> 3, e2 = col & 7;
#pragma unroll 8
            for (int rr = half * 64; rr < half * 64 + 64; ++rr)
              a += __bfloat162float(*reinterpret_cast<const bf16*>(st_c + sw128_chunk(rr, j) + e2 * 2));
            atomicAdd(&s_colsum[nb0 + col], a);
          }
        }
        if (gtid == 0) {
          tma_store_2d(&tmC, st_c, nb0, m0);
          if (p.has_aux) tma_store_2d(&tmAux, st_b, nb0, m0);
          tma_store_commit();
          if (p.in_kind) {
            // every thread of the group has consumed st_b (barrier above): prefetch the next block's input tile
            int nt = t, ncb = cb + G;
            if (ncb >= n_blocks) { nt = t + gridDim.x; ncb = grp; }
            if (nt < num_tiles) {
              mbar_expect_tx(&in_bar[grp], STAGING_BYTES);
              tma_load_2d(st_b, &tmIn, &in_bar[grp], (nt % p.tiles_n) * p.block_n + ncb * 64, (nt / p.tiles_n) * BLOCK_M);
            }
            tma_store_wait_read();  // st_c is free again before the group passes the next block's first barrier
          }
        }
      }
      // this warp is done reading the accumulator
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
    if (gtid == 0) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (colsum_out)
    for (int i = threadIdx.x; i < p.N; i += nthreads) {
      const float v = s_colsum[i];
      if (v != 0.f) atomicAdd(colsum_out + i, v);
    }
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

bool tmap2d(CUtensorMap* tm, const void* ptr, long long inner, long long outer, long long ld, int box_outer) {
  const long long dims[2] = {inner, outer};
  const long long strides[1] = {ld};
  const int box[2] = {64, box_outer};
  return make_tmap(tm, ptr, 2, dims, strides, box);
}

int pick_block_n2(int N, bool b_trans) {
  if (b_trans) {
    if (N <= 64) return 64;
    if (N <= 128) return 128;
    if (N % 256 == 0 || N > 512) return 256;
    return (N % 192 == 0) ? 192 : 128;
  }
  if (N <= 256) return ((N + 15) / 16) * 16;  // single n-tile: the partial 64-column block is clipped at the tensor edge
  // several n-tiles: whole 64-column staging blocks only
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  return 256;
}

}  // namespace

// bf16 in / bf16 out, no split-K.  colsum_out (nullable, float[N], +=) receives the column sums of C.
int lnx_gemm_tc2(const GemmArgs& g, float* colsum_out, cudaStream_t st) {
  if (g.N % 8 != 0 || g.lda % 8 != 0 || g.ldb % 8 != 0 || g.accumulate) return LNX_ERR_UNSUPPORTED;
  if (!lnx_aligned16(g.A) || !lnx_aligned16(g.B) || !lnx_aligned16(g.C) || !lnx_aligned16(g.aux_out) || !lnx_aligned16(g.act_grad_in) ||
      !lnx_aligned16(g.residual))
    return LNX_ERR_UNSUPPORTED;
  if (colsum_out && g.N > 4096) return LNX_ERR_UNSUPPORTED;
  // buffer B of a group holds the aux output OR one element-wise input tile
  const int n_b_users = (g.aux_out ? 1 : 0) + (g.act_grad_in ? 1 : 0) + (g.residual ? 1 : 0);
  if (n_b_users > 1) return LNX_ERR_UNSUPPORTED;
  Tc2Params p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.a_trans = g.a_trans; p.b_trans = g.b_trans;
  p.block_n = pick_block_n2(g.N, g.b_trans != 0);
  p.tiles_m = (g.M + BLOCK_M - 1) / BLOCK_M;
  p.tiles_n = (g.N + p.block_n - 1) / p.block_n;
  p.has_aux = g.aux_out ? 1 : 0;
  p.in_kind = g.residual ? 1 : (g.act_grad_in ? 2 : 0);
  const int stage_bytes = A_STAGE_BYTES + p.block_n * BLOCK_K * 2;
  const int num_kb = (g.K + BLOCK_K - 1) / BLOCK_K;
  const int n_blocks = (p.block_n + 63) / 64;
  // small K: the epilogue is the whole cost -> 4 groups; large K: deeper operand ring, 2 groups
  p.groups = (num_kb <= 4 && n_blocks >= 3) ? 4 : 2;
  auto fixed_bytes = [&](int groups) { return 1024 + groups * 2 * STAGING_BYTES + 512 + 4 * 2 * 256 * 4 + (colsum_out ? g.N * 4 : 0); };
  int stages = (MAX_SMEM - fixed_bytes(p.groups)) / stage_bytes;
  if (stages < 2 && p.groups == 4) {
    p.groups = 2;
    stages = (MAX_SMEM - fixed_bytes(p.groups)) / stage_bytes;
  }
  if (stages < 2) return LNX_ERR_UNSUPPORTED;
  p.stages = min(min(stages, 6), max(2, 2 * num_kb));
  const size_t smem = (size_t)p.stages * stage_bytes + fixed_bytes(p.groups);

  CUtensorMap tmA, tmB, tmC, tmAux, tmIn;
  bool ok;
  if (!g.a_trans) ok = tmap2d(&tmA, g.A, g.K, g.M, g.lda, BLOCK_M);
  else ok = tmap2d(&tmA, g.A, g.M, g.K, g.lda, 64);
  if (!g.b_trans) ok = ok && tmap2d(&tmB, g.B, g.K, g.N, g.ldb, p.block_n);
  else ok = ok && tmap2d(&tmB, g.B, g.N, g.K, g.ldb, 64);
  ok = ok && tmap2d(&tmC, g.C, g.N, g.M, g.N, BLOCK_M);
  if (g.aux_out) ok = ok && tmap2d(&tmAux, g.aux_out, g.N, g.M, g.N, BLOCK_M);
  else tmAux = tmC;
  if (p.in_kind) ok = ok && tmap2d(&tmIn, p.in_kind == 1 ? g.residual : g.act_grad_in, g.N, g.M, g.N, BLOCK_M);
  else tmIn = tmC;
  if (!ok) return LNX_ERR_UNSUPPORTED;

  static int smem_set = 0;
  if ((int)smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    smem_set = (int)smem;
  }
  const int grid = min(p.tiles_m * p.tiles_n, kNumSMs);
  gemm_tc2_kernel<<<grid, 64 + 128 * p.groups, smem, st>>>(tmA, tmB, tmC, tmAux, tmIn, g, p, colsum_out);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
