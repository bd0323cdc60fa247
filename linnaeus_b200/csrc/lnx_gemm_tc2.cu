// Persistent, warp-specialised bf16 GEMM (tcgen05 + TMEM + TMA) with a TMA-fed, TMA-stored epilogue.
//
// One CTA per SM walks the output tiles (n fastest).  Roles (64 + 512 threads):
//   warp 0      TMA producer: A/B k-blocks into a 128B-swizzled smem ring (full/empty mbarriers)
//   warp 1      MMA issuer: tcgen05.mma cta_group::1, M = 128, N = block_n, accumulators double
//               buffered in TMEM (2 x 256 columns) so the next tile's MMAs overlap this tile's epilogue
//   warps 2-17  epilogue: 4 groups of 4 warps (a warp may only read its own TMEM lane quarter).  The unit of
//               work is one [32 rows x 64 columns] block; the 64-column blocks of the CTA's tile sequence are
//               dealt round-robin to the groups ACROSS tiles, so 3-block tiles (N = 192) still keep all four
//               groups busy.  Every warp is its own pipeline -- no cross-warp barrier in the epilogue: its
//               element-wise input block (residual or saved pre-activation) arrives by a per-warp TMA load
//               prefetched one block ahead, the accumulator is read 32 columns at a time with tcgen05.ld,
//               bias / GELU / act' / layer scale / DropPath mask / residual are applied in registers, results
//               go to the warp's private swizzled staging block and leave with one TMA store (full 128-byte
//               lines, clipped at the M/N edges by the tensor map).
// The epilogue is compiled per (activation, input kind, aux output, scaling, column sums) so each variant
// carries only its own instructions: with K = 96 the conv-stage GEMMs are pure epilogue/HBM work.
// Optional fused column sums of the output (bias gradients) are accumulated per CTA in shared memory across its
// tiles and flushed with one atomic per column.
#include <stdlib.h>

#include "lnx_gemm.cuh"
#include "lnx_tc_common.cuh"

using namespace lnx;
using namespace lnx_tc;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int MAX_GROUPS = 4;
constexpr int MAX_EPI_WARPS = 4 * MAX_GROUPS;
constexpr int MAX_THREADS = 64 + 32 * MAX_EPI_WARPS;
constexpr int WARP_STAGE_BYTES = 32 * 128;  // [32 rows][64 bf16]
constexpr int TMEM_STAGE_COLS = 256;
constexpr int MAX_SMEM = 232448;

struct Tc2Params {
  int M, N, K;
  int block_n, stages;
  int groups;  // epilogue groups of 4 warps: 4 when the epilogue is the whole cost (small K), 2 with a deeper operand ring
  int bufs;    // staging blocks per epilogue warp (2: aux output / element-wise input / double-buffered C)
  int a_trans, b_trans;
  int tiles_m, tiles_n;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void unpack8(const uint4& raw, float* u) {
  u[0] = __uint_as_float(raw.x << 16); u[1] = __uint_as_float(raw.x & 0xffff0000u);
  u[2] = __uint_as_float(raw.y << 16); u[3] = __uint_as_float(raw.y & 0xffff0000u);
  u[4] = __uint_as_float(raw.z << 16); u[5] = __uint_as_float(raw.z & 0xffff0000u);
  u[6] = __uint_as_float(raw.w << 16); u[7] = __uint_as_float(raw.w & 0xffff0000u);
}

// ACT: LNX_ACT_*.  IN_KIND: 0 none, 1 residual (added last), 2 saved pre-activation (act' multiplies the accumulator).
// AUX: second output = pre-activation (accumulator + bias).  SCALE: col_scale and/or row_scale.  COLSUM: column sums.
template <int ACT, int IN_KIND, bool AUX, bool SCALE, bool COLSUM>
__global__ void __launch_bounds__(MAX_THREADS, 1)
    gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                    const __grid_constant__ CUtensorMap tmAux, const __grid_constant__ CUtensorMap tmIn, const GemmArgs g, const Tc2Params p,
                    float* __restrict__ colsum_out) {
  static_assert(!(AUX && IN_KIND != 0), "the second staging block holds the aux output OR the element-wise input");
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int b_stage_bytes = p.block_n * BLOCK_K * 2;
  unsigned char* smem_a = base;
  unsigned char* smem_b = smem_a + (size_t)p.stages * A_STAGE_BYTES;
  unsigned char* staging = smem_b + (size_t)p.stages * b_stage_bytes;  // [epilogue warps][bufs][4 KB]
  const int EPI_WARPS = 4 * p.groups;
  const int GROUPS = p.groups;
  const int NUM_THREADS = blockDim.x;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + (size_t)EPI_WARPS * p.bufs * WARP_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;        // [2]
  uint64_t* in_bar = tempty_bar + 2;           // [16 warps][4 buffers]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_bar + MAX_EPI_WARPS * 4);
  float* s_colsum = reinterpret_cast<float*>(tmem_slot + 4);  // [N] when COLSUM

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int num_tiles = p.tiles_m * p.tiles_n;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmC);
    if (AUX) prefetch_tmap(&tmAux);
    if (IN_KIND) prefetch_tmap(&tmIn);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4 * p.groups);  // one arrival per epilogue warp
    }
    for (int s = 0; s < MAX_EPI_WARPS * 4; ++s) mbar_init(&in_bar[s], 1);
    mbar_fence_init();
  }
  if (COLSUM)
    for (int i = threadIdx.x; i < p.N; i += NUM_THREADS) s_colsum[i] = 0.f;
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
          mbar_wait_relaxed(&empty_bar[s], ph ^ 1u);
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
          mbar_wait_relaxed(&full_bar[s], ph);
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
    const int grp = e >> 2;                 // 0 .. GROUPS-1
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int my_tiles = ((int)blockIdx.x < num_tiles) ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if constexpr (IN_KIND != 0) {
      // ---- element-wise input variants: [32 rows x 32 columns] blocks (64B swizzle) computed IN PLACE in a ring of four
      // 2 KB buffers per warp: the input block arrives by TMA three blocks ahead of its use, the result overwrites it
      // and leaves with a TMA store; a buffer is refilled once the store issued one block earlier has read it.
      constexpr int NBUF = 4, BUF_BYTES = 32 * 64;
      unsigned char* stb = staging + (size_t)e * (2 * WARP_STAGE_BYTES);
      uint64_t* my_in = &in_bar[e * NBUF];
      const int n_blocks = (p.block_n + 31) / 32;
      const int total_blocks = my_tiles * n_blocks;
      auto block_coords = [&](int b, int& col0, int& row0) {
        const int tl = b / n_blocks, cb = b - tl * n_blocks;
        const int t = (int)blockIdx.x + tl * (int)gridDim.x;
        col0 = (t % p.tiles_n) * p.block_n + cb * 32;
        row0 = (t / p.tiles_n) * BLOCK_M + q * 32;
      };
      auto issue_in = [&](int b, int buf) {
        int c0, r0;
        block_coords(b, c0, r0);
        mbar_expect_tx(&my_in[buf], BUF_BYTES);
        tma_load_2d(stb + buf * BUF_BYTES, &tmIn, &my_in[buf], c0, r0);
      };
      if (lane == 0)
        for (int k = 0; k < NBUF - 1; ++k)
          if (grp + k * GROUPS < total_blocks) issue_in(grp + k * GROUPS, k);
      const uint32_t sw = (uint32_t)((lane >> 1) & 3);
      int b = grp, i = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int t = (int)blockIdx.x + tl * (int)gridDim.x;
        const int m0 = (t / p.tiles_n) * BLOCK_M, n0 = (t % p.tiles_n) * p.block_n;
        const uint32_t as = tl & 1u, aph = (tl >> 1) & 1u;
        const int m = m0 + q * 32 + lane;
        const bool row_ok = m < p.M;
        float rs = 1.f;
        if (SCALE) rs = (g.row_scale && row_ok) ? g.row_scale[m / g.rows_per_group] : 1.f;
        mbar_wait(&tfull_bar[as], aph);
        tcgen05_fence_after();
        const uint32_t trow = tmem_base + as * TMEM_STAGE_COLS + ((uint32_t)(q * 32) << 16);
        const int b_end = (tl + 1) * n_blocks;
        for (; b < b_end; b += GROUPS, ++i) {
          const int cb = b - tl * n_blocks;
          const int nb0 = n0 + cb * 32;
          const int ncols = min(32, p.block_n - cb * 32);
          const int buf = i & (NBUF - 1);
          unsigned char* st = stb + buf * BUF_BYTES;
          uint32_t acc[32];
          tmem_ld32_nowait(trow + cb * 32, acc);
          mbar_wait(&my_in[buf], (uint32_t)(i >> 2) & 1u);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = nb0 + j * 8;
            const uint32_t soff = (uint32_t)(lane * 64) + ((((uint32_t)j) ^ sw) << 4);
            float vv[8], u[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) vv[k] = __uint_as_float(acc[j * 8 + k]);
            if (g.bias && n < p.N) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(g.bias + n) + 1);
              vv[0] += b0.x; vv[1] += b0.y; vv[2] += b0.z; vv[3] += b0.w;
              vv[4] += b1.x; vv[5] += b1.y; vv[6] += b1.z; vv[7] += b1.w;
            }
            unpack8(*reinterpret_cast<const uint4*>(st + soff), u);
            if (IN_KIND == 2) {
              if (ACT == LNX_ACT_GELU) {
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                  const float2 r = __fmul2_rn(make_float2(vv[k], vv[k + 1]), gelu_grad_tanh3_x2(make_float2(u[k], u[k + 1])));
                  vv[k] = r.x;
                  vv[k + 1] = r.y;
                }
              } else if (ACT == LNX_ACT_RELU) {
#pragma unroll
                for (int k = 0; k < 8; ++k) vv[k] = u[k] > 0.f ? vv[k] : 0.f;
              } else if (ACT == LNX_ACT_MUL) {  // the saved tensor is the derivative itself
#pragma unroll
                for (int k = 0; k < 8; ++k) vv[k] *= u[k];
              }
            }
            if (SCALE) {
              if (g.col_scale && n < p.N) {
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(g.col_scale + n));
                const float4 s1 = __ldg(reinterpret_cast<const float4*>(g.col_scale + n) + 1);
                vv[0] *= s0.x * rs; vv[1] *= s0.y * rs; vv[2] *= s0.z * rs; vv[3] *= s0.w * rs;
                vv[4] *= s1.x * rs; vv[5] *= s1.y * rs; vv[6] *= s1.z * rs; vv[7] *= s1.w * rs;
              } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) vv[k] *= rs;
              }
            }
            if (IN_KIND == 1) {
#pragma unroll
              for (int k = 0; k < 8; ++k) vv[k] += u[k];
            }
            if (COLSUM && !row_ok) {
#pragma unroll
              for (int k = 0; k < 8; ++k) vv[k] = 0.f;  // keeps the fused column sums exact at the M edge
            }
            *reinterpret_cast<uint4*>(st + soff) = pack8(vv);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (COLSUM) {
            // column sums of the bf16 values just staged: lane = column, over this warp's 32 rows
            float cs = 0.f;
            const uint32_t cj = (uint32_t)(lane >> 3), ce = (uint32_t)(lane & 7) * 2;
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
              const unsigned short w = *reinterpret_cast<const unsigned short*>(st + rr * 64 + ((cj ^ (uint32_t)((rr >> 1) & 3)) << 4) + ce);
              cs += __uint_as_float((uint32_t)w << 16);
            }
            if (lane < ncols && nb0 + lane < p.N) atomicAdd(&s_colsum[nb0 + lane], cs);
            __syncwarp();
          }
          if (lane == 0) {
            const int row0 = m0 + q * 32;
            if (row0 < p.M) tma_store_2d(&tmC, st, nb0, row0);
            tma_store_commit();
            const int nb = b + (NBUF - 1) * GROUPS;
            if (nb < total_blocks) {
              tma_store_wait_read<1>();  // the store issued one block ago has read buffer (i - 1) & 3 = (i + 3) & 3
              issue_in(nb, (i + NBUF - 1) & (NBUF - 1));
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
      }
      if (lane == 0) tma_store_wait_all();
    } else {
    unsigned char* st0 = staging + (size_t)e * p.bufs * WARP_STAGE_BYTES;
    unsigned char* st1 = st0 + WARP_STAGE_BYTES;  // aux output / element-wise input / second C buffer (bufs == 2)
    const bool dbl = !AUX && !IN_KIND && p.bufs == 2;
    uint64_t* my_in = &in_bar[e];
    const int n_blocks = (p.block_n + 63) / 64;
    const int total_blocks = my_tiles * n_blocks;
    uint32_t in_cnt = 0, cbuf = 0;
    const int r_sw = lane & 7;

    // coordinates of this warp's global block index b (deal order: tile-major, then column block)
    auto block_coords = [&](int b, int& col0, int& row0) {
      const int tl = b / n_blocks, cb = b - tl * n_blocks;
      const int t = (int)blockIdx.x + tl * (int)gridDim.x;
      col0 = (t % p.tiles_n) * p.block_n + cb * 64;
      row0 = (t / p.tiles_n) * BLOCK_M + q * 32;
    };
    if (IN_KIND && lane == 0 && grp < total_blocks) {
      int c0, r0;
      block_coords(grp, c0, r0);
      mbar_expect_tx(my_in, WARP_STAGE_BYTES);
      tma_load_2d(st1, &tmIn, my_in, c0, r0);
    }

    int b = grp;  // next global block of this group
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int t = (int)blockIdx.x + tl * (int)gridDim.x;
      const int m0 = (t / p.tiles_n) * BLOCK_M, n0 = (t % p.tiles_n) * p.block_n;
      const uint32_t as = tl & 1u, aph = (tl >> 1) & 1u;
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < p.M;
      float rs = 1.f;
      if (SCALE) rs = (g.row_scale && row_ok) ? g.row_scale[m / g.rows_per_group] : 1.f;
      int tok_pos = -1;  // image-token index of this row within its sequence (token-position factors)
      if (SCALE && g.tok_scale && row_ok) tok_pos = m % g.tok_period - g.tok_extra;
      const float tok_tx = (float)(tok_pos >= 0 ? tok_pos % g.tok_w : 0), tok_ty = (float)(tok_pos >= 0 ? tok_pos / g.tok_w : 0);
      mbar_wait(&tfull_bar[as], aph);
      tcgen05_fence_after();
      const uint32_t trow = tmem_base + as * TMEM_STAGE_COLS + ((uint32_t)(q * 32) << 16);
      const int b_end = (tl + 1) * n_blocks;
      for (; b < b_end; b += GROUPS) {
        const int cb = b - tl * n_blocks;
        const int nb0 = n0 + cb * 64;
        const int ncols = min(64, p.block_n - cb * 64);
        unsigned char* st_c = (dbl && cbuf) ? st1 : st0;
        // the TMA store that last read this staging block must be done with it
        if (lane == 0) {
          if (dbl) tma_store_wait_read<1>();
          else tma_store_wait_read<0>();
        }
        __syncwarp();
        if (IN_KIND) {
          mbar_wait(my_in, in_cnt & 1u);
          ++in_cnt;
        }
        // token-position factors of this 64-column block (32 pairs): cos(tx fx + ty fy) straight from the learnable frequencies --
        // the two frequency rows are warp-uniform (broadcast loads, 1.5 KB per layer) and the cosine is one MUFU op, so nothing
        // per-element is read from memory (a [pair][position] table cost +34 us per projection, more than the pass it replaced)
        float tokf[32];
        const bool tok_on = SCALE && g.tok_scale && nb0 < 2 * g.tok_dim;
        if (tok_on) {
          const float qs = nb0 < g.tok_dim ? g.tok_qscale : 1.f;
          const float* fx = g.tok_scale + ((nb0 < g.tok_dim ? nb0 : nb0 - g.tok_dim) >> 1);
          const float* fy = fx + (g.tok_dim >> 1);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(fx + i)), c = __ldg(reinterpret_cast<const float4*>(fy + i));
            tokf[i] = tok_pos >= 0 ? __cosf(fmaf(tok_tx, a.x, tok_ty * c.x)) * qs : qs;
            tokf[i + 1] = tok_pos >= 0 ? __cosf(fmaf(tok_tx, a.y, tok_ty * c.y)) * qs : qs;
            tokf[i + 2] = tok_pos >= 0 ? __cosf(fmaf(tok_tx, a.z, tok_ty * c.z)) * qs : qs;
            tokf[i + 3] = tok_pos >= 0 ? __cosf(fmaf(tok_tx, a.w, tok_ty * c.w)) * qs : qs;
          }
        }
        float csum[2] = {0.f, 0.f};
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half * 32 < ncols) {
            uint32_t acc[32];
            tmem_ld32_nowait(trow + cb * 64 + half * 32, acc);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int jg = half * 4 + j;
              const int n = nb0 + jg * 8;
              const uint32_t soff = (uint32_t)(lane * 128 + ((jg ^ r_sw) << 4));
              float vv[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) vv[i] = __uint_as_float(acc[j * 8 + i]);
              if (g.bias && n < p.N) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(g.bias + n) + 1);
                vv[0] += b0.x; vv[1] += b0.y; vv[2] += b0.z; vv[3] += b0.w;
                vv[4] += b1.x; vv[5] += b1.y; vv[6] += b1.z; vv[7] += b1.w;
              }
              if (AUX && ACT == LNX_ACT_GELU_DG) {  // out = gelu(pre), aux = gelu'(pre): one tanh for both
                float dg[8];
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                  float2 gl, d;
                  gelu_both_tanh3_x2(make_float2(vv[i], vv[i + 1]), gl, d);
                  vv[i] = gl.x; vv[i + 1] = gl.y;
                  dg[i] = d.x; dg[i + 1] = d.y;
                }
                *reinterpret_cast<uint4*>(st1 + soff) = pack8(dg);
              } else if (AUX) {
                *reinterpret_cast<uint4*>(st1 + soff) = pack8(vv);
              }
              if (IN_KIND == 2) {
                float u[8];
                unpack8(*reinterpret_cast<const uint4*>(st1 + soff), u);
                if (ACT == LNX_ACT_GELU) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) vv[i] *= gelu_grad_tanh3(u[i]);
                } else if (ACT == LNX_ACT_RELU) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) vv[i] = u[i] > 0.f ? vv[i] : 0.f;
                }
              } else if (ACT == LNX_ACT_GELU) {
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                  const float2 r = gelu_tanh3_x2(make_float2(vv[i], vv[i + 1]));
                  vv[i] = r.x;
                  vv[i + 1] = r.y;
                }
              } else if (ACT == LNX_ACT_RELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) vv[i] = fmaxf(vv[i], 0.f);
              } else if (ACT == LNX_ACT_SWISH) {
#pragma unroll
                for (int i = 0; i < 8; ++i) vv[i] = swish_fast(vv[i]);
              }
              if (SCALE) {
                if (g.tok_scale) {
                  if (tok_on) {  // q / k columns: cos factor of the token position (pairs share one), q also the softmax scale
#pragma unroll
                    for (int i = 0; i < 8; ++i) vv[i] *= tokf[jg * 4 + (i >> 1)];
                  }
                } else if (g.col_scale && n < p.N) {
                  const float4 s0 = __ldg(reinterpret_cast<const float4*>(g.col_scale + n));
                  const float4 s1 = __ldg(reinterpret_cast<const float4*>(g.col_scale + n) + 1);
                  vv[0] *= s0.x * rs; vv[1] *= s0.y * rs; vv[2] *= s0.z * rs; vv[3] *= s0.w * rs;
                  vv[4] *= s1.x * rs; vv[5] *= s1.y * rs; vv[6] *= s1.z * rs; vv[7] *= s1.w * rs;
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) vv[i] *= rs;
                }
              }
              if (IN_KIND == 1) {
                float u[8];
                unpack8(*reinterpret_cast<const uint4*>(st1 + soff), u);
#pragma unroll
                for (int i = 0; i < 8; ++i) vv[i] += u[i];
              }
              if (COLSUM && !row_ok) {
#pragma unroll
                for (int i = 0; i < 8; ++i) vv[i] = 0.f;  // keeps the fused column sums exact at the M edge
              }
              *reinterpret_cast<uint4*>(st_c + soff) = pack8(vv);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (COLSUM) {
          // column sums of the bf16 values just staged: lane = column pair (2 lane, 2 lane + 1) over this warp's 32 rows
          const int jg = lane >> 2;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(st_c + rr * 128 + ((jg ^ (rr & 7)) << 4) + (lane & 3) * 4);
            csum[0] += __uint_as_float(w << 16);
            csum[1] += __uint_as_float(w & 0xffff0000u);
          }
          const int c = nb0 + 2 * lane;
          if (2 * lane < ncols && c < p.N) {
            atomicAdd(&s_colsum[c], csum[0]);
            atomicAdd(&s_colsum[c + 1], csum[1]);
          }
          __syncwarp();
        }
        if (lane == 0) {
          const int row0 = m0 + q * 32;
          if (row0 < p.M) {
            tma_store_2d(&tmC, st_c, nb0, row0);
            if (AUX) tma_store_2d(&tmAux, st1, nb0, row0);
          }
          tma_store_commit();
          if (IN_KIND) {
            // every lane has consumed st1 (syncwarp above): prefetch this warp's next input block
            const int nb = b + GROUPS;
            if (nb < total_blocks) {
              int c0, r0;
              block_coords(nb, c0, r0);
              mbar_expect_tx(my_in, WARP_STAGE_BYTES);
              tma_load_2d(st1, &tmIn, my_in, c0, r0);
            }
          }
        }
        cbuf ^= 1u;
      }
      // this warp is done reading the accumulator of tile tl
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
    if (lane == 0) tma_store_wait_all();
    }  // generic (no element-wise input) path
  }

  tcgen05_fence_before();
  __syncthreads();
  if (COLSUM)
    for (int i = threadIdx.x; i < p.N; i += NUM_THREADS) {
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
// [32 rows][32 columns] blocks, 64-byte swizzle (element-wise input variants)
bool tmap2d_32(CUtensorMap* tm, const void* ptr, long long inner, long long outer, long long ld) {
  const long long dims[2] = {inner, outer};
  const long long strides[1] = {ld};
  const int box[2] = {32, 32};
  return make_tmap(tm, ptr, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
}

int pick_block_n2(int N, bool b_trans) {
  static const int forced = getenv("LNX_GEMM_BN") ? atoi(getenv("LNX_GEMM_BN")) : 0;  // experiments
  if (forced && N > 256 && N % forced == 0) return forced;
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

struct Launch {
  CUtensorMap tmA, tmB, tmC, tmAux, tmIn;
  GemmArgs g;
  Tc2Params p;
  float* colsum_out;
  int grid;
  size_t smem;
  cudaStream_t st;
};

template <int ACT, int IN_KIND, bool AUX, bool SCALE, bool COLSUM>
int launch_variant(const Launch& L) {
  auto kern = gemm_tc2_kernel<ACT, IN_KIND, AUX, SCALE, COLSUM>;
  static int smem_set = 0;
  if ((int)L.smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    smem_set = (int)L.smem;
  }
  kern<<<L.grid, 64 + 128 * L.p.groups, L.smem, L.st>>>(L.tmA, L.tmB, L.tmC, L.tmAux, L.tmIn, L.g, L.p, L.colsum_out);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

template <int ACT>
int launch_act(const Launch& L, int in_kind, bool aux, bool scale, bool colsum) {
  if (in_kind == 2) {
    if (aux || scale) return LNX_ERR_UNSUPPORTED;
    return colsum ? launch_variant<ACT, 2, false, false, true>(L) : launch_variant<ACT, 2, false, false, false>(L);
  }
  if (colsum || scale || in_kind) return LNX_ERR_UNSUPPORTED;
  return aux ? launch_variant<ACT, 0, true, false, false>(L) : launch_variant<ACT, 0, false, false, false>(L);
}

int launch_none(const Launch& L, int in_kind, bool aux, bool scale, bool colsum) {
  if (in_kind == 2 || aux) return LNX_ERR_UNSUPPORTED;  // act' of "none" is the identity: callers never pass it
  if (colsum) {
    if (in_kind || scale) return LNX_ERR_UNSUPPORTED;
    return launch_variant<LNX_ACT_NONE, 0, false, false, true>(L);
  }
  if (in_kind == 1) return scale ? launch_variant<LNX_ACT_NONE, 1, false, true, false>(L) : launch_variant<LNX_ACT_NONE, 1, false, false, false>(L);
  return scale ? launch_variant<LNX_ACT_NONE, 0, false, true, false>(L) : launch_variant<LNX_ACT_NONE, 0, false, false, false>(L);
}

}  // namespace

// bf16 in / bf16 out, no split-K.  colsum_out (nullable, float[N], +=) receives the column sums of C.
int lnx_gemm_tc2(const GemmArgs& g, float* colsum_out, cudaStream_t st) {
  if (g.N % 8 != 0 || g.lda % 8 != 0 || g.ldb % 8 != 0 || g.accumulate) return LNX_ERR_UNSUPPORTED;
  if (!lnx_aligned16(g.A) || !lnx_aligned16(g.B) || !lnx_aligned16(g.C) || !lnx_aligned16(g.aux_out) || !lnx_aligned16(g.act_grad_in) ||
      !lnx_aligned16(g.residual) || !lnx_aligned16(g.bias) || !lnx_aligned16(g.col_scale))
    return LNX_ERR_UNSUPPORTED;
  if (colsum_out && g.N > 4096) return LNX_ERR_UNSUPPORTED;
  // the second staging block of a warp holds the aux output OR one element-wise input block
  const int n_b_users = (g.aux_out ? 1 : 0) + (g.act_grad_in ? 1 : 0) + (g.residual ? 1 : 0);
  if (n_b_users > 1) return LNX_ERR_UNSUPPORTED;
  Launch L;
  Tc2Params& p = L.p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.a_trans = g.a_trans; p.b_trans = g.b_trans;
  p.block_n = pick_block_n2(g.N, g.b_trans != 0);
  p.tiles_m = (g.M + BLOCK_M - 1) / BLOCK_M;
  p.tiles_n = (g.N + p.block_n - 1) / p.block_n;
  const int in_kind = g.residual ? 1 : (g.act_grad_in ? 2 : 0);
  const int stage_bytes = A_STAGE_BYTES + p.block_n * BLOCK_K * 2;
  const int num_kb = (g.K + BLOCK_K - 1) / BLOCK_K;
  // small K: the epilogue is the whole cost -> 4 groups, double staging; large K: deeper operand ring first
  auto fixed_bytes = [&](int groups, int bufs) { return 1024 + 4 * groups * bufs * WARP_STAGE_BYTES + 1024 + (colsum_out ? g.N * 4 : 0); };
  const int want = min(4, 2 * num_kb);
  // the token-factor epilogue (cos-RoPE qkv projection) is heavy enough to want 16 epilogue warps at K = 384 too: 0.074 -> 0.068 ms
  p.groups = (num_kb <= 4 || (g.tok_scale && num_kb <= 12)) ? 4 : 2;
  p.bufs = 2;
  int stages = (MAX_SMEM - fixed_bytes(p.groups, p.bufs)) / stage_bytes;
  if (stages < want && n_b_users == 0) {
    p.bufs = 1;
    stages = (MAX_SMEM - fixed_bytes(p.groups, p.bufs)) / stage_bytes;
  }
  if (stages < want && p.groups == 4) {
    p.groups = 2;
    p.bufs = 2;
    stages = (MAX_SMEM - fixed_bytes(p.groups, p.bufs)) / stage_bytes;
    if (stages < want && n_b_users == 0) {
      p.bufs = 1;
      stages = (MAX_SMEM - fixed_bytes(p.groups, p.bufs)) / stage_bytes;
    }
  }
  if (stages < 2) return LNX_ERR_UNSUPPORTED;
  p.stages = min(min(stages, 8), max(2, 2 * num_kb));
  L.smem = (size_t)p.stages * stage_bytes + fixed_bytes(p.groups, p.bufs);

  bool ok;
  if (!g.a_trans) ok = tmap2d(&L.tmA, g.A, g.K, g.M, g.lda, BLOCK_M);
  else ok = tmap2d(&L.tmA, g.A, g.M, g.K, g.lda, 64);
  if (!g.b_trans) ok = ok && tmap2d(&L.tmB, g.B, g.K, g.N, g.ldb, p.block_n);
  else ok = ok && tmap2d(&L.tmB, g.B, g.N, g.K, g.ldb, 64);
  if (in_kind) {
    ok = ok && tmap2d_32(&L.tmC, g.C, g.N, g.M, g.N);
    ok = ok && tmap2d_32(&L.tmIn, in_kind == 1 ? g.residual : g.act_grad_in, g.N, g.M, g.N);
    L.tmAux = L.tmC;
  } else {
    ok = ok && tmap2d(&L.tmC, g.C, g.N, g.M, g.N, 32);
    if (g.aux_out) ok = ok && tmap2d(&L.tmAux, g.aux_out, g.N, g.M, g.N, 32);
    else L.tmAux = L.tmC;
    L.tmIn = L.tmC;
  }
  if (!ok) return LNX_ERR_UNSUPPORTED;

  L.g = g;
  L.colsum_out = colsum_out;
  L.grid = min(p.tiles_m * p.tiles_n, kNumSMs);
  L.st = st;
  const bool aux = g.aux_out != nullptr, scale = g.col_scale || g.row_scale || g.tok_scale, colsum = colsum_out != nullptr;
  if (g.tok_scale && (in_kind || aux || colsum || g.col_scale || g.row_scale || g.act != LNX_ACT_NONE || g.tok_dim % 64 != 0 || g.tok_w <= 0 || !lnx_aligned16(g.tok_scale)))
    return LNX_ERR_UNSUPPORTED;
  switch (g.act) {
    case LNX_ACT_NONE: return launch_none(L, in_kind, aux, scale, colsum);
    case LNX_ACT_GELU: return launch_act<LNX_ACT_GELU>(L, in_kind, aux, scale, colsum);
    case LNX_ACT_RELU: return launch_act<LNX_ACT_RELU>(L, in_kind, aux, scale, colsum);
    case LNX_ACT_SWISH:
      if (aux || in_kind || scale || colsum) return LNX_ERR_UNSUPPORTED;
      return launch_variant<LNX_ACT_SWISH, 0, false, false, false>(L);
    case LNX_ACT_GELU_DG:
      if (!aux || in_kind || scale || colsum) return LNX_ERR_UNSUPPORTED;
      return launch_variant<LNX_ACT_GELU_DG, 0, true, false, false>(L);
    case LNX_ACT_MUL:
      if (in_kind != 2 || aux || scale) return LNX_ERR_UNSUPPORTED;
      return colsum ? launch_variant<LNX_ACT_MUL, 2, false, false, true>(L) : launch_variant<LNX_ACT_MUL, 2, false, false, false>(L);
  }
  return LNX_ERR_UNSUPPORTED;
}
