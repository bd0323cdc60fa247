// bf16 GEMM on CTA pairs: C[M, N] = A[M, K] B[N, K]^T (+ bias), tcgen05.mma.cta_group::2.
//
// A CTA pair computes a 256 x BN tile: each CTA stages its own 128 rows of A and only HALF of the B tile; the pair's tensor cores read
// the other half from the peer's shared memory (per SM and k-step 4 + 4 KB of fills instead of 4 + 8).  Measured on the
// transformer-stage shapes (profiles/r02_microbench.md): same 1.16 PFLOP/s as the 1-CTA kernel where the main loop dominates
// (K = 3072), slower at K = 384 with this file's plain epilogue -- so the train step does not use it; it is the tested 2-SM building
// block (cluster launch, 2-SM TMA, multicast commit, pair-allocated TMEM) for cluster-multicast / stream-K work.
//
// Structure (persistent, one cluster of 2 CTAs per SM pair; rank 0 = leader):
//   warp 0  TMA producer (both CTAs): A rows [m0 + 128 rank, +128), B rows [n0 + BN/2 rank, +BN/2) of every k-block into the CTA's own
//           ring; the bytes of BOTH CTAs complete on the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 1  MMA issuer (leader only): tcgen05.mma.cta_group::2, M = 256, N = BN; tcgen05.commit ... multicast::cluster releases the
//           ring stage / publishes the accumulator in both CTAs
//   warps 2-5  epilogue (both CTAs): each CTA owns the 128 accumulator rows in its own tensor memory; TMEM -> registers -> bias ->
//           bf16 -> global; "accumulator drained" arrives on the leader's barrier from all 8 epilogue warps of the pair
// TMEM: two accumulator stages of BN columns (allocated with cta_group::2), so the next tile's main loop overlaps this epilogue.
#include <stdio.h>
#include <stdlib.h>

#include "lnx_gemm.cuh"
#include "lnx_tc_common.cuh"

using namespace lnx;
using namespace lnx_tc;

namespace {

constexpr int BLOCK_M = 128;  // rows per CTA (256 per pair)
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 192;
constexpr int MAX_SMEM = 232448;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: the even (leader) CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// executed by both CTAs; the transaction bytes land on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once all previously issued MMAs are complete) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((unsigned short)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair512(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair512(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

struct PairParams {
  int M, N, K;
  int block_n;  // 128, 192 or 256: columns of the pair's tile (each CTA stages block_n / 2 rows of B)
  int stages;
  int tiles_m, tiles_n;  // tiles_m counts 256-row pair tiles
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
    gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const float* __restrict__ bias,
                     bf16* __restrict__ C, const PairParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = BLOCK_M * BLOCK_K * 2;             // 16 KB
  const int b_bytes = (p.block_n / 2) * BLOCK_K * 2;      // this CTA's half of the B tile
  unsigned char* smem_a = base;
  unsigned char* smem_b = smem_a + (size_t)p.stages * a_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * b_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;  // [2]
  uint64_t* tempty = tfull + 2;        // [2] (used on the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8);  // 4 epilogue warps of each CTA of the pair
    }
    mbar_fence_init();
  }
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything is signalled across the pair
  if (warp == 1) tmem_alloc_pair512(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = pair; t < num_tiles; t += npairs) {
        const int m0 = (t / p.tiles_n) * (2 * BLOCK_M) + (int)rank * BLOCK_M;
        const int n0 = (t % p.tiles_n) * p.block_n + (int)rank * (p.block_n / 2);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % p.stages;
          mbar_wait_relaxed(&empty[s], ((it / p.stages) & 1u) ^ 1u);
          if (rank == 0) mbar_expect_tx(&full[s], 2u * (uint32_t)(a_bytes + b_bytes));  // both CTAs' bytes land here
          unsigned char* sa = smem_a + (size_t)s * a_bytes;
          unsigned char* sb = smem_b + (size_t)s * b_bytes;
          tma_load_2d_pair(sa, &tmA, &full[s], kb * BLOCK_K, m0);
          tma_load_2d_pair(sb, &tmB, &full[s], kb * BLOCK_K, n0);  // box = [block_n / 2 rows][64 k]
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, p.block_n, 0, 0);
      uint32_t it = 0, tl = 0;
      for (int t = pair; t < num_tiles; t += npairs, ++tl) {
        const uint32_t as = tl & 1u;
        mbar_wait(&tempty[as], ((tl >> 1) & 1u) ^ 1u);  // both CTAs have drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + as * 256;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % p.stages;
          mbar_wait(&full[s], (it / p.stages) & 1u);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem_a + (size_t)s * a_bytes);
          const uint32_t sb = smem_u32(smem_b + (size_t)s * b_bytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16_pair(tacc, make_smem_desc(sa + k * 32, 0, 1024), make_smem_desc(sb + k * 32, 0, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(&empty[s]);  // frees the stage in both CTAs
        }
        umma_commit_pair(&tfull[as]);
      }
    }
  } else {
    // ===================== epilogue (both CTAs): own 128 rows =====================
    const int q = warp & 3;
    uint32_t tl = 0;
    for (int t = pair; t < num_tiles; t += npairs, ++tl) {
      const uint32_t as = tl & 1u;
      const int m = (t / p.tiles_n) * (2 * BLOCK_M) + (int)rank * BLOCK_M + q * 32 + lane;
      const int n0 = (t % p.tiles_n) * p.block_n;
      mbar_wait(&tfull[as], (tl >> 1) & 1u);
      tcgen05_fence_after();
      const uint32_t trow = tmem_base + as * 256 + ((uint32_t)(q * 32) << 16);
      for (int c = 0; c < p.block_n; c += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(trow + c, v);
        const int n = n0 + c;
        if (m < p.M && n < p.N) {
          bf16* dst = C + (long long)m * p.N + n;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[j * 8 + i]);
            if (bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n + j * 8));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n + j * 8) + 1);
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            if (n + j * 8 < p.N)
              *reinterpret_cast<uint4*>(dst + j * 8) = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty[as]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA frees tensor memory (or exits) while the pair's MMAs / remote arrivals may still touch it
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_pair512(tmem_base);
  }
}

bool tmap_k(CUtensorMap* tm, const void* ptr, long long inner, long long outer, long long ld, int box_rows) {
  const long long dims[2] = {inner, outer};
  const long long strides[1] = {ld};
  const int box[2] = {BLOCK_K, box_rows};
  return make_tmap(tm, ptr, 2, dims, strides, box);
}

}  // namespace

// a [M, K], b [N, K] bf16 row-major (K % 8 == 0), c [M, N] bf16, bias float [N] (nullable).  N % 64 == 0.
extern "C" int lnx_gemm_pair(const void* a, const void* b, const float* bias, void* c, int64_t M, int N, int K, lnx_stream_t s) {
  LNX_REQUIRE(a && b && c, LNX_ERR_NULL);
  LNX_REQUIRE(M > 0 && M < (1ll << 31) && N > 0 && K > 0, LNX_ERR_SHAPE);
  if (N % 128 != 0 || K % 8 != 0) return LNX_ERR_UNSUPPORTED;
  if (!lnx_aligned16(a) || !lnx_aligned16(b) || !lnx_aligned16(c) || !lnx_aligned16(bias)) return LNX_ERR_ALIGN;
  PairParams p;
  p.M = (int)M; p.N = N; p.K = K;
  static const int forced = getenv("LNX_PAIR_BN") ? atoi(getenv("LNX_PAIR_BN")) : 0;
  p.block_n = (N % 256 == 0) ? 256 : 128;
  if (forced && N % forced == 0) p.block_n = forced;
  p.tiles_m = (int)((M + 2 * BLOCK_M - 1) / (2 * BLOCK_M));
  p.tiles_n = N / p.block_n;
  const int stage = BLOCK_M * BLOCK_K * 2 + (p.block_n / 2) * BLOCK_K * 2;
  p.stages = min(8, (MAX_SMEM - 2048) / stage);
  const size_t smem = (size_t)p.stages * stage + 2048;
  CUtensorMap tmA, tmB;
  if (!tmap_k(&tmA, a, K, M, K, BLOCK_M) || !tmap_k(&tmB, b, K, N, K, p.block_n / 2)) return LNX_ERR_UNSUPPORTED;
  static int smem_set = 0;
  if ((int)smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    smem_set = (int)smem;
  }
  const int grid = 2 * min(p.tiles_m * p.tiles_n, kNumSMs / 2);
  gemm_pair_kernel<<<grid, NTHREADS, smem, (cudaStream_t)s>>>(tmA, tmB, bias, (bf16*)c, p);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
