// bf16 GEMM on the 5th-generation tensor cores: TMA -> 128B-swizzled shared memory ->
// tcgen05.mma (cta_group::1, M = 128, N = block_n <= 256, K = 16 per instruction) with
// the fp32 accumulator in TMEM, read back with tcgen05.ld for the fused epilogue.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA
// issuer (one elected lane), warps 2-5 = epilogue (TMEM lane quarter = warp_id % 4).
// Operands may be K-major (stored [rows, K]) or MN-major (stored [K, rows]); the
// latter serves the data- and weight-gradient GEMMs without any transposed copies.
// Split-K (grid.z) + fp32 atomics serves the weight gradients (reduction over tokens).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "lnx_common.cuh"
#include "lnx_gemm.cuh"

using namespace lnx;

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;             // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int NUM_THREADS = 192;
constexpr uint32_t SPIN_LIMIT = 1u << 24;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  const uint32_t addr = smem_u32(bar);
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > SPIN_LIMIT) __trap();  // never hang the GPU on a pipeline bug
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor (sm_100): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | SWIZZLE_128B (2) <<61
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct TcParams {
  int M, N, K;
  int block_n, stages, tmem_cols;
  int a_trans, b_trans;
  int kb_per_split;
};

template <typename TC>
__device__ __forceinline__ void load8_as_f32(const TC* p, float* v);
template <>
__device__ __forceinline__ void load8_as_f32<float>(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8_as_f32<bf16>(const bf16* p, float* v) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const bf16* h = reinterpret_cast<const bf16*>(&raw);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(h[i]);
}
template <typename TC>
__device__ __forceinline__ void store8_from_f32(TC* p, const float* v);
template <>
__device__ __forceinline__ void store8_from_f32<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8_from_f32<bf16>(bf16* p, const float* v) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = raw;
}

template <typename TC>
__global__ void __launch_bounds__(NUM_THREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                               const GemmArgs g, const TcParams p) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte aligned operand ring (SWIZZLE_128B atoms are 1024 bytes)
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int b_stage_bytes = p.block_n * BLOCK_K * 2;
  unsigned char* smem_a = base;
  unsigned char* smem_b = base + (size_t)p.stages * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * b_stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BLOCK_M, n0 = blockIdx.x * p.block_n;
  const int num_kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
  const int num_kb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // allocate TMEM columns (power of two >= 32); whole warp, then give up the permit
    const uint32_t dst = smem_u32(tmem_slot);
    if (p.tmem_cols == 32) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(dst) : "memory");
    else if (p.tmem_cols == 64) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(dst) : "memory");
    else if (p.tmem_cols == 128) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(dst) : "memory");
    else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(dst) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && num_kb > 0) {
      const uint32_t stage_bytes = A_STAGE_BYTES + (uint32_t)b_stage_bytes;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % p.stages;
        const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], stage_bytes);
        const int k0 = (kb_begin + i) * BLOCK_K;
        unsigned char* sa = smem_a + (size_t)s * A_STAGE_BYTES;
        unsigned char* sb = smem_b + (size_t)s * b_stage_bytes;
        if (!p.a_trans) {
          tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);  // box {64 k, 128 m}
        } else {
          tma_load_2d(sa, &tmA, &full_bar[s], m0, k0);  // 2 boxes {64 m, 64 k}
          tma_load_2d(sa + 8192, &tmA, &full_bar[s], m0 + 64, k0);
        }
        if (!p.b_trans) {
          tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);  // box {64 k, block_n}
        } else {
          for (int j = 0; j < p.block_n / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], n0 + 64 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && num_kb > 0) {
      // instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.a_trans ? 1 : 0) << 15) |
                             ((uint32_t)(p.b_trans ? 1 : 0) << 16) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % p.stages;
        const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem_a + (size_t)s * A_STAGE_BYTES);
        const uint32_t sb = smem_u32(smem_b + (size_t)s * b_stage_bytes);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // K-major: +32 bytes per UMMA_K inside the swizzle row, SBO = 8 rows * 128 B
          // MN-major: +16 k-rows * 128 B per UMMA_K, LBO = next 64-wide MN block (8 KB), SBO = 8 k-rows * 128 B
          const uint64_t ad = p.a_trans ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 0, 1024);
          const uint64_t bd = p.b_trans ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 0, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int m = m0 + row;
    if (num_kb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tcgen05_fence_after();
    }
    TC* C = reinterpret_cast<TC*>(g.C);
    TC* aux = reinterpret_cast<TC*>(g.aux_out);
    const TC* agi = reinterpret_cast<const TC*>(g.act_grad_in);
    const TC* res = reinterpret_cast<const TC*>(g.residual);
    for (int c = 0; c < p.block_n; c += 16) {
      float v[16];
      __syncwarp();  // tcgen05.ld is .sync.aligned: the whole warp must be converged here
      if (num_kb > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = n0 + c + h * 8;
        if (m >= p.M || n >= p.N) continue;  // N % 8 == 0 is guaranteed by the host
        float* vv = v + h * 8;
        const long long idx = (long long)m * p.N + n;
        if (g.accumulate) {
          float* dst = reinterpret_cast<float*>(g.C) + idx;
          atomicAdd(reinterpret_cast<float4*>(dst), make_float4(vv[0], vv[1], vv[2], vv[3]));
          atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(vv[4], vv[5], vv[6], vv[7]));
          continue;
        }
        if (g.bias) {
          const float4 b0 = *reinterpret_cast<const float4*>(g.bias + n), b1 = *reinterpret_cast<const float4*>(g.bias + n + 4);
          vv[0] += b0.x; vv[1] += b0.y; vv[2] += b0.z; vv[3] += b0.w;
          vv[4] += b1.x; vv[5] += b1.y; vv[6] += b1.z; vv[7] += b1.w;
        }
        if (aux) {
          if (g.act == LNX_ACT_GELU_DG) {
            float dg[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) dg[i] = gelu_grad_fast(vv[i]);
            store8_from_f32<TC>(aux + idx, dg);
          } else {
            store8_from_f32<TC>(aux + idx, vv);
          }
        }
        if (agi) {
          float u[8];
          load8_as_f32<TC>(agi + idx, u);
          if (g.act == LNX_ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] *= gelu_grad_fast(u[i]);
          } else if (g.act == LNX_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] = u[i] > 0.f ? vv[i] : 0.f;
          } else if (g.act == LNX_ACT_MUL) {
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] *= u[i];
          }
        } else if (g.act == LNX_ACT_GELU || g.act == LNX_ACT_GELU_DG) {
#pragma unroll
          for (int i = 0; i < 8; ++i) vv[i] = gelu_fast(vv[i]);
        } else if (g.act == LNX_ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 8; ++i) vv[i] = fmaxf(vv[i], 0.f);
        } else if (g.act == LNX_ACT_SWISH) {
#pragma unroll
          for (int i = 0; i < 8; ++i) vv[i] = swish_fast(vv[i]);
        }
        if (g.col_scale) {
          const float4 s0 = *reinterpret_cast<const float4*>(g.col_scale + n), s1 = *reinterpret_cast<const float4*>(g.col_scale + n + 4);
          vv[0] *= s0.x; vv[1] *= s0.y; vv[2] *= s0.z; vv[3] *= s0.w;
          vv[4] *= s1.x; vv[5] *= s1.y; vv[6] *= s1.z; vv[7] *= s1.w;
        }
        if (g.row_scale) {
          const float rs = g.row_scale[m / g.rows_per_group];
#pragma unroll
          for (int i = 0; i < 8; ++i) vv[i] *= rs;
        }
        if (res) {
          float r[8];
          load8_as_f32<TC>(res + idx, r);
#pragma unroll
          for (int i = 0; i < 8; ++i) vv[i] += r[i];
        }
        store8_from_f32<TC>(C + idx, vv);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (p.tmem_cols == 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_base) : "memory");
    else if (p.tmem_cols == 64) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base) : "memory");
    else if (p.tmem_cols == 128) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// 2-D bf16 tensor [outer][inner] with row pitch ld elements; box {64, box_outer}; 128B swizzle; OOB -> 0
bool make_tmap(CUtensorMap* tm, const void* ptr, long long inner, long long outer, long long ld, int box_outer) {
  auto enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int pick_block_n(int N, bool b_trans) {
  if (b_trans) {  // MN-major B: whole 64-wide swizzle atoms
    if (N <= 64) return 64;
    if (N <= 128) return 128;
    if (N % 256 == 0 || N > 512) return 256;
    return (N % 192 == 0) ? 192 : 128;
  }
  if (N <= 256 && N % 16 == 0) return N;
  if (N <= 256) return ((N + 15) / 16) * 16;
  // prefer a divisor of N among multiples of 16 in [128, 256]
  for (int bn = 256; bn >= 128; bn -= 16)
    if (N % bn == 0) return bn;
  return 256;
}

template <typename TC>
int launch_tc(const GemmArgs& g, cudaStream_t st) {
  TcParams p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.a_trans = g.a_trans; p.b_trans = g.b_trans;
  p.block_n = pick_block_n(g.N, g.b_trans != 0);
  p.tmem_cols = p.block_n <= 32 ? 32 : p.block_n <= 64 ? 64 : p.block_n <= 128 ? 128 : 256;
  const int stage_bytes = A_STAGE_BYTES + p.block_n * BLOCK_K * 2;
  const int num_kb = (g.K + BLOCK_K - 1) / BLOCK_K;
  p.stages = p.block_n <= 128 ? 3 : 4;
  p.stages = max(2, min(p.stages, num_kb));
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 + (2 * p.stages + 1) * 8 + 16;

  const int gx = (g.N + p.block_n - 1) / p.block_n, gy = (g.M + BLOCK_M - 1) / BLOCK_M;
  int splits = 1;
  if (g.accumulate) {
    const int tiles = gx * gy;
    splits = max(1, min(num_kb / 8, (kNumSMs * 2 + tiles - 1) / tiles));
  }
  p.kb_per_split = (num_kb + splits - 1) / splits;
  splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;

  CUtensorMap tmA, tmB;
  bool ok;
  if (!g.a_trans) ok = make_tmap(&tmA, g.A, g.K, g.M, g.lda, BLOCK_M);
  else ok = make_tmap(&tmA, g.A, g.M, g.K, g.lda, 64);
  if (!g.b_trans) ok = ok && make_tmap(&tmB, g.B, g.K, g.N, g.ldb, p.block_n);
  else ok = ok && make_tmap(&tmB, g.B, g.N, g.K, g.ldb, 64);
  if (!ok) return LNX_ERR_UNSUPPORTED;

  auto kern = gemm_tc_kernel<TC>;
  static int smem_set[2] = {0, 0};
  int& cur = smem_set[sizeof(TC) == 4 ? 0 : 1];
  if ((int)smem > cur) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    cur = (int)smem;
  }
  dim3 grid(gx, gy, splits);
  kern<<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, g, p);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

}  // namespace

int lnx_gemm_tc2(const GemmArgs& g, float* colsum_out, cudaStream_t st);

int lnx_gemm_tc(const GemmArgs& g, int c_dtype, cudaStream_t st) {
  // shape gate: TMA needs 16-byte pitches/bases; the vector epilogue needs N % 8 == 0
  if (g.N % 8 != 0 || g.lda % 8 != 0 || g.ldb % 8 != 0) return LNX_ERR_UNSUPPORTED;
  if (!lnx_aligned16(g.A) || !lnx_aligned16(g.B) || !lnx_aligned16(g.C)) return LNX_ERR_UNSUPPORTED;
  if (!lnx_aligned16(g.aux_out) || !lnx_aligned16(g.act_grad_in) || !lnx_aligned16(g.residual) || !lnx_aligned16(g.bias) ||
      !lnx_aligned16(g.col_scale))
    return LNX_ERR_UNSUPPORTED;
  if (g.accumulate && c_dtype != LNX_F32) return LNX_ERR_DTYPE;
  if (c_dtype == LNX_BF16) return launch_tc<bf16>(g, st);
  if (c_dtype == LNX_F32) return launch_tc<float>(g, st);
  return LNX_ERR_DTYPE;
}

// qkv[M, 3 D] = (x[M, K] w[3 D, K]^T + bias) with the cos factors of the reference's 2-D "RoPE" and the softmax scale applied to the
// q / k columns in the epilogue (bf16, tcgen05 path only): replaces nn.Linear + the separate q / k scaling pass of
// rope_2d_mhsa.py:432-501.  freqs [2, D / 2] float32 = the learnable (fx, fy) per (head, pair); image tokens form a grid_w-wide grid
// after the n_extra leading tokens of each sequence of `tokens` rows.
extern "C" int lnx_qkv_rope_gemm(const void* x, const void* w, const float* bias, const float* freqs, void* qkv, int64_t M, int D, int K,
                                 int tokens, int n_extra, int grid_w, float q_scale, lnx_stream_t s) {
  LNX_REQUIRE(x && w && freqs && qkv, LNX_ERR_NULL);
  LNX_REQUIRE(M > 0 && M < (1ll << 31) && D > 0 && K > 0 && tokens > 0 && n_extra >= 0 && n_extra <= tokens && grid_w > 0, LNX_ERR_SHAPE);
  GemmArgs g;
  g.A = x; g.B = w; g.C = qkv; g.lda = K; g.ldb = K; g.M = (int)M; g.N = 3 * D; g.K = K;
  g.a_trans = 0; g.b_trans = 0;
  g.bias = bias; g.act = LNX_ACT_NONE; g.aux_out = nullptr; g.act_grad_in = nullptr; g.residual = nullptr; g.col_scale = nullptr;
  g.row_scale = nullptr; g.rows_per_group = 0; g.accumulate = 0;
  g.tok_scale = freqs; g.tok_period = tokens; g.tok_extra = n_extra; g.tok_dim = D; g.tok_w = grid_w; g.tok_qscale = q_scale;
  return lnx_gemm_tc2(g, nullptr, (cudaStream_t)s);
}

// ------------------------------------------------------------------ public dispatcher
extern "C" int lnx_gemm(int ab_dtype, const void* A, int64_t lda, int a_trans, const void* B, int64_t ldb, int b_trans, void* C,
                        int c_dtype, int M, int N, int K, const float* bias, int act, void* aux_out, const void* act_grad_in,
                        const void* residual, const float* col_scale, const float* row_scale, int rows_per_group, float* colsum_out,
                        int accumulate, int force_simt, lnx_stream_t s) {
  LNX_REQUIRE(A && B && C, LNX_ERR_NULL);
  LNX_REQUIRE(M > 0 && N > 0 && K > 0 && lda > 0 && ldb > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(!(accumulate && (bias || act || aux_out || act_grad_in || residual || col_scale || row_scale)), LNX_ERR_UNSUPPORTED);
  LNX_REQUIRE(!row_scale || rows_per_group > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(!accumulate || c_dtype == LNX_F32, LNX_ERR_DTYPE);
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.lda = lda; g.ldb = ldb; g.M = M; g.N = N; g.K = K;
  g.a_trans = a_trans ? 1 : 0; g.b_trans = b_trans ? 1 : 0;
  g.bias = bias; g.act = act; g.aux_out = aux_out; g.act_grad_in = act_grad_in; g.residual = residual; g.col_scale = col_scale; g.row_scale = row_scale; g.rows_per_group = rows_per_group;
  g.accumulate = accumulate ? 1 : 0;
  cudaStream_t st = (cudaStream_t)s;
  LNX_REQUIRE(!(colsum_out && accumulate), LNX_ERR_UNSUPPORTED);
  static const bool use_v1 = getenv("LNX_GEMM_V1") != nullptr;
  int r = LNX_ERR_UNSUPPORTED;
  if (ab_dtype == LNX_BF16 && !force_simt) {
    if (c_dtype == LNX_BF16 && !accumulate && !use_v1) {
      r = lnx_gemm_tc2(g, colsum_out, st);  // persistent kernel, TMA-store epilogue, fused column sums
      if (r == LNX_OK) return r;
    }
    if (r == LNX_ERR_UNSUPPORTED) r = lnx_gemm_tc(g, c_dtype, st);
  }
  if (r == LNX_ERR_UNSUPPORTED) r = lnx_gemm_simt(g, ab_dtype, c_dtype, st);
  if (r == LNX_OK && colsum_out) r = lnx_colsum(C, colsum_out, M, N, c_dtype, s);
  return r;
}
