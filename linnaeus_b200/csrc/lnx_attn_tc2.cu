// Attention for the mFormer token counts (N <= 240 forward / N <= 256 backward, head_dim 64) on tcgen05, second
// generation: persistent and software pipelined (forward), one CTA per (batch, head) with every gradient
// accumulated in TMEM (backward).
//
// Forward.  One CTA per SM walks (batch, head) problems; K and V of a problem are loaded once and serve all of its
// 128-query tiles.  Roles (320 threads):
//   warp 0      TMA producer: K/V into a 2-deep ring, Q tiles into a 2-deep ring
//   warp 1      MMA issuer:   S = Q K^T into one of two TMEM buffers (so the S of the NEXT tile is computed while the
//               softmax of this one runs), then O = P V into the first 64 columns of the same buffer
//   warps 2-9   softmax, TWO threads per query row (each owns half of the keys): exact two-pass softmax straight out
//               of TMEM (the whole key set is resident, no online rescaling), exp2 with raw MUFU, P as bf16 into
//               swizzled smem; they move on to the next tile as soon as P is published
//   warps 10-13 output: wait for O = P V, scale by 1/rowsum, write out and the LSE, release the TMEM buffer
// q is pre-scaled (softmax scale and cos factors folded in by the RoPE kernel).
//
// Backward.  CTA = one (batch, head): all K, V, Q, dO tiles of the problem are resident (<= 128 KB).  For every
// (key tile, query tile) pair: S = Q K^T and dP = dO V^T on the tensor core, P = exp(S - lse) and dS = P (dP - delta)
// by 256 threads (two per query row) as bf16 into smem, then dV += P^T dO, dK += dS^T Q, dQ[query tile] += dS K.
// TMEM holds S, dP, dV, dK and BOTH dQ tiles (512 columns exactly), so dQ never leaves the chip until it is final:
// no fp32 atomics, no zero fill, no cast pass.
#include <stdio.h>
#include <stdlib.h>

#include "lnx_tc_common.cuh"

using namespace lnx;
using namespace lnx_tc;

namespace {

constexpr int HD = 64;
constexpr int QT = 128;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// 16 consecutive columns [col0, col0 + 16) of row r of a K-major tile made of [128 rows][64 cols] 128B-swizzled blocks
__device__ __forceinline__ void store16_sw(unsigned char* tile, int r, int col0, const float* v) {
  unsigned char* blk = tile + (col0 >> 6) * (QT * 128);
  const int j = (col0 & 63) >> 3;
  *reinterpret_cast<uint4*>(blk + sw128_chunk(r, j)) = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
  *reinterpret_cast<uint4*>(blk + sw128_chunk(r, j + 1)) =
      make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct FwdParams {
  int B, heads, N, nkp;  // nkp = keys rounded up to 16
  int n_qt, n_bh;
  int c_split;           // first key column of the second softmax half (multiple of 16)
  int hd;                // head dim <= 64 (tiles are zero padded to 64 by the TMA out-of-bounds fill)
  int qkv4d;             // 1: q/k/v are read straight from a [B, N, 3, heads, hd] qkv matrix (4-D tensor maps)
  float scale;           // logits = scale * q k^T (+ bias); 1 when q arrives pre-scaled
  const float* bias;     // nullable, [heads, N, N] added to the scaled logits (relative position bias)
};

// barrier indices
enum { KV_FULL = 0, KV_EMPTY = 2, Q_FULL = 4, Q_EMPTY = 6, S_FULL = 8, O_FULL = 10, T_EMPTY = 12, P_FULL = 14, P_EMPTY = 15, NBARS = 16 };

template <bool BIAS>
__global__ void __launch_bounds__(448, 1) attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                               const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ out,
                                                               float* __restrict__ lse, const FwdParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  const int kv_bytes = p.nkp * 128;                    // one K or V tile [nkp keys][64 d]
  unsigned char* sKV = base;                           // [2]{K, V}
  unsigned char* sQ = sKV + 4 * (size_t)kv_bytes;      // [2][128][64]
  unsigned char* sP = sQ + 2 * 16384;                  // ceil(nkp / 64) blocks of [128][64]
  const int p_blocks = (p.nkp + 63) / 64;
  float* sMax = reinterpret_cast<float*>(sP + (size_t)p_blocks * 16384);  // [2 TMEM buffers][2 halves][128]
  float* sSum = sMax + 512;                                             // [2][2][128]
  float* sBias = sSum + 512;                                            // [8 softmax warps][32 rows][17] (bias staging, BIAS only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + (BIAS ? 8 * 32 * 17 : 0));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_bh = ((int)blockIdx.x < p.n_bh) ? (p.n_bh - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int T = my_bh * p.n_qt;  // (problem, query tile) items of this CTA

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[KV_FULL + i], 1);
      mbar_init(&bars[KV_EMPTY + i], 1);
      mbar_init(&bars[Q_FULL + i], 1);
      mbar_init(&bars[Q_EMPTY + i], 1);
      mbar_init(&bars[S_FULL + i], 1);
      mbar_init(&bars[O_FULL + i], 1);
      mbar_init(&bars[T_EMPTY + i], 128);
    }
    mbar_init(&bars[P_FULL], 256);
    mbar_init(&bars[P_EMPTY], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int t = 0;
      for (int j = 0; j < my_bh; ++j) {
        const int bh = (int)blockIdx.x + j * (int)gridDim.x;
        const int kvs = j & 1;
        mbar_wait_relaxed(&bars[KV_EMPTY + kvs], (((uint32_t)j >> 1) & 1u) ^ 1u);
        mbar_expect_tx(&bars[KV_FULL + kvs], 2 * (uint32_t)kv_bytes);
        const int b_ = bh / p.heads, h_ = bh - b_ * p.heads;
        if (p.qkv4d) {
          tma_load_4d(sKV + (size_t)kvs * 2 * kv_bytes, &tmK, &bars[KV_FULL + kvs], 0, h_, 0, b_);
          tma_load_4d(sKV + (size_t)kvs * 2 * kv_bytes + kv_bytes, &tmV, &bars[KV_FULL + kvs], 0, h_, 0, b_);
        } else {
          tma_load_3d(sKV + (size_t)kvs * 2 * kv_bytes, &tmK, &bars[KV_FULL + kvs], 0, 0, bh);
          tma_load_3d(sKV + (size_t)kvs * 2 * kv_bytes + kv_bytes, &tmV, &bars[KV_FULL + kvs], 0, 0, bh);
        }
        for (int qt = 0; qt < p.n_qt; ++qt, ++t) {
          const int qs = t & 1;
          mbar_wait_relaxed(&bars[Q_EMPTY + qs], (((uint32_t)t >> 1) & 1u) ^ 1u);
          mbar_expect_tx(&bars[Q_FULL + qs], 16384);
          if (p.qkv4d) tma_load_4d(sQ + qs * 16384, &tmQ, &bars[Q_FULL + qs], 0, h_, qt * QT, b_);
          else tma_load_3d(sQ + qs * 16384, &tmQ, &bars[Q_FULL + qs], 0, qt * QT, bh);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(QT, p.nkp, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(QT, HD, 0, 1);
      const uint32_t aP = smem_u32(sP);
      auto issue_s = [&](int t) {
        const int j = t / p.n_qt, qt = t - j * p.n_qt;
        const int kvs = j & 1, qs = t & 1, buf = t & 1;
        if (qt == 0) mbar_wait(&bars[KV_FULL + kvs], ((uint32_t)j >> 1) & 1u);
        mbar_wait(&bars[Q_FULL + qs], ((uint32_t)t >> 1) & 1u);
        mbar_wait(&bars[T_EMPTY + buf], (((uint32_t)t >> 1) & 1u) ^ 1u);  // epilogue of item t - 2 has drained this buffer
        tcgen05_fence_after();
        const uint32_t aq = smem_u32(sQ + qs * 16384), ak = smem_u32(sKV + (size_t)kvs * 2 * kv_bytes);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem + buf * 256, make_smem_desc(aq + k * 32, 0, 1024), make_smem_desc(ak + k * 32, 0, 1024), idesc_s, k > 0);
        umma_commit(&bars[S_FULL + buf]);
        umma_commit(&bars[Q_EMPTY + qs]);
      };
      if (T > 0) issue_s(0);
      for (int t = 0; t < T; ++t) {
        if (t + 1 < T) issue_s(t + 1);
        const int j = t / p.n_qt, qt = t - j * p.n_qt;
        const int kvs = j & 1, buf = t & 1;
        mbar_wait(&bars[P_FULL], (uint32_t)t & 1u);
        tcgen05_fence_after();
        const uint32_t av = smem_u32(sKV + (size_t)kvs * 2 * kv_bytes + kv_bytes);
        for (int k = 0; k < p.nkp / 16; ++k)
          umma_bf16(tmem + buf * 256, make_smem_desc(aP + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024), make_smem_desc(av + k * 2048, 0, 1024),
                    idesc_o, k > 0);
        umma_commit(&bars[O_FULL + buf]);
        umma_commit(&bars[P_EMPTY]);
        if (qt == p.n_qt - 1) umma_commit(&bars[KV_EMPTY + kvs]);
      }
    }
  } else if (warp < 10) {
    // ===================== softmax: two threads per query row =====================
    const int sw = warp - 2;
    const int q = warp & 3;       // TMEM lane quarter
    const int half = sw >> 2;     // which half of the keys
    const int r = q * 32 + lane;  // query row inside the tile
    const int c_begin = half ? p.c_split : 0, c_end = half ? p.nkp : p.c_split;
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      const uint32_t trow = tmem + buf * 256 + ((uint32_t)(q * 32) << 16);
      float* bMax = sMax + buf * 256;
      float* bSum = sSum + buf * 256;
      mbar_wait(&bars[S_FULL + buf], ((uint32_t)t >> 1) & 1u);
      tcgen05_fence_after();
      // logits = scale * S (+ bias), handled in the log2 domain.  BIAS: the [32 rows x 16 keys] bias chunk of this warp is
      // staged through shared memory with coalesced 16-byte loads (a lane owns one ROW, so direct loads would touch 32
      // cache lines per instruction).  The loads run two chunks ahead of their use (registers), and pass 1 writes the
      // biased logits back into the S columns of TMEM so pass 2 needs neither the bias nor the scale again.
      const float sl = p.scale * LOG2E;
      float* wb = sBias + sw * (32 * 17);
      const float* bhead = nullptr;
      int qt_ = 0;
      if constexpr (BIAS) {
        const int j_ = t / p.n_qt;
        qt_ = t - j_ * p.n_qt;
        bhead = p.bias + (long long)(((int)blockIdx.x + j_ * (int)gridDim.x) % p.heads) * p.N * p.N;
      }
      auto load_bias = [&](int c, float4* dst) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rl = it * 8 + (lane >> 2), c4 = c + (lane & 3) * 4;
          const int gm = min(qt_ * QT + q * 32 + rl, p.N - 1);  // rows past N read row N - 1 (never stored)
          dst[it] = (c < c_end && c4 < p.N) ? __ldg(reinterpret_cast<const float4*>(bhead + (long long)gm * p.N + c4))  // N % 4 == 0
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto stage_bias = [&](const float4* src) {
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          float* d = wb + (it * 8 + (lane >> 2)) * 17 + (lane & 3) * 4;
          d[0] = src[it].x; d[1] = src[it].y; d[2] = src[it].z; d[3] = src[it].w;
        }
        __syncwarp();
      };
      // pass 1: row max over this half's valid keys
      float mx = -INFINITY;
      float4 b0[4], b1[4], b2[4];
      if constexpr (BIAS) {
        load_bias(c_begin, b0);
        load_bias(c_begin + 16, b1);
      }
      for (int c = c_begin; c < c_end; c += 16) {
        float v[16];
        if constexpr (BIAS) load_bias(c + 32, b2);
        tmem_ld16(trow + c, v);
        if constexpr (BIAS) {
          stage_bias(b0);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], sl, wb[lane * 17 + i] * LOG2E);
          tmem_st16(trow + c, v);
#pragma unroll
          for (int it = 0; it < 4; ++it) b0[it] = b1[it], b1[it] = b2[it];
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= sl;
        }
        if (c + 16 <= p.N) {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c + i < p.N) mx = fmaxf(mx, v[i]);
        }
      }
      if constexpr (BIAS) tmem_st_wait();
      bMax[half * 128 + r] = mx;
      named_bar(1, 256);
      mx = fmaxf(mx, bMax[(half ^ 1) * 128 + r]);
      const float mxl = mx;
      mbar_wait(&bars[P_EMPTY], ((uint32_t)t & 1u) ^ 1u);  // P V of the previous item has consumed the P tile
      // pass 2: P = exp2(logit - max) -> bf16 smem, row sum in fp32
      float sum = 0.f;
      for (int c = c_begin; c < c_end; c += 16) {
        float v[16];
        tmem_ld16(trow + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = BIAS ? v[i] - mxl : fmaf(v[i], sl, -mxl);  // BIAS: TMEM already holds the biased log2-domain logit
          float e = ex2f(a);
          if (c + 16 > p.N && c + i >= p.N) e = 0.f;
          sum += e;
          v[i] = e;
        }
        store16_sw(sP, r, c, v);
      }
      bSum[half * 128 + r] = sum;
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive(&bars[P_FULL]);
    }
  } else {
    // ===================== output: one thread per query row =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    for (int t = 0; t < T; ++t) {
      const int j = t / p.n_qt, qt = t - j * p.n_qt;
      const int bh = (int)blockIdx.x + j * (int)gridDim.x;
      const int buf = t & 1;
      const int m = qt * QT + r;
      const uint32_t trow = tmem + buf * 256 + ((uint32_t)(q * 32) << 16);
      mbar_wait(&bars[O_FULL + buf], ((uint32_t)t >> 1) & 1u);  // implies P_FULL: both halves' max / sum are published
      tcgen05_fence_after();
      const float sum = sSum[buf * 256 + r] + sSum[buf * 256 + 128 + r];
      const float mx = fmaxf(sMax[buf * 256 + r], sMax[buf * 256 + 128 + r]);
      const float inv = 1.0f / sum;
      const int b = bh / p.heads, h = bh - b * p.heads;
      for (int c = 0; c < p.hd; c += 16) {
        float v[16];
        __syncwarp();
        tmem_ld16(trow + c, v);
        if (m < p.N) {
          bf16* dst = out + (((long long)b * p.N + m) * p.heads + h) * p.hd + c;
          *reinterpret_cast<uint4*>(dst) = make_uint4(pack2(v[0] * inv, v[1] * inv), pack2(v[2] * inv, v[3] * inv), pack2(v[4] * inv, v[5] * inv),
                                                      pack2(v[6] * inv, v[7] * inv));
          *reinterpret_cast<uint4*>(dst + 8) = make_uint4(pack2(v[8] * inv, v[9] * inv), pack2(v[10] * inv, v[11] * inv),
                                                          pack2(v[12] * inv, v[13] * inv), pack2(v[14] * inv, v[15] * inv));
        }
      }
      if (lse && m < p.N) lse[(long long)bh * p.N + m] = mx * (1.0f / LOG2E) + __logf(sum);  // mx is kept in the log2 domain
      tcgen05_fence_before();
      mbar_arrive(&bars[T_EMPTY + buf]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------ forward, long sequences (N > 240)
// Same roles as above, but the keys are streamed in tiles of 128 with an online softmax: an item is one
// (problem, query tile, key tile).  The softmax warps keep the running row max / row sum in registers and publish the
// rescale factor alpha = exp(m_old - m_new) of every item; the output warps own the fp32 output accumulator in
// REGISTERS (O = O * alpha + P_j V_j, P_j V_j read from TMEM after each non-accumulating MMA), so TMEM is never
// written by threads.  K/V tiles come through a 3-deep ring (re-read per query tile: they hit in L2).
struct LongParams {
  int B, heads, N;
  int n_qt, n_kt, n_bh;
};
constexpr int KT = 128;
constexpr int L_KV_STAGES = 3;
enum { L_KV_FULL = 0, L_KV_EMPTY = 3, L_Q_FULL = 6, L_Q_EMPTY = 8, L_S_FULL = 10, L_O_FULL = 12, L_T_EMPTY = 14, L_P_FULL = 16, L_P_EMPTY = 17, L_NBARS = 18 };

__global__ void __launch_bounds__(448, 1) attn_fwd_long_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                                    const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ out,
                                                                    float* __restrict__ lse, const LongParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* sKV = base;                               // [3]{K tile, V tile} of [128][64]
  unsigned char* sQ = sKV + L_KV_STAGES * 32768;           // [2][128][64]
  unsigned char* sP = sQ + 2 * 16384;                      // 2 blocks of [128][64]
  float* sMax = reinterpret_cast<float*>(sP + 32768);      // [2 halves][128]  (per item, exchanged inside the softmax warps)
  float* sAlpha = sMax + 256;                              // [2 TMEM buffers][128]  rescale factor of the item
  float* sFin = sAlpha + 256;                              // [2 buffers][3][128]: final max, row sum of half 0 / half 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(sFin + 768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + L_NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = p.n_bh * p.n_qt;  // (problem, query tile)
  const int my_units = ((int)blockIdx.x < n_units) ? (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int T = my_units * p.n_kt;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    for (int i = 0; i < L_KV_STAGES; ++i) {
      mbar_init(&bars[L_KV_FULL + i], 1);
      mbar_init(&bars[L_KV_EMPTY + i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[L_Q_FULL + i], 1);
      mbar_init(&bars[L_Q_EMPTY + i], 1);
      mbar_init(&bars[L_S_FULL + i], 1);
      mbar_init(&bars[L_O_FULL + i], 1);
      mbar_init(&bars[L_T_EMPTY + i], 128);
    }
    mbar_init(&bars[L_P_FULL], 256);
    mbar_init(&bars[L_P_EMPTY], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int t = 0;
      for (int u = 0; u < my_units; ++u) {
        const int unit = (int)blockIdx.x + u * (int)gridDim.x;
        const int bh = unit / p.n_qt, qt = unit - bh * p.n_qt;
        const int qs = u & 1;
        mbar_wait_relaxed(&bars[L_Q_EMPTY + qs], (((uint32_t)u >> 1) & 1u) ^ 1u);
        mbar_expect_tx(&bars[L_Q_FULL + qs], 16384);
        tma_load_3d(sQ + qs * 16384, &tmQ, &bars[L_Q_FULL + qs], 0, qt * QT, bh);
        for (int j = 0; j < p.n_kt; ++j, ++t) {
          const int ks = t % L_KV_STAGES;
          mbar_wait_relaxed(&bars[L_KV_EMPTY + ks], (((uint32_t)(t / L_KV_STAGES)) & 1u) ^ 1u);
          mbar_expect_tx(&bars[L_KV_FULL + ks], 32768);
          tma_load_3d(sKV + ks * 32768, &tmK, &bars[L_KV_FULL + ks], 0, j * KT, bh);
          tma_load_3d(sKV + ks * 32768 + 16384, &tmV, &bars[L_KV_FULL + ks], 0, j * KT, bh);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(QT, KT, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(QT, HD, 0, 1);
      const uint32_t aP = smem_u32(sP);
      auto issue_s = [&](int t) {
        const int u = t / p.n_kt, j = t - u * p.n_kt;
        const int ks = t % L_KV_STAGES, qs = u & 1, buf = t & 1;
        if (j == 0) mbar_wait(&bars[L_Q_FULL + qs], ((uint32_t)u >> 1) & 1u);
        mbar_wait(&bars[L_KV_FULL + ks], ((uint32_t)(t / L_KV_STAGES)) & 1u);
        mbar_wait(&bars[L_T_EMPTY + buf], (((uint32_t)t >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t aq = smem_u32(sQ + qs * 16384), ak = smem_u32(sKV + ks * 32768);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem + buf * 256, make_smem_desc(aq + k * 32, 0, 1024), make_smem_desc(ak + k * 32, 0, 1024), idesc_s, k > 0);
        umma_commit(&bars[L_S_FULL + buf]);
        if (j == p.n_kt - 1) umma_commit(&bars[L_Q_EMPTY + qs]);
      };
      if (T > 0) issue_s(0);
      for (int t = 0; t < T; ++t) {
        if (t + 1 < T) issue_s(t + 1);
        const int ks = t % L_KV_STAGES, buf = t & 1;
        mbar_wait(&bars[L_P_FULL], (uint32_t)t & 1u);
        tcgen05_fence_after();
        const uint32_t av = smem_u32(sKV + ks * 32768 + 16384);
#pragma unroll
        for (int k = 0; k < KT / 16; ++k)
          umma_bf16(tmem + buf * 256, make_smem_desc(aP + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024), make_smem_desc(av + k * 2048, 0, 1024),
                    idesc_o, k > 0);
        umma_commit(&bars[L_O_FULL + buf]);
        umma_commit(&bars[L_P_EMPTY]);
        umma_commit(&bars[L_KV_EMPTY + ks]);
      }
    }
  } else if (warp < 10) {
    // ===================== softmax: two threads per query row, 64 keys of the tile each =====================
    const int sw = warp - 2;
    const int q = warp & 3;
    const int half = sw >> 2;
    const int r = q * 32 + lane;
    float m_run = -INFINITY, l_run = 0.f;  // running max (shared by both halves), running sum of this half
    for (int t = 0; t < T; ++t) {
      const int u = t / p.n_kt, j = t - u * p.n_kt;
      const int buf = t & 1;
      const uint32_t trow = tmem + buf * 256 + ((uint32_t)(q * 32) << 16);
      const int key0 = j * KT + half * 64;
      if (j == 0) {
        m_run = -INFINITY;
        l_run = 0.f;
      }
      mbar_wait(&bars[L_S_FULL + buf], ((uint32_t)t >> 1) & 1u);
      tcgen05_fence_after();
      float mx = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float v[16];
        tmem_ld16(trow + half * 64 + cc * 16, v);
        const int kb = key0 + cc * 16;
        if (kb + 16 <= p.N) {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (kb + i < p.N) mx = fmaxf(mx, v[i]);
        }
      }
      sMax[half * 128 + r] = mx;
      named_bar(1, 256);
      mx = fmaxf(mx, sMax[(half ^ 1) * 128 + r]);
      const float m_new = fmaxf(m_run, mx);  // finite: key tile 0 always holds valid keys
      const float alpha = ex2f((m_run - m_new) * LOG2E);  // 0 on the first tile (m_run = -inf)
      m_run = m_new;
      const float mxl = m_new * LOG2E;
      mbar_wait(&bars[L_P_EMPTY], ((uint32_t)t & 1u) ^ 1u);
      float sum = 0.f;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float v[16];
        tmem_ld16(trow + half * 64 + cc * 16, v);
        const int kb = key0 + cc * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float e = ex2f(fmaf(v[i], LOG2E, -mxl));
          if (kb + 16 > p.N && kb + i >= p.N) e = 0.f;
          sum += e;
          v[i] = e;
        }
        store16_sw(sP, r, half * 64 + cc * 16, v);
      }
      l_run = fmaf(l_run, alpha, sum);
      if (half == 0) sAlpha[buf * 128 + r] = alpha;
      if (j == p.n_kt - 1) {
        if (half == 0) sFin[buf * 384 + r] = m_run;
        sFin[buf * 384 + 128 + half * 128 + r] = l_run;
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive(&bars[L_P_FULL]);
      named_bar(2, 256);  // sMax is rewritten by the next item
    }
  } else {
    // ===================== output: one thread per query row, fp32 accumulator in registers =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    float acc[HD];
    for (int t = 0; t < T; ++t) {
      const int u = t / p.n_kt, j = t - u * p.n_kt;
      const int buf = t & 1;
      const uint32_t trow = tmem + buf * 256 + ((uint32_t)(q * 32) << 16);
      mbar_wait(&bars[L_O_FULL + buf], ((uint32_t)t >> 1) & 1u);
      tcgen05_fence_after();
      const float alpha = (j == 0) ? 0.f : sAlpha[buf * 128 + r];
#pragma unroll
      for (int c = 0; c < HD; c += 16) {
        float v[16];
        __syncwarp();
        tmem_ld16(trow + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[c + i] = (j == 0) ? v[i] : fmaf(acc[c + i], alpha, v[i]);
      }
      if (j == p.n_kt - 1) {
        const int unit = (int)blockIdx.x + u * (int)gridDim.x;
        const int bh = unit / p.n_qt, qt = unit - bh * p.n_qt;
        const int m = qt * QT + r;
        const float mx = sFin[buf * 384 + r];
        const float sum = sFin[buf * 384 + 128 + r] + sFin[buf * 384 + 256 + r];
        const float inv = 1.0f / sum;
        const int b = bh / p.heads, h = bh - b * p.heads;
        if (m < p.N) {
          bf16* dst = out + (((long long)b * p.N + m) * p.heads + h) * HD;
#pragma unroll
          for (int c = 0; c < HD; c += 8)
            *reinterpret_cast<uint4*>(dst + c) = make_uint4(pack2(acc[c] * inv, acc[c + 1] * inv), pack2(acc[c + 2] * inv, acc[c + 3] * inv),
                                                            pack2(acc[c + 4] * inv, acc[c + 5] * inv), pack2(acc[c + 6] * inv, acc[c + 7] * inv));
          lse[(long long)bh * p.N + m] = mx + __logf(sum);
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&bars[L_T_EMPTY + buf]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ------------------------------------------------------------------ backward
// TMEM columns: S [0,128) | dP [128,256) | dV [256,320) | dK [320,384) | dQ tile 0 [384,448) | dQ tile 1 [448,512)
struct BwdParams {
  int B, heads, N;
  int n_t;    // 128-row tiles of the sequence (1 or 2)
  int qkv4d;  // 1: q / k / v are read straight from the [B, N, 3, heads, 64] qkv matrix (4-D tensor maps)
};
enum { B_LOAD = 0, B_SDP = 1, B_PDS = 2, B_ACC = 3, B_NBARS = 4 };

__global__ void __launch_bounds__(320, 1) attn_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                                                               const bf16* __restrict__ o, const bf16* __restrict__ dout,
                                                               const float* __restrict__ lse, bf16* __restrict__ dq, bf16* __restrict__ dk,
                                                               bf16* __restrict__ dv, const BwdParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* sK = base;                    // [n_t][128 keys][128 B]
  unsigned char* sV = sK + p.n_t * 16384;
  unsigned char* sQ = sV + p.n_t * 16384;      // [n_t][128 queries][128 B]
  unsigned char* sdO = sQ + p.n_t * 16384;
  unsigned char* sP = sdO + p.n_t * 16384;     // [128 queries][128 keys] = 2 blocks x 16 KB
  unsigned char* sdS = sP + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x;
  const int b = bh / p.heads, h = bh - b * p.heads;
  const int n_it = p.n_t * p.n_t;  // (key tile, query tile) pairs, key tile outer

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmdO);
    mbar_init(&bars[B_LOAD], 1);
    mbar_init(&bars[B_SDP], 1);
    mbar_init(&bars[B_PDS], 256);
    mbar_init(&bars[B_ACC], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&bars[B_LOAD], (uint32_t)(4 * p.n_t * 16384));
      for (int i = 0; i < p.n_t; ++i) {
        if (p.qkv4d) {
          tma_load_4d(sK + i * 16384, &tmK, &bars[B_LOAD], 0, h, i * QT, b);
          tma_load_4d(sV + i * 16384, &tmV, &bars[B_LOAD], 0, h, i * QT, b);
          tma_load_4d(sQ + i * 16384, &tmQ, &bars[B_LOAD], 0, h, i * QT, b);
        } else {
          tma_load_3d(sK + i * 16384, &tmK, &bars[B_LOAD], 0, i * QT, bh);
          tma_load_3d(sV + i * 16384, &tmV, &bars[B_LOAD], 0, i * QT, bh);
          tma_load_3d(sQ + i * 16384, &tmQ, &bars[B_LOAD], 0, i * QT, bh);
        }
        tma_load_3d(sdO + i * 16384, &tmdO, &bars[B_LOAD], h * HD, i * QT, b);  // dout is [B][N][heads*64]
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t aP = smem_u32(sP), aS = smem_u32(sdS);
      const uint32_t id_s = make_idesc_bf16(QT, QT, 0, 0);   // S, dP: [128 q] x [128 keys], K = 64
      const uint32_t id_kv = make_idesc_bf16(QT, HD, 1, 1);  // dV, dK: A = P^T / dS^T (MN-major), B = dO / Q (MN-major)
      const uint32_t id_q = make_idesc_bf16(QT, HD, 0, 1);   // dQ: A = dS (K-major), B = K (MN-major)
      mbar_wait(&bars[B_LOAD], 0);
      tcgen05_fence_after();
      for (int it = 0; it < n_it; ++it) {
        const int kt = it / p.n_t, qt = it - kt * p.n_t;
        const uint32_t ph = (uint32_t)it & 1u;
        const uint32_t aK = smem_u32(sK + kt * 16384), aV = smem_u32(sV + kt * 16384);
        const uint32_t aQ = smem_u32(sQ + qt * 16384), aO = smem_u32(sdO + qt * 16384);
        // S / dP TMEM of the previous pair was drained before its compute threads arrived on B_PDS (waited below), and
        // MMAs execute in issue order, so the next S / dP can be queued right behind the previous pair's dV / dK / dQ
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem + 0, make_smem_desc(aQ + k * 32, 0, 1024), make_smem_desc(aK + k * 32, 0, 1024), id_s, k > 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem + 128, make_smem_desc(aO + k * 32, 0, 1024), make_smem_desc(aV + k * 32, 0, 1024), id_s, k > 0);
        umma_commit(&bars[B_SDP]);
        mbar_wait(&bars[B_PDS], ph);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < QT / 16; ++k) {  // reduction over the 128 queries of this tile
          const uint32_t acc = (qt > 0 || k > 0) ? 1u : 0u;
          umma_bf16(tmem + 256, make_smem_desc(aP + k * 2048, 16384, 1024), make_smem_desc(aO + k * 2048, 0, 1024), id_kv, acc);
          umma_bf16(tmem + 320, make_smem_desc(aS + k * 2048, 16384, 1024), make_smem_desc(aQ + k * 2048, 0, 1024), id_kv, acc);
        }
#pragma unroll
        for (int k = 0; k < QT / 16; ++k)  // reduction over the 128 keys of this key tile
          umma_bf16(tmem + 384 + qt * 64, make_smem_desc(aS + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024), make_smem_desc(aK + k * 2048, 0, 1024),
                    id_q, (kt > 0 || k > 0) ? 1u : 0u);
        umma_commit(&bars[B_ACC]);
      }
    }
  } else {
    // 256 compute threads: two per query row, each owns 64 of the 128 keys of the pair
    const int sw = warp - 2;
    const int q = warp & 3;
    const int half = sw >> 2;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    // per query tile: lse (log2 domain) and delta = rowsum(dO * O) of this thread's row
    float Ls[2] = {0.f, 0.f}, deltas[2] = {0.f, 0.f};
    for (int qt = 0; qt < p.n_t; ++qt) {
      const int m = qt * QT + r;
      if (m < p.N) {
        Ls[qt] = lse[(long long)bh * p.N + m] * LOG2E;
        const long long oidx = (((long long)b * p.N + m) * p.heads + h) * HD;
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < HD; c += 8) {
          const uint4 ra = __ldg(reinterpret_cast<const uint4*>(o + oidx + c));
          const uint4 rb = __ldg(reinterpret_cast<const uint4*>(dout + oidx + c));
          const uint32_t* pa = reinterpret_cast<const uint32_t*>(&ra);
          const uint32_t* pb = reinterpret_cast<const uint32_t*>(&rb);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            d = fmaf(__uint_as_float(pa[e] << 16), __uint_as_float(pb[e] << 16), d);
            d = fmaf(__uint_as_float(pa[e] & 0xffff0000u), __uint_as_float(pb[e] & 0xffff0000u), d);
          }
        }
        deltas[qt] = d;
      }
    }
    for (int it = 0; it < n_it; ++it) {
      const int kt = it / p.n_t, qt = it - kt * p.n_t;
      const uint32_t ph = (uint32_t)it & 1u;
      const int m = qt * QT + r;
      const float L = qt ? Ls[1] : Ls[0], delta = qt ? deltas[1] : deltas[0];
      const bool row_ok = m < p.N;
      const int key0 = kt * QT + half * 64;
      mbar_wait(&bars[B_SDP], ph);
      tcgen05_fence_after();
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = half * 64 + cc * 16;
        float s[16], dp[16];
        __syncwarp();
        tmem_ld16(trow + c, s);
        tmem_ld16(trow + 128 + c, dp);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const bool ok = row_ok && (key0 + cc * 16 + e < p.N);
          const float pr = ok ? ex2f(fmaf(s[e], LOG2E, -L)) : 0.f;
          s[e] = pr;
          dp[e] = pr * (dp[e] - delta);
        }
        store16_sw(sP, r, c, s);
        store16_sw(sdS, r, c, dp);
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive(&bars[B_PDS]);
      // after the last query tile of this key tile: dV, dK are final (rows = keys kt*128 + r; each half stores 32 columns)
      if (qt == p.n_t - 1) {
        mbar_wait(&bars[B_ACC], ph);
        tcgen05_fence_after();
        const int key = kt * QT + r;
#pragma unroll
        for (int tsel = 0; tsel < 2; ++tsel) {
          bf16* dst_base = (tsel == 0 ? dv : dk);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = half * 32 + cc * 16;
            float v[16];
            __syncwarp();
            tmem_ld16(trow + 256 + tsel * 64 + c, v);
            if (key < p.N) {
              bf16* dst = dst_base + ((long long)bh * p.N + key) * HD + c;
              *reinterpret_cast<uint4*>(dst) = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
              *reinterpret_cast<uint4*>(dst + 8) = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
            }
          }
        }
        tcgen05_fence_before();
      }
    }
    // dQ of both query tiles is final after the last pair
    mbar_wait(&bars[B_ACC], (uint32_t)(n_it - 1) & 1u);
    tcgen05_fence_after();
    for (int qt = 0; qt < p.n_t; ++qt) {
      const int m = qt * QT + r;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 32 + cc * 16;
        float v[16];
        __syncwarp();
        tmem_ld16(trow + 384 + qt * 64 + c, v);
        if (m < p.N) {
          bf16* dst = dq + ((long long)bh * p.N + m) * HD + c;
          *reinterpret_cast<uint4*>(dst) = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
          *reinterpret_cast<uint4*>(dst + 8) = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

bool head_tmap(CUtensorMap* tm, const void* ptr, int BH, int N, int box_rows) {
  const long long dims[3] = {HD, N, BH};
  const long long strides[2] = {HD, (long long)N * HD};
  const int box[3] = {HD, box_rows, 1};
  return make_tmap(tm, ptr, 3, dims, strides, box);
}

int g_fwd_smem_set = 0, g_fwd_bias_smem_set = 0;  // largest dynamic shared-memory size each attn_fwd_tc2_kernel variant was configured for

}  // namespace

int lnx_attn_fwd_tc2(const void* q, const void* k, const void* v, void* out, float* lse, int B, int heads, int N, int hd, cudaStream_t st) {
  if (hd != HD || N > 240 || N < 1) return LNX_ERR_UNSUPPORTED;
  FwdParams p;
  p.B = B; p.heads = heads; p.N = N;
  p.nkp = (N + 15) / 16 * 16;
  p.n_qt = (N + QT - 1) / QT;
  p.n_bh = B * heads;
  p.c_split = ((p.nkp / 16 + 1) / 2) * 16;
  p.hd = HD; p.qkv4d = 0; p.scale = 1.0f; p.bias = nullptr;
  CUtensorMap tq, tk, tv;
  if (!head_tmap(&tq, q, p.n_bh, N, QT) || !head_tmap(&tk, k, p.n_bh, N, p.nkp) || !head_tmap(&tv, v, p.n_bh, N, p.nkp)) return LNX_ERR_UNSUPPORTED;
  const size_t smem = 1024 + 4 * (size_t)p.nkp * 128 + 2 * 16384 + (size_t)((p.nkp + 63) / 64) * 16384 + 8 * 128 * 4 + NBARS * 8 + 64;
  if (smem > 232448) return LNX_ERR_UNSUPPORTED;
  if ((int)smem > g_fwd_smem_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    g_fwd_smem_set = (int)smem;
  }
  attn_fwd_tc2_kernel<false><<<min(p.n_bh, kNumSMs), 448, smem, st>>>(tq, tk, tv, (bf16*)out, lse, p);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

// softmax(scale q k^T + bias[h]) v with q/k/v read straight from qkv [B, N, 3, heads, hd] (hd <= 64, N <= 240): the
// RelativeAttention of mFormerV0.  Head dims below 64 ride on the TMA out-of-bounds zero fill (tiles stay [rows][64]).
int lnx_attn_bias_fwd_tc2(const void* qkv, const float* bias, void* out, float* lse, int B, int heads, int N, int hd, float scale, cudaStream_t st) {
  if (hd > HD || hd % 16 != 0 || N > 240 || N < 1) return LNX_ERR_UNSUPPORTED;
  if (!lnx_aligned16(qkv) || !lnx_aligned16(out)) return LNX_ERR_UNSUPPORTED;
  if (bias && (N % 4 != 0 || !lnx_aligned16(bias))) return LNX_ERR_UNSUPPORTED;  // the bias rows are staged with 16-byte loads
  FwdParams p;
  p.B = B; p.heads = heads; p.N = N;
  p.nkp = (N + 15) / 16 * 16;
  p.n_qt = (N + QT - 1) / QT;
  p.n_bh = B * heads;
  p.c_split = ((p.nkp / 16 + 1) / 2) * 16;
  p.hd = hd; p.qkv4d = 1; p.scale = scale; p.bias = bias;
  CUtensorMap tm[3];
  for (int which = 0; which < 3; ++which) {
    const long long dims[4] = {hd, heads, N, B};
    const long long strides[3] = {hd, 3LL * heads * hd, (long long)N * 3 * heads * hd};
    const int box[4] = {HD, 1, which == 0 ? QT : p.nkp, 1};
    if (!make_tmap(&tm[which], reinterpret_cast<const bf16*>(qkv) + (long long)which * heads * hd, 4, dims, strides, box)) {
      if (getenv("LNX_ATTN_NO_FALLBACK")) fprintf(stderr, "lnx_attn_bias_fwd_tc2: tensor map %d failed (hd %d heads %d N %d B %d)\n", which, hd, heads, N, B);
      return LNX_ERR_UNSUPPORTED;
    }
  }
  const size_t smem = 1024 + 4 * (size_t)p.nkp * 128 + 2 * 16384 + (size_t)((p.nkp + 63) / 64) * 16384 + 8 * 128 * 4 + (bias ? 8 * 32 * 17 * 4 : 0) +
                      NBARS * 8 + 64;
  if (smem > 232448) return LNX_ERR_UNSUPPORTED;
  int& set = bias ? g_fwd_bias_smem_set : g_fwd_smem_set;
  if ((int)smem > set) {
    cudaError_t e = bias ? cudaFuncSetAttribute(attn_fwd_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                         : cudaFuncSetAttribute(attn_fwd_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    set = (int)smem;
  }
  if (bias) attn_fwd_tc2_kernel<true><<<min(p.n_bh, kNumSMs), 448, smem, st>>>(tm[0], tm[1], tm[2], (bf16*)out, lse, p);
  else attn_fwd_tc2_kernel<false><<<min(p.n_bh, kNumSMs), 448, smem, st>>>(tm[0], tm[1], tm[2], (bf16*)out, lse, p);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

int lnx_attn_fwd_long_tc2(const void* q, const void* k, const void* v, void* out, float* lse, int B, int heads, int N, int hd, cudaStream_t st) {
  if (hd != HD || N < 1) return LNX_ERR_UNSUPPORTED;
  LongParams p;
  p.B = B; p.heads = heads; p.N = N;
  p.n_qt = (N + QT - 1) / QT;
  p.n_kt = (N + KT - 1) / KT;
  p.n_bh = B * heads;
  CUtensorMap tq, tk, tv;
  if (!head_tmap(&tq, q, p.n_bh, N, QT) || !head_tmap(&tk, k, p.n_bh, N, KT) || !head_tmap(&tv, v, p.n_bh, N, KT)) return LNX_ERR_UNSUPPORTED;
  const size_t smem = 1024 + L_KV_STAGES * 32768 + 2 * 16384 + 32768 + (256 + 256 + 768) * 4 + L_NBARS * 8 + 64;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_long_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    attr = true;
  }
  attn_fwd_long_tc2_kernel<<<min(p.n_bh * p.n_qt, kNumSMs), 448, smem, st>>>(tq, tk, tv, (bf16*)out, lse, p);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

// dq / dk / dv are written as bf16 (no accumulation, no workspace), head-major [B * heads, N, 64].  qkv != nullptr: q / k / v are read
// straight from the [B, N, 3, heads, 64] matrix the (RoPE-fused) qkv projection wrote, instead of head-major copies.
static int attn_bwd_tc2_launch(const void* q, const void* k, const void* v, const void* qkv, const void* out, const void* dout, const float* lse,
                               void* dq, void* dk, void* dv, int B, int heads, int N, int hd, cudaStream_t st) {
  if (hd != HD || N > 256 || N < 1) return LNX_ERR_UNSUPPORTED;
  BwdParams p;
  p.B = B; p.heads = heads; p.N = N;
  p.n_t = (N + QT - 1) / QT;
  p.qkv4d = qkv ? 1 : 0;
  CUtensorMap tq, tk, tv, tdo;
  if (qkv) {
    CUtensorMap* tm[3] = {&tq, &tk, &tv};
    for (int which = 0; which < 3; ++which) {
      const long long dims[4] = {HD, heads, N, B};
      const long long strides[3] = {HD, 3LL * heads * HD, (long long)N * 3 * heads * HD};
      const int box[4] = {HD, 1, QT, 1};
      if (!make_tmap(tm[which], reinterpret_cast<const bf16*>(qkv) + (long long)which * heads * HD, 4, dims, strides, box)) return LNX_ERR_UNSUPPORTED;
    }
  } else if (!head_tmap(&tq, q, B * heads, N, QT) || !head_tmap(&tk, k, B * heads, N, QT) || !head_tmap(&tv, v, B * heads, N, QT)) {
    return LNX_ERR_UNSUPPORTED;
  }
  {  // dout [B][N][heads*64]: box {64, 128, 1} at column h*64
    const long long dims[3] = {(long long)heads * HD, N, B};
    const long long strides[2] = {(long long)heads * HD, (long long)N * heads * HD};
    const int box[3] = {HD, QT, 1};
    if (!make_tmap(&tdo, dout, 3, dims, strides, box)) return LNX_ERR_UNSUPPORTED;
  }
  const size_t smem = 1024 + 4 * (size_t)p.n_t * 16384 + 2 * 32768 + B_NBARS * 8 + 64;
  static int smem_set = 0;
  if ((int)smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    smem_set = (int)smem;
  }
  attn_bwd_tc2_kernel<<<B * heads, 320, smem, st>>>(tq, tk, tv, tdo, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dq, (bf16*)dk, (bf16*)dv, p);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

int lnx_attn_bwd_tc2(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse, void* dq, void* dk,
                     void* dv, int B, int heads, int N, int hd, cudaStream_t st) {
  return attn_bwd_tc2_launch(q, k, v, nullptr, out, dout, lse, dq, dk, dv, B, heads, N, hd, st);
}

int lnx_attn_qkv_bwd_tc2(const void* qkv, const void* out, const void* dout, const float* lse, void* dq, void* dk, void* dv, int B, int heads,
                         int N, int hd, cudaStream_t st) {
  return attn_bwd_tc2_launch(nullptr, nullptr, nullptr, qkv, out, dout, lse, dq, dk, dv, B, heads, N, hd, st);
}
