// Attention on the 5th-gen tensor cores for the mFormer token counts (N <= 256, head_dim 64):
// the whole key/value set of one (batch, head) fits in shared memory, so softmax is a
// single exact pass over an S tile held in TMEM (no online rescaling needed).
//
//   forward  CTA = (128-query tile, batch*head):  TMA Q,K,V -> S = Q K^T (tcgen05, TMEM)
//            -> 128 softmax threads, one S row each (tcgen05.ld), P (bf16) to swizzled smem
//            -> O = P V (tcgen05; V consumed MN-major straight from its [key][d] tile) -> out, LSE
//   backward CTA = (128-key tile, batch*head), loops over query tiles:
//            S = Q K^T, dP = dO V^T -> P = exp(S - lse), dS = P (dP - delta) (bf16, smem)
//            -> dV += P^T dO, dK += dS^T Q (MN-major A straight from the P/dS tiles), dQ = dS K
//            dQ leaves through fp32 vector atomics (two key tiles per head), dK/dV stay in TMEM
//            until the last query tile.
// q is pre-scaled (softmax scale and cos factors folded in by the RoPE kernel).
#include "lnx_tc_common.cuh"

using namespace lnx;
using namespace lnx_tc;

namespace {

constexpr int HD = 64;
constexpr int QT = 128;    // rows per M tile
constexpr int KMAX = 256;  // max keys handled by the forward tile
constexpr float LOG2E = 1.4426950408889634f;

struct AttnParams {
  int B, heads, N, nkp;  // nkp = keys rounded up to 16
};

__device__ __forceinline__ void store_bf16x16_sw(unsigned char* tile, int r, int col0, const float* v) {
  // 16 consecutive columns starting at col0 (multiple of 16) of row r, in a [rows][64]-per-16KB-block swizzled tile
  const int kb = col0 >> 6;
  const int j = (col0 & 63) >> 3;
  unsigned char* blk = tile + kb * (QT * 128);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint4 raw;
    __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) p2[e] = __floats2bfloat162_rn(v[h * 8 + 2 * e], v[h * 8 + 2 * e + 1]);
    *reinterpret_cast<uint4*>(blk + sw128_chunk(r, j + h)) = raw;
  }
}

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(160) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                          const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ out,
                                                          float* __restrict__ lse, const AttnParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* sQ = base;               // 16 KB; P block 0 after S is done
  unsigned char* sK = base + 16384;       // 32 KB; P blocks 1, 2
  unsigned char* sV = base + 65536;       // 32 KB   (P block 3 lives at base + 48 KB)
  unsigned char* sP = base;               // 4 x 16 KB, contiguous
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 98304);
  uint64_t* bar_load = bars;
  uint64_t* bar_s = bars + 1;
  uint64_t* bar_p = bars + 2;
  uint64_t* bar_o = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int m0 = blockIdx.x * QT;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_load, 16384 + 32768 + 32768);
      tma_load_3d(sQ, &tmQ, bar_load, 0, m0, bh);
      tma_load_3d(sK, &tmK, bar_load, 0, 0, bh);
      tma_load_3d(sV, &tmV, bar_load, 0, 0, bh);
      mbar_wait(bar_load, 0);
      tcgen05_fence_after();
      const uint32_t idesc_s = make_idesc_bf16(QT, p.nkp, 0, 0);
      const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        umma_bf16(tmem, make_smem_desc(aq + k * 32, 0, 1024), make_smem_desc(ak + k * 32, 0, 1024), idesc_s, k > 0);
      umma_commit(bar_s);
      mbar_wait(bar_p, 0);
      tcgen05_fence_after();
      const uint32_t idesc_o = make_idesc_bf16(QT, HD, 0, 1);
      const uint32_t ap = smem_u32(sP), av = smem_u32(sV);
      for (int k = 0; k < p.nkp / 16; ++k)
        umma_bf16(tmem, make_smem_desc(ap + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024), make_smem_desc(av + k * 2048, 0, 1024),
                  idesc_o, k > 0);
      umma_commit(bar_o);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int m = m0 + r;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    mbar_wait(bar_s, 0);
    tcgen05_fence_after();
    // pass 1: row max over the valid keys
    float mx = -INFINITY;
    for (int c = 0; c < p.nkp; c += 16) {
      float v[16];
      tmem_ld16(trow + c, v);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c + i < p.N) mx = fmaxf(mx, v[i]);
    }
    const float mxl = mx * LOG2E;
    // pass 2: P = exp(S - max) -> bf16 smem (Q/K tiles are dead once S is complete), row sum in fp32
    float sum = 0.f;
    for (int c = 0; c < p.nkp; c += 16) {
      float v[16];
      tmem_ld16(trow + c, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float e = (c + i < p.N) ? exp2f(fmaf(v[i], LOG2E, -mxl)) : 0.f;
        sum += e;
        v[i] = e;
      }
      store_bf16x16_sw(sP, r, c, v);
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    mbar_arrive(bar_p);
    mbar_wait(bar_o, 0);
    tcgen05_fence_after();
    const float inv = 1.0f / sum;
    const int b = bh / p.heads, h = bh % p.heads;
#pragma unroll
    for (int c = 0; c < HD; c += 16) {
      float v[16];
      __syncwarp();
      tmem_ld16(trow + c, v);
      if (m < p.N) {
        bf16* dst = out + (((long long)b * p.N + m) * p.heads + h) * HD + c;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint4 raw;
          __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
          for (int e = 0; e < 4; ++e) p2[e] = __floats2bfloat162_rn(v[hh * 8 + 2 * e] * inv, v[hh * 8 + 2 * e + 1] * inv);
          *reinterpret_cast<uint4*>(dst + hh * 8) = raw;
        }
      }
    }
    if (m < p.N) lse[(long long)bh * p.N + m] = mx + logf(sum);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

// ------------------------------------------------------------------ backward
// TMEM columns: S [0,128) | dP [128,256) | dV [256,320) | dK [320,384) | dQ [384,448)
__global__ void __launch_bounds__(160) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                                                          const bf16* __restrict__ o, const bf16* __restrict__ dout,
                                                          const float* __restrict__ lse, float* __restrict__ dq, bf16* __restrict__ dk,
                                                          bf16* __restrict__ dv, const AttnParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* sK = base;             // [128 keys][128 B]
  unsigned char* sV = base + 16384;
  unsigned char* sQ = base + 32768;     // [128 queries][128 B]
  unsigned char* sdO = base + 49152;
  unsigned char* sP = base + 65536;     // [128 queries][128 keys] = 2 blocks x 16 KB
  unsigned char* sdS = base + 98304;    // same
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 131072);
  uint64_t* bar_kv = bars;
  uint64_t* bar_q = bars + 1;    // Q_i, dO_i landed
  uint64_t* bar_sdp = bars + 2;  // S, dP in TMEM
  uint64_t* bar_pds = bars + 3;  // P, dS in smem (128 arrivals)
  uint64_t* bar_acc = bars + 4;  // dV, dK, dQ MMAs done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.heads, h = bh % p.heads;
  const int k0 = blockIdx.x * QT;  // first key of this CTA
  const int n_qt = (p.N + QT - 1) / QT;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmdO);
    mbar_init(bar_kv, 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_sdp, 1);
    mbar_init(bar_pds, 128);
    mbar_init(bar_acc, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_kv, 2 * 16384);
      tma_load_3d(sK, &tmK, bar_kv, 0, k0, bh);
      tma_load_3d(sV, &tmV, bar_kv, 0, k0, bh);
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), aO = smem_u32(sdO), aP = smem_u32(sP), aS = smem_u32(sdS);
      const uint32_t id_s = make_idesc_bf16(QT, QT, 0, 0);     // S, dP: [128 q] x [128 keys], K = 64
      const uint32_t id_kv = make_idesc_bf16(QT, HD, 1, 1);    // dV, dK: A = P^T / dS^T (MN-major), B = dO / Q (MN-major)
      const uint32_t id_q = make_idesc_bf16(QT, HD, 0, 1);     // dQ: A = dS (K-major), B = K (MN-major)
      for (int i = 0; i < n_qt; ++i) {
        const uint32_t ph = (uint32_t)i & 1u;
        if (i > 0) mbar_wait(bar_acc, ph ^ 1u);  // previous tile's MMAs no longer read sQ / sdO / sP / sdS
        mbar_expect_tx(bar_q, 2 * 16384);
        tma_load_3d(sQ, &tmQ, bar_q, 0, i * QT, bh);
        tma_load_3d(sdO, &tmdO, bar_q, h * HD, i * QT, b);  // dout is [B][N][heads*64]
        if (i == 0) mbar_wait(bar_kv, 0);
        mbar_wait(bar_q, ph);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem + 0, make_smem_desc(aQ + k * 32, 0, 1024), make_smem_desc(aK + k * 32, 0, 1024), id_s, k > 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16(tmem + 128, make_smem_desc(aO + k * 32, 0, 1024), make_smem_desc(aV + k * 32, 0, 1024), id_s, k > 0);
        umma_commit(bar_sdp);
        mbar_wait(bar_pds, ph);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < QT / 16; ++k) {  // reduction over the 128 queries of this tile
          const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
          umma_bf16(tmem + 256, make_smem_desc(aP + k * 2048, 16384, 1024), make_smem_desc(aO + k * 2048, 0, 1024), id_kv, acc);
          umma_bf16(tmem + 320, make_smem_desc(aS + k * 2048, 16384, 1024), make_smem_desc(aQ + k * 2048, 0, 1024), id_kv, acc);
        }
#pragma unroll
        for (int k = 0; k < QT / 16; ++k)  // reduction over the 128 keys of this CTA
          umma_bf16(tmem + 384, make_smem_desc(aS + (k >> 2) * 16384 + (k & 3) * 32, 0, 1024), make_smem_desc(aK + k * 2048, 0, 1024),
                    id_q, k > 0);
        umma_commit(bar_acc);
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    for (int i = 0; i < n_qt; ++i) {
      const uint32_t ph = (uint32_t)i & 1u;
      const int m = i * QT + r;  // this thread's query
      float L = 0.f, delta = 0.f;
      if (m < p.N) {
        L = lse[(long long)bh * p.N + m] * LOG2E;
        const long long oidx = (((long long)b * p.N + m) * p.heads + h) * HD;
#pragma unroll
        for (int c = 0; c < HD; c += 8) {
          const uint4 ra = *reinterpret_cast<const uint4*>(o + oidx + c);
          const uint4 rb = *reinterpret_cast<const uint4*>(dout + oidx + c);
          const bf16* pa = reinterpret_cast<const bf16*>(&ra);
          const bf16* pb = reinterpret_cast<const bf16*>(&rb);
#pragma unroll
          for (int e = 0; e < 8; ++e) delta = fmaf(__bfloat162float(pa[e]), __bfloat162float(pb[e]), delta);
        }
      }
      mbar_wait(bar_sdp, ph);
      tcgen05_fence_after();
      for (int c = 0; c < QT; c += 16) {
        float s[16], dp[16];
        __syncwarp();
        tmem_ld16(trow + c, s);
        tmem_ld16(trow + 128 + c, dp);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const bool ok = (m < p.N) && (k0 + c + e < p.N);
          const float pr = ok ? exp2f(fmaf(s[e], LOG2E, -L)) : 0.f;
          s[e] = pr;
          dp[e] = pr * (dp[e] - delta);
        }
        store_bf16x16_sw(sP, r, c, s);
        store_bf16x16_sw(sdS, r, c, dp);
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive(bar_pds);
      mbar_wait(bar_acc, ph);
      tcgen05_fence_after();
#pragma unroll
      for (int c = 0; c < HD; c += 16) {
        float v[16];
        __syncwarp();
        tmem_ld16(trow + 384 + c, v);
        if (m < p.N) {
          float* dst = dq + ((long long)bh * p.N + m) * HD + c;
#pragma unroll
          for (int e = 0; e < 16; e += 4) atomicAdd(reinterpret_cast<float4*>(dst + e), make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]));
        }
      }
      tcgen05_fence_before();  // orders these TMEM reads before the next tile's MMAs (via bar_pds of the next iteration)
    }
    // dV, dK of this key tile (row r = key k0 + r)
    const int key = k0 + r;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      bf16* dst_base = (t == 0 ? dv : dk);
#pragma unroll
      for (int c = 0; c < HD; c += 16) {
        float v[16];
        __syncwarp();
        tmem_ld16(trow + 256 + t * 64 + c, v);
        if (key < p.N) {
          bf16* dst = dst_base + ((long long)bh * p.N + key) * HD + c;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint4 raw;
            __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
            for (int e = 0; e < 4; ++e) p2[e] = __floats2bfloat162_rn(v[hh * 8 + 2 * e], v[hh * 8 + 2 * e + 1]);
            *reinterpret_cast<uint4*>(dst + hh * 8) = raw;
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
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

}  // namespace

int lnx_attn_fwd_tc(const void* q, const void* k, const void* v, void* out, float* lse, int B, int heads, int N, int hd, cudaStream_t st) {
  if (hd != HD || N > KMAX || N < 1) return LNX_ERR_UNSUPPORTED;
  CUtensorMap tq, tk, tv;
  if (!head_tmap(&tq, q, B * heads, N, QT) || !head_tmap(&tk, k, B * heads, N, KMAX) || !head_tmap(&tv, v, B * heads, N, KMAX))
    return LNX_ERR_UNSUPPORTED;
  AttnParams p{B, heads, N, (N + 15) / 16 * 16};
  const size_t smem = 98304 + 1024 + 64;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    attr = true;
  }
  dim3 grid((N + QT - 1) / QT, B * heads);
  attn_fwd_tc_kernel<<<grid, 160, smem, st>>>(tq, tk, tv, (bf16*)out, lse, p);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

// dq_f32 must be zero-filled by the caller: [B*heads, N, 64] float32 (atomically accumulated)
int lnx_attn_bwd_tc(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse, float* dq_f32,
                    void* dk, void* dv, int B, int heads, int N, int hd, cudaStream_t st) {
  if (hd != HD || N < 1) return LNX_ERR_UNSUPPORTED;
  CUtensorMap tq, tk, tv, tdo;
  if (!head_tmap(&tq, q, B * heads, N, QT) || !head_tmap(&tk, k, B * heads, N, QT) || !head_tmap(&tv, v, B * heads, N, QT))
    return LNX_ERR_UNSUPPORTED;
  {  // dout [B][N][heads*64]: box {64, 128, 1} at column h*64
    const long long dims[3] = {(long long)heads * HD, N, B};
    const long long strides[2] = {(long long)heads * HD, (long long)N * heads * HD};
    const int box[3] = {HD, QT, 1};
    if (!make_tmap(&tdo, dout, 3, dims, strides, box)) return LNX_ERR_UNSUPPORTED;
  }
  AttnParams p{B, heads, N, (N + 15) / 16 * 16};
  const size_t smem = 131072 + 1024 + 64;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    attr = true;
  }
  dim3 grid((N + QT - 1) / QT, B * heads);
  attn_bwd_tc_kernel<<<grid, 160, smem, st>>>(tq, tk, tv, tdo, (const bf16*)out, (const bf16*)dout, lse, dq_f32, (bf16*)dk, (bf16*)dv, p);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
