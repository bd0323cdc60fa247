// 2-D "RoPE" of the reference = cos-scaling of (even, odd) q/k pairs of the image
// tokens.  Forward splits qkv into head-major q/k/v (the layout the
// attention kernels read with TMA / coalesced loads) and folds the softmax scale into q.
// Backward folds the same factors into dqkv and reduces dtheta for the learnable freqs.
#include "lnx_common.cuh"

using namespace lnx;

namespace {

template <typename T>
__device__ __forceinline__ void ld8(const T* p, float* x);
template <>
__device__ __forceinline__ void ld8<float>(const float* p, float* x) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
template <>
__device__ __forceinline__ void ld8<bf16>(const bf16* p, float* x) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  x[0] = __uint_as_float(r.x << 16); x[1] = __uint_as_float(r.x & 0xffff0000u);
  x[2] = __uint_as_float(r.y << 16); x[3] = __uint_as_float(r.y & 0xffff0000u);
  x[4] = __uint_as_float(r.z << 16); x[5] = __uint_as_float(r.z & 0xffff0000u);
  x[6] = __uint_as_float(r.w << 16); x[7] = __uint_as_float(r.w & 0xffff0000u);
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float* x);
template <>
__device__ __forceinline__ void st8<float>(float* p, const float* x) {
  *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
  *(reinterpret_cast<float4*>(p) + 1) = make_float4(x[4], x[5], x[6], x[7]);
}
template <>
__device__ __forceinline__ void st8<bf16>(bf16* p, const float* x) {
  uint4 raw;
  __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) hh[e] = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = raw;
}

__global__ void rope_table_kernel(const float* __restrict__ freqs, float* __restrict__ cos_out, float* __restrict__ sin_out, int H,
                                  int W, int heads, int half) {
  const int total = H * W * heads * half;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i % half;
    const int h = (i / half) % heads;
    const int n = i / (half * heads);
    const float tx = (float)(n % W), ty = (float)(n / W);
    const float th = tx * freqs[(0 * heads + h) * half + j] + ty * freqs[(1 * heads + h) * half + j];
    cos_out[i] = cosf(th);
    if (sin_out) sin_out[i] = sinf(th);
  }
}

// CTA = TOK * ITER consecutive tokens; thread = (token slot, which in {q,k,v}, head, 8-element chunk) and walks ITER tokens
// TOK apart: the index decomposition is done once, the only division is 32-bit, all ITER 16-byte row reads are issued before
// the first use; coalesced reads of the [B*N, 3*heads*hd] rows, 128-byte segments on the head-major side.
template <typename T, int TOK, int ITER>
__global__ void rope_qk_fwd_kernel(const T* __restrict__ qkv, const float* __restrict__ cos_tab, T* __restrict__ q, T* __restrict__ k,
                                   T* __restrict__ v, unsigned BN, int N, int heads, int hd, int n_extra, float q_scale) {
  const int chunks = hd / 8;
  const int per_tok = 3 * heads * chunks;
  const int t = threadIdx.x;
  if (t >= per_tok * TOK) return;
  const int tok = t / per_tok;
  const int w = t - tok * per_tok;
  const int ch = w % chunks;
  const int h = (w / chunks) % heads;
  const int which = w / (chunks * heads);
  const unsigned row0 = blockIdx.x * (unsigned)(TOK * ITER) + tok;  // = b * N + n
  const float sc = (which == 0) ? q_scale : 1.0f;
  T* const dst0 = (which == 0 ? q : which == 1 ? k : v) + (long long)h * N * hd + ch * 8;
  float x[ITER][8];
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const unsigned row = row0 + it * TOK;
    if (row < BN) ld8<T>(qkv + ((long long)row * per_tok + w) * 8, x[it]);
  }
  unsigned b = row0 / (unsigned)N;
  int n = (int)(row0 - b * (unsigned)N);
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    if (row0 + it * TOK < BN) {
      if (which < 2) {
        if (n >= n_extra) {
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(cos_tab + ((long long)(n - n_extra) * heads + h) * (hd / 2) + ch * 4));
          const float ct[4] = {c4.x * sc, c4.y * sc, c4.z * sc, c4.w * sc};
#pragma unroll
          for (int e = 0; e < 8; ++e) x[it][e] *= ct[e >> 1];
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) x[it][e] *= sc;
        }
      }
      st8<T>(dst0 + ((long long)b * heads * N + n) * hd, x[it]);
    }
    n += TOK;
    while (n >= N) {
      n -= N;
      ++b;
    }
  }
}

// grid: (N, batch chunks); thread = (which in {q,k,v}, head, pair-chunk of 4 pairs)
// ROT: qkv holds the ALREADY scaled q' = q cos s / k' = k cos (the qkv projection applied the factors in its epilogue), so
// g q s (-sin) = -g q' tan(theta).  bf16 keeps its 8 relative bits down to 1e-38, so q' = q cos s is as informative as q even where cos
// is tiny (the term there is the LARGEST, |sin| ~ 1: it must not be dropped); only an exactly vanishing cos loses it.
template <typename T, bool ROT>
__global__ void rope_qk_bwd_kernel(const T* __restrict__ dq, const T* __restrict__ dk, const T* __restrict__ dv,
                                   const T* __restrict__ qkv, const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                                   T* __restrict__ dqkv, float* __restrict__ dtheta, int B, int N, int heads, int hd, int n_extra,
                                   float q_scale, int b_per_block) {
  const int n = blockIdx.x;
  const int chunks = hd / 8;
  const int work = 3 * heads * chunks;
  const int b0 = blockIdx.y * b_per_block, b1 = min(B, b0 + b_per_block);
  for (int t = threadIdx.x; t < work; t += blockDim.x) {
    const int ch = t % chunks;
    const int h = (t / chunks) % heads;
    const int which = t / (chunks * heads);
    const bool img = n >= n_extra;
    float c[4] = {1.f, 1.f, 1.f, 1.f}, sn[4] = {0.f, 0.f, 0.f, 0.f};
    if (img && which < 2) {
      const long long o = ((long long)(n - n_extra) * heads + h) * (hd / 2) + ch * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) c[e] = cos_tab[o + e], sn[e] = sin_tab[o + e];
    }
    const float sc = (which == 0) ? q_scale : 1.0f;
    float mt[4];  // multiplier of g * x in dtheta
#pragma unroll
    for (int e = 0; e < 4; ++e) mt[e] = ROT ? (fabsf(c[e]) > 1e-30f ? -sn[e] / c[e] : 0.f) : -sn[e] * sc;
    float dth[4] = {0.f, 0.f, 0.f, 0.f};
    const T* gsrc = (which == 0 ? dq : which == 1 ? dk : dv);
#pragma unroll 4
    for (int b = b0; b < b1; ++b) {
      const T* gp = gsrc + ((((long long)b * heads + h) * N + n) * hd) + ch * 8;
      const long long o = ((((long long)b * N + n) * 3 + which) * heads + h) * hd + ch * 8;
      float g[8], x[8];
      ld8<T>(gp, g);
      if (which < 2 && img) {
        ld8<T>(qkv + o, x);
#pragma unroll
        for (int e = 0; e < 8; ++e) dth[e >> 1] += g[e] * x[e] * mt[e >> 1];
      }
      if (which < 2) {
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] *= c[e >> 1] * sc;
      }
      st8<T>(dqkv + o, g);
    }
    if (which < 2 && img) {
      float* dst = dtheta + ((long long)(n - n_extra) * heads + h) * (hd / 2) + ch * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(dst + e, dth[e]);
    }
  }
}

__global__ void rope_freq_grad_kernel(const float* __restrict__ dtheta, float* __restrict__ dfreqs, int H, int W, int heads, int half) {
  // one warp per (h, j): lanes stride over the grid positions, shuffle reduction
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= heads * half) return;
  float gx = 0.f, gy = 0.f;
  for (int n = lane; n < H * W; n += 32) {
    const float d = dtheta[(long long)n * heads * half + i];
    gx += (float)(n % W) * d;
    gy += (float)(n / W) * d;
  }
  gx = warp_sum(gx);
  gy = warp_sum(gy);
  if (lane == 0) {
    atomicAdd(dfreqs + i, gx);
    atomicAdd(dfreqs + heads * half + i, gy);
  }
}

}  // namespace

extern "C" int lnx_rope_table(const float* freqs, float* cos_out, float* sin_out, int H, int W, int heads, int half, lnx_stream_t s) {
  LNX_REQUIRE(freqs && cos_out, LNX_ERR_NULL);
  LNX_REQUIRE(H > 0 && W > 0 && heads > 0 && half > 0, LNX_ERR_SHAPE);
  const int total = H * W * heads * half;
  rope_table_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)s>>>(freqs, cos_out, sin_out, H, W, heads, half);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_rope_qk_fwd(const void* qkv, const float* cos_tab, void* q, void* k, void* v, int B, int N, int heads, int hd,
                               int n_extra, float q_scale, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(qkv && cos_tab && q && k && v, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && N > n_extra && heads > 0 && hd % 8 == 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(qkv) && lnx_aligned16(q) && lnx_aligned16(k) && lnx_aligned16(v), LNX_ERR_ALIGN);
  const int per_tok = 3 * heads * (hd / 8);
  LNX_REQUIRE(per_tok <= 1024 && hd % 8 == 0 && (hd / 2) % 4 == 0, LNX_ERR_UNSUPPORTED);
  const long long BN = (long long)B * N;
  cudaStream_t st = (cudaStream_t)s;
  LNX_REQUIRE(BN < (1LL << 31), LNX_ERR_SHAPE);
#define LNX_ROPE_F(T, TOK)                                                                                                         \
  rope_qk_fwd_kernel<T, TOK, 4><<<(unsigned)((BN + TOK * 4 - 1) / (TOK * 4)), ((per_tok * TOK + 31) / 32) * 32, 0, st>>>(            \
      (const T*)qkv, cos_tab, (T*)q, (T*)k, (T*)v, (unsigned)BN, N, heads, hd, n_extra, q_scale)
  if (dtype == LNX_F32) {
    if (per_tok * 4 <= 1024) LNX_ROPE_F(float, 4); else LNX_ROPE_F(float, 1);
  } else if (dtype == LNX_BF16) {
    if (per_tok * 4 <= 1024) LNX_ROPE_F(bf16, 4); else LNX_ROPE_F(bf16, 1);
  } else
    return LNX_ERR_DTYPE;
#undef LNX_ROPE_F
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

static int rope_qk_bwd_launch(bool rot, const void* dq, const void* dk, const void* dv, const void* qkv, const float* cos_tab, const float* sin_tab,
                               void* dqkv, float* dtheta, int B, int N, int heads, int hd, int n_extra, float q_scale, int dtype,
                               lnx_stream_t s) {
  LNX_REQUIRE(dq && dk && dv && qkv && cos_tab && sin_tab && dqkv && dtheta, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && N > n_extra && heads > 0 && hd % 8 == 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(dq) && lnx_aligned16(dk) && lnx_aligned16(dv) && lnx_aligned16(qkv) && lnx_aligned16(dqkv), LNX_ERR_ALIGN);
  const int work = 3 * heads * (hd / 8);
  const int threads = min(256, ((work + 31) / 32) * 32);
  // batch chunk per CTA: the grid (N x chunks) should fill whole waves of resident CTAs (6 per SM at 64 registers x 160
  // threads) - the old fixed N x 6 grid ran as 1.35 waves; among chunkings with >= 2 waves take the one wasting the least
  // of its last wave, preferring larger chunks (fewer dtheta atomics)
  const double slots = 6.0 * kNumSMs;
  int bpb = B;
  double best = -1.0;
  for (int c = B; c >= 1; --c) {
    const int chunks_c = (B + c - 1) / c;
    const double waves = (double)N * chunks_c / slots;
    if (waves < 2.0 && c > 1) continue;
    const double eff = waves / ceil(waves);
    if (eff > best + 0.02) best = eff, bpb = c;
    if (waves > 12.0) break;
  }
  const int by = (B + bpb - 1) / bpb;
  dim3 grid(N, by);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_F32)
    (rot ? rope_qk_bwd_kernel<float, true> : rope_qk_bwd_kernel<float, false>)<<<grid, threads, 0, st>>>((const float*)dq, (const float*)dk, (const float*)dv, (const float*)qkv, cos_tab, sin_tab, (float*)dqkv, dtheta, B, N, heads, hd, n_extra, q_scale, bpb);
  else if (dtype == LNX_BF16)
    (rot ? rope_qk_bwd_kernel<bf16, true> : rope_qk_bwd_kernel<bf16, false>)<<<grid, threads, 0, st>>>((const bf16*)dq, (const bf16*)dk, (const bf16*)dv, (const bf16*)qkv, cos_tab, sin_tab, (bf16*)dqkv, dtheta, B, N, heads, hd, n_extra, q_scale, bpb);
  else
    return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_rope_qk_bwd(const void* dq, const void* dk, const void* dv, const void* qkv, const float* cos_tab, const float* sin_tab,
                               void* dqkv, float* dtheta, int B, int N, int heads, int hd, int n_extra, float q_scale, int dtype,
                               lnx_stream_t s) {
  return rope_qk_bwd_launch(false, dq, dk, dv, qkv, cos_tab, sin_tab, dqkv, dtheta, B, N, heads, hd, n_extra, q_scale, dtype, s);
}

// Same, for the fused projection (lnx_qkv_rope_gemm): qkv_scaled holds q cos s / k cos / v, the un-scaled q / k were never stored.
extern "C" int lnx_rope_qk_bwd_scaled(const void* dq, const void* dk, const void* dv, const void* qkv_scaled, const float* cos_tab,
                                      const float* sin_tab, void* dqkv, float* dtheta, int B, int N, int heads, int hd, int n_extra,
                                      float q_scale, int dtype, lnx_stream_t s) {
  return rope_qk_bwd_launch(true, dq, dk, dv, qkv_scaled, cos_tab, sin_tab, dqkv, dtheta, B, N, heads, hd, n_extra, q_scale, dtype, s);
}

extern "C" int lnx_rope_freq_grad(const float* dtheta, float* dfreqs, int H, int W, int heads, int half, lnx_stream_t s) {
  LNX_REQUIRE(dtheta && dfreqs, LNX_ERR_NULL);
  const int n = heads * half;
  rope_freq_grad_kernel<<<(n * 32 + 127) / 128, 128, 0, (cudaStream_t)s>>>(dtheta, dfreqs, H, W, heads, half);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
