// fp32 CUDA-core attention (head_dim 64): the 1e-4 parity path, and the bf16-mode
// fallback for shapes the tcgen05 kernel does not take.  softmax(q k^T) v with q
// pre-scaled; fp32 online softmax; LSE saved for the backward.
//   fwd : thread per query row, K/V streamed through shared memory in 64-key tiles
//   bwd1: thread per query row  -> delta, dQ
//   bwd2: two threads per key   -> dK, dV   (no atomics anywhere)
#include "lnx_common.cuh"

using namespace lnx;

namespace {

constexpr int HD = 64;
constexpr int KT = 64;  // keys (or queries) per shared tile

template <typename T>
__device__ __forceinline__ void load_row64(const T* p, float* out) {
  if (sizeof(T) == 2) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 raw = reinterpret_cast<const uint4*>(p)[c];
      const bf16* h = reinterpret_cast<const bf16*>(&raw);
#pragma unroll
      for (int e = 0; e < 8; ++e) out[c * 8 + e] = __bfloat162float(h[e]);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4 a = reinterpret_cast<const float4*>(p)[c];
      out[c * 4] = a.x; out[c * 4 + 1] = a.y; out[c * 4 + 2] = a.z; out[c * 4 + 3] = a.w;
    }
  }
}

// cooperative: tile[r][d] = src[(row0 + r) * HD + d] as fp32, zero beyond N
template <typename T>
__device__ __forceinline__ void load_tile64(float* tile, const T* __restrict__ src, int row0, int N) {
  for (int i = threadIdx.x; i < KT * HD / 4; i += blockDim.x) {
    const int r = i / (HD / 4), d4 = (i % (HD / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) {
      const T* p = src + (long long)(row0 + r) * HD + d4;
      v = make_float4(to_f32(p[0]), to_f32(p[1]), to_f32(p[2]), to_f32(p[3]));
    }
    *reinterpret_cast<float4*>(tile + r * HD + d4) = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(64) attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                                      T* __restrict__ out, float* __restrict__ lse, int heads, int N) {
  __shared__ __align__(16) float Ks[KT * HD];
  __shared__ __align__(16) float Vs[KT * HD];
  const int bh = blockIdx.y;
  const int b = bh / heads, h = bh % heads;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < N;
  const T* qb = q + (long long)bh * N * HD;
  const T* kb = k + (long long)bh * N * HD;
  const T* vb = v + (long long)bh * N * HD;
  float qr[HD], o[HD];
  if (valid) load_row64(qb + (long long)i * HD, qr);
  else {
#pragma unroll
    for (int d = 0; d < HD; ++d) qr[d] = 0.f;
  }
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < N; j0 += KT) {
    __syncthreads();
    load_tile64(Ks, kb, j0, N);
    load_tile64(Vs, vb, j0, N);
    __syncthreads();
    const int jn = min(KT, N - j0);
    for (int jc = 0; jc < jn; jc += 8) {
      float sc[8];
      float cmax = -INFINITY;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float a = 0.f;
        if (jc + u < jn) {
          const float4* kr = reinterpret_cast<const float4*>(Ks + (jc + u) * HD);
#pragma unroll
          for (int d4 = 0; d4 < HD / 4; ++d4) {
            const float4 kk = kr[d4];
            a = fmaf(qr[d4 * 4], kk.x, a); a = fmaf(qr[d4 * 4 + 1], kk.y, a);
            a = fmaf(qr[d4 * 4 + 2], kk.z, a); a = fmaf(qr[d4 * 4 + 3], kk.w, a);
          }
        } else {
          a = -INFINITY;
        }
        sc[u] = a;
        cmax = fmaxf(cmax, a);
      }
      const float mn = fmaxf(m, cmax);
      const float corr = expf(m - mn);  // m = -inf on the first chunk -> 0
      l *= corr;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] *= corr;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (jc + u < jn) {
          const float pexp = expf(sc[u] - mn);
          l += pexp;
          const float4* vr = reinterpret_cast<const float4*>(Vs + (jc + u) * HD);
#pragma unroll
          for (int d4 = 0; d4 < HD / 4; ++d4) {
            const float4 vv = vr[d4];
            o[d4 * 4] = fmaf(pexp, vv.x, o[d4 * 4]); o[d4 * 4 + 1] = fmaf(pexp, vv.y, o[d4 * 4 + 1]);
            o[d4 * 4 + 2] = fmaf(pexp, vv.z, o[d4 * 4 + 2]); o[d4 * 4 + 3] = fmaf(pexp, vv.w, o[d4 * 4 + 3]);
          }
        }
      }
      m = mn;
    }
  }
  if (valid) {
    const float inv = 1.0f / l;
    T* dst = out + (((long long)b * N + i) * heads + h) * HD;  // [B, N, heads*HD]
#pragma unroll
    for (int d = 0; d < HD; ++d) dst[d] = from_f32<T>(o[d] * inv);
    lse[(long long)bh * N + i] = m + logf(l);
  }
}

// dQ and delta: thread per query
template <typename T>
__global__ void __launch_bounds__(64) attn_bwd_dq_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                                         const T* __restrict__ out, const T* __restrict__ dout,
                                                         const float* __restrict__ lse, T* __restrict__ dq, float* __restrict__ delta,
                                                         int heads, int N) {
  __shared__ __align__(16) float Ks[KT * HD];
  __shared__ __align__(16) float Vs[KT * HD];
  const int bh = blockIdx.y;
  const int b = bh / heads, h = bh % heads;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < N;
  const T* kb = k + (long long)bh * N * HD;
  const T* vb = v + (long long)bh * N * HD;
  float qr[HD], go[HD], acc[HD];
  float L = 0.f, dl = 0.f;
  if (valid) {
    load_row64(q + ((long long)bh * N + i) * HD, qr);
    const long long oidx = (((long long)b * N + i) * heads + h) * HD;
    load_row64(dout + oidx, go);
    float orow[HD];
    load_row64(out + oidx, orow);
#pragma unroll
    for (int d = 0; d < HD; ++d) dl = fmaf(go[d], orow[d], dl);
    L = lse[(long long)bh * N + i];
    delta[(long long)bh * N + i] = dl;
  } else {
#pragma unroll
    for (int d = 0; d < HD; ++d) qr[d] = 0.f, go[d] = 0.f;
  }
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  for (int j0 = 0; j0 < N; j0 += KT) {
    __syncthreads();
    load_tile64(Ks, kb, j0, N);
    load_tile64(Vs, vb, j0, N);
    __syncthreads();
    const int jn = min(KT, N - j0);
    for (int j = 0; j < jn; ++j) {
      const float4* kr = reinterpret_cast<const float4*>(Ks + j * HD);
      const float4* vr = reinterpret_cast<const float4*>(Vs + j * HD);
      float sdot = 0.f, dp = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < HD / 4; ++d4) {
        const float4 kk = kr[d4], vv = vr[d4];
        sdot = fmaf(qr[d4 * 4], kk.x, sdot); sdot = fmaf(qr[d4 * 4 + 1], kk.y, sdot);
        sdot = fmaf(qr[d4 * 4 + 2], kk.z, sdot); sdot = fmaf(qr[d4 * 4 + 3], kk.w, sdot);
        dp = fmaf(go[d4 * 4], vv.x, dp); dp = fmaf(go[d4 * 4 + 1], vv.y, dp);
        dp = fmaf(go[d4 * 4 + 2], vv.z, dp); dp = fmaf(go[d4 * 4 + 3], vv.w, dp);
      }
      const float pr = expf(sdot - L);
      const float ds = pr * (dp - dl);
#pragma unroll
      for (int d4 = 0; d4 < HD / 4; ++d4) {
        const float4 kk = kr[d4];
        acc[d4 * 4] = fmaf(ds, kk.x, acc[d4 * 4]); acc[d4 * 4 + 1] = fmaf(ds, kk.y, acc[d4 * 4 + 1]);
        acc[d4 * 4 + 2] = fmaf(ds, kk.z, acc[d4 * 4 + 2]); acc[d4 * 4 + 3] = fmaf(ds, kk.w, acc[d4 * 4 + 3]);
      }
    }
  }
  if (valid) {
    T* dst = dq + ((long long)bh * N + i) * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) dst[d] = from_f32<T>(acc[d]);
  }
}

// dK, dV: two threads per key (each owns 32 of the 64 dims); block = 128 threads = 64 keys
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                                           const T* __restrict__ dout, const float* __restrict__ lse,
                                                           const float* __restrict__ delta, T* __restrict__ dk, T* __restrict__ dv,
                                                           int heads, int N) {
  __shared__ __align__(16) float Qs[KT * HD];
  __shared__ __align__(16) float Gs[KT * HD];
  __shared__ float Ls[KT], Ds[KT];
  const int bh = blockIdx.y;
  const int b = bh / heads, h = bh % heads;
  const int j = blockIdx.x * 64 + (threadIdx.x >> 1);
  const int half = threadIdx.x & 1;
  const bool valid = j < N;
  float kr[32], vr[32], dkr[32], dvr[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    kr[d] = valid ? to_f32(k[((long long)bh * N + j) * HD + half * 32 + d]) : 0.f;
    vr[d] = valid ? to_f32(v[((long long)bh * N + j) * HD + half * 32 + d]) : 0.f;
    dkr[d] = 0.f;
    dvr[d] = 0.f;
  }
  const T* qb = q + (long long)bh * N * HD;
  for (int i0 = 0; i0 < N; i0 += KT) {
    __syncthreads();
    load_tile64(Qs, qb, i0, N);
    // dout rows live in [B, N, heads*HD]
    for (int t = threadIdx.x; t < KT * HD / 4; t += blockDim.x) {
      const int r = t / (HD / 4), d4 = (t % (HD / 4)) * 4;
      float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i0 + r < N) {
        const T* p = dout + (((long long)b * N + (i0 + r)) * heads + h) * HD + d4;
        g4 = make_float4(to_f32(p[0]), to_f32(p[1]), to_f32(p[2]), to_f32(p[3]));
      }
      *reinterpret_cast<float4*>(Gs + r * HD + d4) = g4;
    }
    for (int t = threadIdx.x; t < KT; t += blockDim.x) {
      Ls[t] = (i0 + t < N) ? lse[(long long)bh * N + i0 + t] : 0.f;
      Ds[t] = (i0 + t < N) ? delta[(long long)bh * N + i0 + t] : 0.f;
    }
    __syncthreads();
    const int in = min(KT, N - i0);
    for (int i = 0; i < in; ++i) {
      const float* qrow = Qs + i * HD + half * 32;
      const float* grow = Gs + i * HD + half * 32;
      float sdot = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        sdot = fmaf(qrow[d], kr[d], sdot);
        dp = fmaf(grow[d], vr[d], dp);
      }
      sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
      dp += __shfl_xor_sync(0xffffffffu, dp, 1);
      const float pr = expf(sdot - Ls[i]);
      const float ds = pr * (dp - Ds[i]);
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        dvr[d] = fmaf(pr, grow[d], dvr[d]);
        dkr[d] = fmaf(ds, qrow[d], dkr[d]);
      }
    }
  }
  if (valid) {
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      dk[((long long)bh * N + j) * HD + half * 32 + d] = from_f32<T>(dkr[d]);
      dv[((long long)bh * N + j) * HD + half * 32 + d] = from_f32<T>(dvr[d]);
    }
  }
}

}  // namespace

int lnx_attn_fwd_simt(const void* q, const void* k, const void* v, void* out, float* lse, int B, int heads, int N, int hd, int dtype,
                      cudaStream_t st) {
  if (hd != HD) return LNX_ERR_UNSUPPORTED;
  dim3 grid((N + 63) / 64, B * heads);
  if (dtype == LNX_F32)
    attn_fwd_kernel<float><<<grid, 64, 0, st>>>((const float*)q, (const float*)k, (const float*)v, (float*)out, lse, heads, N);
  else if (dtype == LNX_BF16)
    attn_fwd_kernel<bf16><<<grid, 64, 0, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)out, lse, heads, N);
  else
    return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

int lnx_attn_bwd_simt(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse, void* dq,
                      void* dk, void* dv, float* delta, int B, int heads, int N, int hd, int dtype, cudaStream_t st) {
  if (hd != HD) return LNX_ERR_UNSUPPORTED;
  dim3 grid((N + 63) / 64, B * heads);
  if (dtype == LNX_F32) {
    attn_bwd_dq_kernel<float><<<grid, 64, 0, st>>>((const float*)q, (const float*)k, (const float*)v, (const float*)out, (const float*)dout, lse, (float*)dq, delta, heads, N);
    attn_bwd_dkv_kernel<float><<<grid, 128, 0, st>>>((const float*)q, (const float*)k, (const float*)v, (const float*)dout, lse, delta, (float*)dk, (float*)dv, heads, N);
  } else if (dtype == LNX_BF16) {
    attn_bwd_dq_kernel<bf16><<<grid, 64, 0, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dq, delta, heads, N);
    attn_bwd_dkv_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)dout, lse, delta, (bf16*)dk, (bf16*)dv, heads, N);
  } else
    return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

int lnx_attn_fwd_tc(const void* q, const void* k, const void* v, void* out, float* lse, int B, int heads, int N, int hd, cudaStream_t st);
int lnx_attn_bwd_tc(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse, float* dq_f32,
                    void* dk, void* dv, int B, int heads, int N, int hd, cudaStream_t st);

int lnx_attn_fwd_tc2(const void* q, const void* k, const void* v, void* out, float* lse, int B, int heads, int N, int hd, cudaStream_t st);
int lnx_attn_fwd_long_tc2(const void* q, const void* k, const void* v, void* out, float* lse, int B, int heads, int N, int hd, cudaStream_t st);
int lnx_attn_bwd_tc2(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse, void* dq, void* dk,
                     void* dv, int B, int heads, int N, int hd, cudaStream_t st);

int lnx_attn_bias_fwd_tc2(const void* qkv, const float* bias, void* out, float* lse, int B, int heads, int N, int hd, float scale, cudaStream_t st);
int lnx_attn_qkv_bwd_tc2(const void* qkv, const void* out, const void* dout, const float* lse, void* dq, void* dk, void* dv, int B, int heads,
                         int N, int hd, cudaStream_t st);

// Attention straight from the [B, N, 3, heads, hd] matrix of the (RoPE-fused) qkv projection: no head-major copies of q / k / v.
// bf16, hd = 64, N <= 240 (the tcgen05 kernels); anything else is LNX_ERR_UNSUPPORTED and the caller keeps the split path.
extern "C" int lnx_attn_qkv_fwd(const void* qkv, void* out, float* lse, int B, int heads, int N, int hd, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(qkv && out && lse, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && heads > 0 && N > 0, LNX_ERR_SHAPE);
  if (dtype != LNX_BF16 || hd != 64) return LNX_ERR_UNSUPPORTED;
  return lnx_attn_bias_fwd_tc2(qkv, nullptr, out, lse, B, heads, N, hd, 1.0f, (cudaStream_t)s);
}

// dq / dk / dv: head-major [B, heads, N, hd] bf16 (lnx_rope_qk_bwd_scaled turns them into dqkv)
extern "C" int lnx_attn_qkv_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dq, void* dk, void* dv, int B,
                                int heads, int N, int hd, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(qkv && out && dout && lse && dq && dk && dv, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && heads > 0 && N > 0, LNX_ERR_SHAPE);
  if (dtype != LNX_BF16 || hd != 64 || N > 240) return LNX_ERR_UNSUPPORTED;
  return lnx_attn_qkv_bwd_tc2(qkv, out, dout, lse, dq, dk, dv, B, heads, N, hd, (cudaStream_t)s);
}

extern "C" int lnx_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int heads, int N, int hd,
                            int dtype, int force_simt, lnx_stream_t s) {
  LNX_REQUIRE(q && k && v && out && lse, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && heads > 0 && N > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(q) && lnx_aligned16(k) && lnx_aligned16(v) && lnx_aligned16(out), LNX_ERR_ALIGN);
  if (dtype == LNX_BF16 && !force_simt) {
    int r = lnx_attn_fwd_tc2(q, k, v, out, lse, B, heads, N, hd, (cudaStream_t)s);  // persistent, pipelined (N <= 240)
    if (r != LNX_ERR_UNSUPPORTED) return r;
    r = lnx_attn_fwd_long_tc2(q, k, v, out, lse, B, heads, N, hd, (cudaStream_t)s);  // streamed key tiles, online softmax
    if (r != LNX_ERR_UNSUPPORTED) return r;
    r = lnx_attn_fwd_tc(q, k, v, out, lse, B, heads, N, hd, (cudaStream_t)s);
    if (r != LNX_ERR_UNSUPPORTED) return r;
  }
  return lnx_attn_fwd_simt(q, k, v, out, lse, B, heads, N, hd, dtype, (cudaStream_t)s);
}

extern "C" int lnx_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse, void* dq,
                            void* dk, void* dv, float* delta_ws, int B, int heads, int N, int hd, int dtype, int force_simt,
                            lnx_stream_t s) {
  LNX_REQUIRE(q && k && v && out && dout && lse && dq && dk && dv && delta_ws, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && heads > 0 && N > 0, LNX_ERR_SHAPE);
  if (dtype == LNX_BF16 && !force_simt && hd == 64) {
    // N <= 256: one CTA per (batch, head), dQ accumulated in TMEM and written once as bf16
    const int r2 = lnx_attn_bwd_tc2(q, k, v, out, dout, lse, dq, dk, dv, B, heads, N, hd, (cudaStream_t)s);
    if (r2 != LNX_ERR_UNSUPPORTED) return r2;
    // longer sequences: dQ is accumulated in fp32 (two key tiles per head) in the workspace, then cast
    float* dq32 = delta_ws + (((long long)B * heads * N + 3) / 4) * 4;  // keep the accumulator 16-byte aligned
    const long long n = (long long)B * heads * N * hd;
    cudaError_t e = cudaMemsetAsync(dq32, 0, sizeof(float) * n, (cudaStream_t)s);
    if (e != cudaSuccess) return lnx_set_cuda_error(e);
    const int r = lnx_attn_bwd_tc(q, k, v, out, dout, lse, dq32, dk, dv, B, heads, N, hd, (cudaStream_t)s);
    if (r == LNX_OK) return lnx_cast_f32_to_bf16(dq32, dq, n, s);
    if (r != LNX_ERR_UNSUPPORTED) return r;
  }
  return lnx_attn_bwd_simt(q, k, v, out, dout, lse, dq, dk, dv, delta_ws, B, heads, N, hd, dtype, (cudaStream_t)s);
}
