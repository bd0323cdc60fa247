// Layout / gather kernels (all HBM-bound, 16-byte vectorised where the shape allows)
// and the library-wide error plumbing.
#include <stdio.h>
#include <string.h>

#include "lnx_common.cuh"

using namespace lnx;

// ---------------------------------------------------------------- errors
static thread_local char g_cuda_err[256] = "";

int lnx_set_cuda_error(cudaError_t e) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "CUDA error: %s", cudaGetErrorString(e));
  return LNX_ERR_CUDA;
}

extern "C" int lnx_version(void) { return 100; }

extern "C" const char* lnx_strerror(int code) {
  switch (code) {
    case LNX_OK: return "ok";
    case LNX_ERR_SHAPE: return "bad shape";
    case LNX_ERR_DTYPE: return "unsupported dtype";
    case LNX_ERR_ALIGN: return "misaligned pointer or pitch";
    case LNX_ERR_CUDA: return g_cuda_err[0] ? g_cuda_err : "CUDA error";
    case LNX_ERR_UNSUPPORTED: return "unsupported configuration";
    case LNX_ERR_NULL: return "null pointer";
    default: return "unknown error";
  }
}

// ---------------------------------------------------------------- patchify
// Thread = (output row, 4-column group): 16 groups per row when Kpad = 64.  A group is one (c, kh) run of p = 4
// pixels = one aligned 16-byte load from the NCHW image; the 16 threads of a row write one full 128-byte line.
template <typename T>
__global__ void patchify4_kernel(const float* __restrict__ x, T* __restrict__ out, int B, int Cin, int H, int W, int Kpad) {
  const int Ho = H / 4, Wo = W / 4;
  const int gpr = Kpad / 4;  // groups per output row
  const int K4 = Cin * 4;    // valid groups: (c, kh)
  const long long total = (long long)B * Ho * Wo * gpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int gidx = (int)(i % gpr);
    const long long row = i / gpr;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gidx < K4) {
      const int kh = gidx & 3, c = gidx >> 2;
      const int ow = (int)(row % Wo);
      const int oh = (int)((row / Wo) % Ho);
      const int b = (int)(row / ((long long)Wo * Ho));
      v = *reinterpret_cast<const float4*>(x + (((long long)b * Cin + c) * H + (oh * 4 + kh)) * W + ow * 4);
    }
    T* dst = out + row * Kpad + gidx * 4;
    if (sizeof(T) == 2) {
      __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), bq = __floats2bfloat162_rn(v.z, v.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&bq);
      *reinterpret_cast<uint2*>(dst) = o;
    } else {
      *reinterpret_cast<float4*>(dst) = v;
    }
  }
}

// generic fallback: one thread per output element
template <typename T>
__global__ void patchify_kernel(const float* __restrict__ x, T* __restrict__ out, int B, int Cin, int H, int W, int p, int Kpad) {
  const int Ho = H / p, Wo = W / p;
  const int K = Cin * p * p;
  const long long total = (long long)B * Ho * Wo * Kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % Kpad);
    const long long row = i / Kpad;
    float v = 0.f;
    if (kk < K) {
      const int kw = kk % p, kh = (kk / p) % p, c = kk / (p * p);
      const int ow = (int)(row % Wo);
      const int oh = (int)((row / Wo) % Ho);
      const int b = (int)(row / ((long long)Wo * Ho));
      v = x[(((long long)b * Cin + c) * H + (oh * p + kh)) * W + (ow * p + kw)];
    }
    out[i] = from_f32<T>(v);
  }
}

extern "C" int lnx_patchify_nchw(const float* x, void* out, int B, int Cin, int H, int W, int p, int Kpad, int out_dtype, lnx_stream_t s) {
  LNX_REQUIRE(x && out, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && Cin > 0 && p > 0 && H % p == 0 && W % p == 0 && Kpad >= Cin * p * p, LNX_ERR_SHAPE);
  const int threads = 256;
  cudaStream_t st = (cudaStream_t)s;
  if (p == 4 && Kpad % 4 == 0 && W % 4 == 0 && lnx_aligned16(x) && lnx_aligned16(out) && (out_dtype == LNX_F32 || out_dtype == LNX_BF16)) {
    const long long total = (long long)B * (H / 4) * (W / 4) * (Kpad / 4);
    const int blocks = (int)min((long long)kNumSMs * 16, (total + threads - 1) / threads);
    if (out_dtype == LNX_F32) patchify4_kernel<float><<<blocks, threads, 0, st>>>(x, (float*)out, B, Cin, H, W, Kpad);
    else patchify4_kernel<bf16><<<blocks, threads, 0, st>>>(x, (bf16*)out, B, Cin, H, W, Kpad);
    LNX_CHECK_LAUNCH();
    return LNX_OK;
  }
  const long long total = (long long)B * (H / p) * (W / p) * Kpad;
  const int blocks = (int)min((long long)kNumSMs * 16, (total + threads - 1) / threads);
  if (out_dtype == LNX_F32)
    patchify_kernel<float><<<blocks, threads, 0, st>>>(x, (float*)out, B, Cin, H, W, p, Kpad);
  else if (out_dtype == LNX_BF16)
    patchify_kernel<bf16><<<blocks, threads, 0, st>>>(x, (bf16*)out, B, Cin, H, W, p, Kpad);
  else
    return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

// ---------------------------------------------------------------- space to depth
// 16-byte chunks; chunk index over the [rows, 4C] side
template <typename T, bool INVERSE>
__global__ void s2d_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int H, int W, int C) {
  constexpr int V = Vec16<T>::N;
  const int Cv = C / V;
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * 4 * Cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % Cv);
    long long r = i / Cv;
    const int kw = (int)(r % 2); r /= 2;
    const int kh = (int)(r % 2); r /= 2;
    const int ow = (int)(r % Wo); r /= Wo;
    const int oh = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const long long nhwc = ((((long long)b * H + (2 * oh + kh)) * W) + (2 * ow + kw)) * C + (long long)cv * V;
    const long long s2d = i * V;
    if (!INVERSE)
      st16(out + s2d, ld16(in + nhwc));
    else
      st16(out + nhwc, ld16(in + s2d));
  }
}

extern "C" int lnx_space_to_depth(const void* x, void* out, int B, int H, int W, int C, int inverse, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(x && out, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && H % 2 == 0 && W % 2 == 0 && C > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(x) && lnx_aligned16(out), LNX_ERR_ALIGN);
  const int V = dtype == LNX_F32 ? 4 : 8;
  LNX_REQUIRE(C % V == 0, LNX_ERR_SHAPE);
  const long long total = (long long)B * H * W * (C / V);
  const int threads = 256;
  const int blocks = (int)min((long long)kNumSMs * 16, (total + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_F32) {
    if (inverse) s2d_kernel<float, true><<<blocks, threads, 0, st>>>((const float*)x, (float*)out, B, H, W, C);
    else s2d_kernel<float, false><<<blocks, threads, 0, st>>>((const float*)x, (float*)out, B, H, W, C);
  } else if (dtype == LNX_BF16) {
    if (inverse) s2d_kernel<bf16, true><<<blocks, threads, 0, st>>>((const bf16*)x, (bf16*)out, B, H, W, C);
    else s2d_kernel<bf16, false><<<blocks, threads, 0, st>>>((const bf16*)x, (bf16*)out, B, H, W, C);
  } else
    return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

// ---------------------------------------------------------------- token assemble / split
template <typename T>
__global__ void tokens_assemble_kernel(const T* __restrict__ cls, long long cls_stride, const T* __restrict__ extras,
                                       const T* __restrict__ patches, T* __restrict__ tokens, int B, int n_meta, int n_patch, int D) {
  constexpr int V = Vec16<T>::N;
  const int Dv = D / V;
  const int N = 1 + n_meta + n_patch;
  const long long total = (long long)B * N * Dv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int dv = (int)(i % Dv);
    const long long r = i / Dv;
    const int n = (int)(r % N);
    const long long b = r / N;
    Vec16<T> v;
    if (n == 0) {
      v = ld16(cls + b * cls_stride + (long long)dv * V);
    } else if (n <= n_meta) {
      if (extras) v = ld16(extras + ((b * n_meta + (n - 1)) * D) + (long long)dv * V);
      else {
#pragma unroll
        for (int j = 0; j < V; ++j) v.set(j, 0.f);
      }
    } else {
      v = ld16(patches + ((b * n_patch + (n - 1 - n_meta)) * D) + (long long)dv * V);
    }
    st16(tokens + i * V, v);
  }
}

template <typename T>
__global__ void tokens_split_kernel(const T* __restrict__ tokens, T* __restrict__ cls_out, T* __restrict__ extras_out,
                                    T* __restrict__ patches_out, int B, int n_meta, int n_patch, int D) {
  constexpr int V = Vec16<T>::N;
  const int Dv = D / V;
  const int N = 1 + n_meta + n_patch;
  const long long total = (long long)B * N * Dv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int dv = (int)(i % Dv);
    const long long r = i / Dv;
    const int n = (int)(r % N);
    const long long b = r / N;
    T* dst = nullptr;
    if (n == 0) {
      if (cls_out) dst = cls_out + b * D;
    } else if (n <= n_meta) {
      if (extras_out) dst = extras_out + (b * n_meta + (n - 1)) * D;
    } else {
      if (patches_out) dst = patches_out + (b * n_patch + (n - 1 - n_meta)) * D;
    }
    if (dst) st16(dst + (long long)dv * V, ld16(tokens + i * V));
  }
}

extern "C" int lnx_tokens_assemble(const void* cls, int64_t cls_stride, const void* extras, const void* patches, void* tokens,
                                   int B, int n_meta, int n_patch, int D, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(cls && patches && tokens, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && n_meta >= 0 && n_patch > 0 && D > 0, LNX_ERR_SHAPE);
  const int V = dtype == LNX_F32 ? 4 : 8;
  LNX_REQUIRE(D % V == 0 && cls_stride % V == 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(cls) && lnx_aligned16(patches) && lnx_aligned16(tokens) && lnx_aligned16(extras), LNX_ERR_ALIGN);
  const long long total = (long long)B * (1 + n_meta + n_patch) * (D / V);
  const int threads = 256;
  const int blocks = (int)min((long long)kNumSMs * 16, (total + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_F32)
    tokens_assemble_kernel<float><<<blocks, threads, 0, st>>>((const float*)cls, cls_stride, (const float*)extras, (const float*)patches, (float*)tokens, B, n_meta, n_patch, D);
  else if (dtype == LNX_BF16)
    tokens_assemble_kernel<bf16><<<blocks, threads, 0, st>>>((const bf16*)cls, cls_stride, (const bf16*)extras, (const bf16*)patches, (bf16*)tokens, B, n_meta, n_patch, D);
  else
    return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_tokens_split(const void* tokens, void* cls_out, void* extras_out, void* patches_out,
                                int B, int n_meta, int n_patch, int D, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(tokens, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && n_meta >= 0 && n_patch > 0 && D > 0, LNX_ERR_SHAPE);
  const int V = dtype == LNX_F32 ? 4 : 8;
  LNX_REQUIRE(D % V == 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(tokens) && lnx_aligned16(cls_out) && lnx_aligned16(extras_out) && lnx_aligned16(patches_out), LNX_ERR_ALIGN);
  const long long total = (long long)B * (1 + n_meta + n_patch) * (D / V);
  const int threads = 256;
  const int blocks = (int)min((long long)kNumSMs * 16, (total + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_F32)
    tokens_split_kernel<float><<<blocks, threads, 0, st>>>((const float*)tokens, (float*)cls_out, (float*)extras_out, (float*)patches_out, B, n_meta, n_patch, D);
  else if (dtype == LNX_BF16)
    tokens_split_kernel<bf16><<<blocks, threads, 0, st>>>((const bf16*)tokens, (bf16*)cls_out, (bf16*)extras_out, (bf16*)patches_out, B, n_meta, n_patch, D);
  else
    return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

// ---------------------------------------------------------------- column sum
// block = 32 x 8 threads; each block walks a slab of rows; one atomic per column per block
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, float* __restrict__ out, long long M, int N, long long rows_per_block) {
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float acc = 0.f;
  if (col < N)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) acc += to_f32(x[r * N + col]);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    atomicAdd(out + col, t);
  }
}

// 16-byte loads: a block owns up to 256 column chunks x a slab of rows; thread = (row-in-pass, chunk)
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, float* __restrict__ out, long long M, int N,
                                                         long long rows_per_block) {
  constexpr int V = Vec16<T>::N;
  __shared__ float red[256 * V];
  const int cpr = N / V;
  const int cx0 = blockIdx.x * 256;
  const int cw = min(256, cpr - cx0);
  const int rp = 256 / cw;
  const int rr = threadIdx.x / cw, cc = threadIdx.x % cw;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  if (rr < rp) {
    const T* px = x + (long long)(cx0 + cc) * V;
    long long r = r0 + rr;
    for (; r + 3LL * rp < r1; r += 4LL * rp) {  // 4 independent 16-byte loads in flight
      const Vec16<T> a = ld16(px + r * N), b = ld16(px + (r + rp) * N), c = ld16(px + (r + 2LL * rp) * N), d = ld16(px + (r + 3LL * rp) * N);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += (a.get(i) + b.get(i)) + (c.get(i) + d.get(i));
    }
    for (; r < r1; r += rp) {
      const Vec16<T> a = ld16(px + r * N);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += a.get(i);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) red[(rr * cw + cc) * V + i] = acc[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cw * V; i += 256) {
    float t = 0.f;
    for (int j = 0; j < rp; ++j) t += red[j * cw * V + i];
    atomicAdd(out + (long long)cx0 * V + i, t);
  }
}

extern "C" int lnx_colsum(const void* x, float* out, int64_t M, int N, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(x && out, LNX_ERR_NULL);
  LNX_REQUIRE(M > 0 && N > 0, LNX_ERR_SHAPE);
  {
    const int V = dtype == LNX_F32 ? 4 : 8;
    if ((dtype == LNX_F32 || dtype == LNX_BF16) && N % V == 0 && lnx_aligned16(x)) {
      const int cpr = N / V;
      const int gx = (cpr + 255) / 256;
      const long long Ml = (long long)M;
      int gy = (int)max(1LL, min((Ml + 255) / 256, (long long)(kNumSMs * 8 + gx - 1) / gx));
      const long long rpb = (Ml + gy - 1) / gy;
      gy = (int)((Ml + rpb - 1) / rpb);
      dim3 grid(gx, gy);
      if (dtype == LNX_F32) colsum_vec_kernel<float><<<grid, 256, 0, (cudaStream_t)s>>>((const float*)x, out, Ml, N, rpb);
      else colsum_vec_kernel<bf16><<<grid, 256, 0, (cudaStream_t)s>>>((const bf16*)x, out, Ml, N, rpb);
      LNX_CHECK_LAUNCH();
      return LNX_OK;
    }
  }
  const int gx = (N + 31) / 32;
  const long long Ml = (long long)M;
  int gy = max(1, min((int)((Ml + 63) / 64), (kNumSMs * 8 + gx - 1) / gx));
  const long long rpb = (Ml + gy - 1) / gy;
  gy = (int)((Ml + rpb - 1) / rpb);
  dim3 grid(gx, gy), block(32, 8);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_F32) colsum_kernel<float><<<grid, block, 0, st>>>((const float*)x, out, M, N, rpb);
  else if (dtype == LNX_BF16) colsum_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)x, out, M, N, rpb);
  else return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

// ---------------------------------------------------------------- cast
__global__ void cast_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(in)[i];
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(out)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < n - n4 * 4) out[n4 * 4 + threadIdx.x] = __float2bfloat16_rn(in[n4 * 4 + threadIdx.x]);
}

extern "C" int lnx_cast_f32_to_bf16(const float* in, void* out, int64_t n, lnx_stream_t s) {
  LNX_REQUIRE(in && out, LNX_ERR_NULL);
  LNX_REQUIRE(n > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(in) && (reinterpret_cast<uintptr_t>(out) & 7u) == 0, LNX_ERR_ALIGN);
  const int threads = 256;
  const int blocks = (int)max(1LL, min((long long)kNumSMs * 16, ((long long)n / 4 + threads - 1) / threads));
  cast_kernel<<<blocks, threads, 0, (cudaStream_t)s>>>(in, (bf16*)out, n);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

// ---------------------------------------------------------------- row scale (DropPath backward)
template <typename T>
__global__ void rowscale_kernel(const T* __restrict__ x, const float* __restrict__ sc, T* __restrict__ out, long long M, int N, int rpg) {
  constexpr int V = Vec16<T>::N;
  const int Nv = N / V;
  const long long total = M * Nv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / Nv;
    const float f = sc[m / rpg];
    Vec16<T> v = ld16(x + i * V);
#pragma unroll
    for (int j = 0; j < V; ++j) v.set(j, v.get(j) * f);
    st16(out + i * V, v);
  }
}

extern "C" int lnx_rowscale(const void* x, const float* s, void* out, int64_t M, int N, int rows_per_group, int dtype, lnx_stream_t st) {
  LNX_REQUIRE(x && s && out, LNX_ERR_NULL);
  LNX_REQUIRE(M > 0 && N > 0 && rows_per_group > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(x) && lnx_aligned16(out), LNX_ERR_ALIGN);
  const int V = dtype == LNX_F32 ? 4 : 8;
  LNX_REQUIRE(N % V == 0, LNX_ERR_SHAPE);
  const long long total = (long long)M * (N / V);
  const int threads = 256, blocks = (int)min((long long)kNumSMs * 16, (total + threads - 1) / threads);
  if (dtype == LNX_F32) rowscale_kernel<float><<<blocks, threads, 0, (cudaStream_t)st>>>((const float*)x, s, (float*)out, (long long)M, N, rows_per_group);
  else if (dtype == LNX_BF16) rowscale_kernel<bf16><<<blocks, threads, 0, (cudaStream_t)st>>>((const bf16*)x, s, (bf16*)out, (long long)M, N, rows_per_group);
  else return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

// ---------------------------------------------------------------- layer scale (ConvNeXt gamma) plumbing of the pointwise pair
// w_eff[n, :] = bf16(w[n, :] * cs[n]): the weight the data-gradient kernel sees (gamma folded in), one launch instead of mul + cast
__global__ void rowscale_cast_kernel(const float* __restrict__ w, const float* __restrict__ cs, bf16* __restrict__ out, int N, int K) {
  const long long total = (long long)N * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(w[i] * cs[i / K]);
}

extern "C" int lnx_rowscale_cast_bf16(const float* w, const float* cs, void* out, int N, int K, lnx_stream_t st) {
  LNX_REQUIRE(w && cs && out, LNX_ERR_NULL);
  LNX_REQUIRE(N > 0 && K > 0, LNX_ERR_SHAPE);
  const long long total = (long long)N * K;
  rowscale_cast_kernel<<<(int)min((long long)kNumSMs * 4, (total + 255) / 256), 256, 0, (cudaStream_t)st>>>(w, cs, (bf16*)out, N, K);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

// From the UN-scaled gradients of y = cs * (h W2^T + b2):  dW2 += cs dW2_raw,  db2 += cs db2_raw,
// dcs += rowsum(dW2_raw * W2) + b2 db2_raw  (no pass over activations: SURVEY 8(a4), convnext.py:84).  One block per output row.
__global__ void layerscale_bwd_kernel(const float* __restrict__ dw_raw, const float* __restrict__ db_raw, const float* __restrict__ w,
                                      const float* __restrict__ b, const float* __restrict__ cs, float* __restrict__ dw,
                                      float* __restrict__ db, float* __restrict__ dcs, int K) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  const float c = cs[n];
  float acc = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float g = dw_raw[(long long)n * K + k];
    acc = fmaf(g, w[(long long)n * K + k], acc);
    dw[(long long)n * K + k] += g * c;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    const float gb = db_raw ? db_raw[n] : 0.f;
    if (db) db[n] += gb * c;
    dcs[n] += t + (b ? b[n] : 0.f) * gb;
  }
}

extern "C" int lnx_layerscale_bwd(const float* dw_raw, const float* db_raw, const float* w, const float* b, const float* cs, float* dw,
                                  float* db, float* dcs, int N, int K, lnx_stream_t st) {
  LNX_REQUIRE(dw_raw && w && cs && dw && dcs, LNX_ERR_NULL);
  LNX_REQUIRE(N > 0 && K > 0, LNX_ERR_SHAPE);
  layerscale_bwd_kernel<<<N, 128, 0, (cudaStream_t)st>>>(dw_raw, db_raw, w, b, cs, dw, db, dcs, K);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

// ---------------------------------------------------------------- activation backward
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ pre, T* __restrict__ out, long long n, int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float u = to_f32(pre[i]), g = to_f32(dy[i]);
    out[i] = from_f32<T>(act == LNX_ACT_GELU ? g * gelu_grad_f(u) : (u > 0.f ? g : 0.f));
  }
}

extern "C" int lnx_act_bwd(const void* dy, const void* pre, void* out, int64_t n, int act, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(dy && pre && out, LNX_ERR_NULL);
  LNX_REQUIRE(n > 0 && (act == LNX_ACT_GELU || act == LNX_ACT_RELU), LNX_ERR_SHAPE);
  const long long nl = (long long)n;
  const int threads = 256, blocks = (int)min((long long)kNumSMs * 8, (nl + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_F32) act_bwd_kernel<float><<<blocks, threads, 0, st>>>((const float*)dy, (const float*)pre, (float*)out, nl, act);
  else if (dtype == LNX_BF16) act_bwd_kernel<bf16><<<blocks, threads, 0, st>>>((const bf16*)dy, (const bf16*)pre, (bf16*)out, nl, act);
  else return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

// ---------------------------------------------------------------- aggregate (Conv1d 2->1, k=1)
template <typename T>
__global__ void aggregate2_fwd_kernel(const T* __restrict__ a, const T* __restrict__ c, const float* __restrict__ w2,
                                      const float* __restrict__ bias1, T* __restrict__ out, long long n) {
  const float w0 = w2[0], w1 = w2[1], b = bias1[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = from_f32<T>(w0 * to_f32(a[i]) + w1 * to_f32(c[i]) + b);
}

template <typename T>
__global__ void aggregate2_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ a, const T* __restrict__ c,
                                      const float* __restrict__ w2, T* __restrict__ da, T* __restrict__ dc,
                                      float* __restrict__ dw2, float* __restrict__ dbias1, long long n) {
  const float w0 = w2[0], w1 = w2[1];
  float s0 = 0.f, s1 = 0.f, sb = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = to_f32(dout[i]);
    da[i] = from_f32<T>(w0 * g);
    dc[i] = from_f32<T>(w1 * g);
    s0 += g * to_f32(a[i]);
    s1 += g * to_f32(c[i]);
    sb += g;
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); sb = warp_sum(sb);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(dw2 + 0, s0);
    atomicAdd(dw2 + 1, s1);
    atomicAdd(dbias1, sb);
  }
}

extern "C" int lnx_aggregate2_fwd(const void* a, const void* c, const float* w2, const float* bias1, void* out, int64_t B, int D, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(a && c && w2 && bias1 && out, LNX_ERR_NULL);
  const long long n = (long long)B * D;
  LNX_REQUIRE(n > 0, LNX_ERR_SHAPE);
  const int threads = 256, blocks = (int)min((long long)kNumSMs * 4, (n + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_F32) aggregate2_fwd_kernel<float><<<blocks, threads, 0, st>>>((const float*)a, (const float*)c, w2, bias1, (float*)out, n);
  else if (dtype == LNX_BF16) aggregate2_fwd_kernel<bf16><<<blocks, threads, 0, st>>>((const bf16*)a, (const bf16*)c, w2, bias1, (bf16*)out, n);
  else return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_aggregate2_bwd(const void* dout, const void* a, const void* c, const float* w2, void* da, void* dc, float* dw2, float* dbias1,
                                  int64_t B, int D, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(dout && a && c && w2 && da && dc && dw2 && dbias1, LNX_ERR_NULL);
  const long long n = (long long)B * D;
  LNX_REQUIRE(n > 0, LNX_ERR_SHAPE);
  const int threads = 256, blocks = (int)min((long long)kNumSMs, (n + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_F32) aggregate2_bwd_kernel<float><<<blocks, threads, 0, st>>>((const float*)dout, (const float*)a, (const float*)c, w2, (float*)da, (float*)dc, dw2, dbias1, n);
  else if (dtype == LNX_BF16) aggregate2_bwd_kernel<bf16><<<blocks, threads, 0, st>>>((const bf16*)dout, (const bf16*)a, (const bf16*)c, w2, (bf16*)da, (bf16*)dc, dw2, dbias1, n);
  else return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
