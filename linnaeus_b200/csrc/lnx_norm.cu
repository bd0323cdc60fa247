// LayerNorm forward / backward over rows of C channels (token-major and NHWC).
// HBM-bound: one pass over x (+residual) and y; rows live in registers as 16-byte
// vectors, statistics by sub-warp shuffle reductions (8/16/32 lanes per row so
// small C does not idle lanes).  fp32 statistics, two-pass variance.
#include "lnx_common.cuh"

using namespace lnx;

namespace {

__device__ __forceinline__ float group_sum(float v, int lanes) {
  for (int o = lanes >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                     const T* __restrict__ res, T* __restrict__ y, float* __restrict__ mean,
                                                     float* __restrict__ rstd, long long rows, int C, float eps, int lanes) {
  constexpr int V = Vec16<T>::N;
  const int Cv = C / V;
  const int gpb = blockDim.x / lanes;
  const int gid = threadIdx.x / lanes, lane = threadIdx.x % lanes;
  const float invC = 1.0f / (float)C;
  for (long long base = (long long)blockIdx.x * gpb; base < rows; base += (long long)gridDim.x * gpb) {
    const long long row = base + gid;
    const bool valid = row < rows;
    Vec16<T> xv[NV];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int cv = lane + j * lanes;
      if (valid && cv < Cv) {
        xv[j] = ld16(x + row * C + (long long)cv * V);
#pragma unroll
        for (int i = 0; i < V; ++i) sum += xv[j].get(i);
      }
    }
    const float mu = group_sum(sum, lanes) * invC;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int cv = lane + j * lanes;
      if (valid && cv < Cv) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float d = xv[j].get(i) - mu;
          sq += d * d;
        }
      }
    }
    const float var = group_sum(sq, lanes) * invC;
    const float rs = 1.0f / sqrtf(var + eps);
    if (valid && lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int cv = lane + j * lanes;
      if (valid && cv < Cv) {
        Vec16<T> o, r;
        if (res) r = ld16(res + row * C + (long long)cv * V);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const int c = cv * V + i;
          float v = (xv[j].get(i) - mu) * rs * w[c] + b[c];
          if (res) v += r.get(i);
          o.set(i, v);
        }
        st16(y + row * C + (long long)cv * V, o);
      }
    }
  }
}

template <typename T, int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ dx,
                                                     float* __restrict__ dw, float* __restrict__ db, long long rows, int C, int lanes) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ float sred[];  // [2][C]
  const int Cv = C / V;
  const int gpb = blockDim.x / lanes;
  const int gid = threadIdx.x / lanes, lane = threadIdx.x % lanes;
  const float invC = 1.0f / (float)C;
  float dw_acc[NV][V], db_acc[NV][V];
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int i = 0; i < V; ++i) dw_acc[j][i] = 0.f, db_acc[j][i] = 0.f;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sred[i] = 0.f;

  for (long long base = (long long)blockIdx.x * gpb; base < rows; base += (long long)gridDim.x * gpb) {
    const long long row = base + gid;
    const bool valid = row < rows;
    const float mu = valid ? mean[row] : 0.f;
    const float rs = valid ? rstd[row] : 0.f;
    Vec16<T> xv[NV], gv[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int cv = lane + j * lanes;
      if (valid && cv < Cv) {
        xv[j] = ld16(x + row * C + (long long)cv * V);
        gv[j] = ld16(dy + row * C + (long long)cv * V);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float xh = (xv[j].get(i) - mu) * rs;
          const float g = gv[j].get(i);
          const float gw = g * w[cv * V + i];
          s1 += gw;
          s2 += gw * xh;
          dw_acc[j][i] += g * xh;
          db_acc[j][i] += g;
        }
      }
    }
    const float c1 = group_sum(s1, lanes) * invC;
    const float c2 = group_sum(s2, lanes) * invC;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int cv = lane + j * lanes;
      if (valid && cv < Cv) {
        Vec16<T> o;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float xh = (xv[j].get(i) - mu) * rs;
          const float gw = gv[j].get(i) * w[cv * V + i];
          o.set(i, rs * (gw - c1 - xh * c2));
        }
        st16(dx + row * C + (long long)cv * V, o);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int cv = lane + j * lanes;
    if (cv < Cv) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        atomicAdd(&sred[cv * V + i], dw_acc[j][i]);
        atomicAdd(&sred[C + cv * V + i], db_acc[j][i]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dw + i, sred[i]);
    atomicAdd(db + i, sred[C + i]);
  }
}

struct LnPlan {
  int lanes, nv;
};

inline bool ln_plan(int Cv, LnPlan* p) {
  if (Cv <= 24) {
    p->lanes = 8;
    p->nv = (Cv + 7) / 8;
  } else if (Cv <= 64) {
    p->lanes = 16;
    p->nv = (Cv + 15) / 16;
  } else {
    p->lanes = 32;
    p->nv = (Cv + 31) / 32;
  }
  if (p->nv > 16) return false;
  if (p->nv > 8) p->nv = 16;
  else if (p->nv > 4) p->nv = 8;
  return true;
}

template <typename T>
int ln_fwd_launch(const void* x, const float* w, const float* b, const void* res, void* y, float* mean, float* rstd,
                  long long rows, int C, float eps, cudaStream_t st) {
  constexpr int V = Vec16<T>::N;
  LnPlan p;
  if (C % V != 0 || !ln_plan(C / V, &p)) return LNX_ERR_SHAPE;
  const int gpb = 256 / p.lanes;
  const int blocks = (int)max(1LL, min((long long)kNumSMs * 8, (rows + gpb - 1) / gpb));
#define LNX_LN_F(NVV)                                                                                                           \
  case NVV:                                                                                                                     \
    ln_fwd_kernel<T, NVV><<<blocks, 256, 0, st>>>((const T*)x, w, b, (const T*)res, (T*)y, mean, rstd, rows, C, eps, p.lanes); \
    break;
  switch (p.nv) {
    LNX_LN_F(1) LNX_LN_F(2) LNX_LN_F(3) LNX_LN_F(4) LNX_LN_F(8) LNX_LN_F(16)
    default: return LNX_ERR_SHAPE;
  }
#undef LNX_LN_F
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

template <typename T>
int ln_bwd_launch(const void* dy, const void* x, const float* w, const float* mean, const float* rstd, void* dx, float* dw,
                  float* db, long long rows, int C, cudaStream_t st) {
  constexpr int V = Vec16<T>::N;
  LnPlan p;
  if (C % V != 0 || !ln_plan(C / V, &p)) return LNX_ERR_SHAPE;
  const int gpb = 256 / p.lanes;
  const int blocks = (int)max(1LL, min((long long)kNumSMs * 4, (rows + gpb - 1) / gpb));
  const size_t smem = 2 * (size_t)C * sizeof(float);
#define LNX_LN_B(NVV)                                                                                                         \
  case NVV:                                                                                                                   \
    ln_bwd_kernel<T, NVV><<<blocks, 256, smem, st>>>((const T*)dy, (const T*)x, w, mean, rstd, (T*)dx, dw, db, rows, C, p.lanes); \
    break;
  switch (p.nv) {
    LNX_LN_B(1) LNX_LN_B(2) LNX_LN_B(3) LNX_LN_B(4) LNX_LN_B(8) LNX_LN_B(16)
    default: return LNX_ERR_SHAPE;
  }
#undef LNX_LN_B
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

}  // namespace

extern "C" int lnx_layernorm_fwd(const void* x, const float* w, const float* b, const void* residual, void* y, float* mean, float* rstd,
                                 int64_t rows, int C, float eps, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(x && w && b && y, LNX_ERR_NULL);
  LNX_REQUIRE(rows > 0 && C > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(x) && lnx_aligned16(y) && lnx_aligned16(residual), LNX_ERR_ALIGN);
  if (dtype == LNX_F32) return ln_fwd_launch<float>(x, w, b, residual, y, mean, rstd, rows, C, eps, (cudaStream_t)s);
  if (dtype == LNX_BF16) return ln_fwd_launch<bf16>(x, w, b, residual, y, mean, rstd, rows, C, eps, (cudaStream_t)s);
  return LNX_ERR_DTYPE;
}

extern "C" int lnx_layernorm_bwd(const void* dy, const void* x, const float* w, const float* mean, const float* rstd, void* dx,
                                 float* dw, float* db, int64_t rows, int C, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(dy && x && w && mean && rstd && dx && dw && db, LNX_ERR_NULL);
  LNX_REQUIRE(rows > 0 && C > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(x) && lnx_aligned16(dy) && lnx_aligned16(dx), LNX_ERR_ALIGN);
  if (dtype == LNX_F32) return ln_bwd_launch<float>(dy, x, w, mean, rstd, dx, dw, db, rows, C, (cudaStream_t)s);
  if (dtype == LNX_BF16) return ln_bwd_launch<bf16>(dy, x, w, mean, rstd, dx, dw, db, rows, C, (cudaStream_t)s);
  return LNX_ERR_DTYPE;
}
