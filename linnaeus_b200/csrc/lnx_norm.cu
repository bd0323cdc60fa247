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
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const T* __restrict__ dres, T* __restrict__ dx, float* __restrict__ dw,
                                                     float* __restrict__ db, long long rows, int C, int lanes) {
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
        Vec16<T> o, rv;
        if (dres) rv = ld16(dres + row * C + (long long)cv * V);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float xh = (xv[j].get(i) - mu) * rs;
          const float gw = gv[j].get(i) * w[cv * V + i];
          float r = rs * (gw - c1 - xh * c2);
          if (dres) r += rv.get(i);
          o.set(i, r);
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


// ------------------------------------------------------------------ bf16 fast path
// Exact fit (lanes * NV 16-byte vectors = one row, lanes in {4, 8, 16, 32}): no idle lanes at C = 96 / 192 / 384 / 768
// (NV = 3).  gamma / beta live in shared memory as float4 (read as conflict-free LDS.128), and the NEXT row of a
// lane group is loaded before the current one is reduced, so every thread keeps 2 * NV (forward) / 4 * NV
// (backward) 16-byte loads in flight -- the kernels are HBM-latency bound otherwise.
__device__ __forceinline__ void unpack8f(const uint4& raw, float* u) {
  u[0] = __uint_as_float(raw.x << 16); u[1] = __uint_as_float(raw.x & 0xffff0000u);
  u[2] = __uint_as_float(raw.y << 16); u[3] = __uint_as_float(raw.y & 0xffff0000u);
  u[4] = __uint_as_float(raw.z << 16); u[5] = __uint_as_float(raw.z & 0xffff0000u);
  u[6] = __uint_as_float(raw.w << 16); u[7] = __uint_as_float(raw.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2f(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint4 pack8f(const float* v) {
  return make_uint4(pack2f(v[0], v[1]), pack2f(v[2], v[3]), pack2f(v[4], v[5]), pack2f(v[6], v[7]));
}
template <int LANES>
__device__ __forceinline__ float group_sum_c(float v) {
#pragma unroll
  for (int o = LANES >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV, int LANES, bool RES>
__global__ void __launch_bounds__(256) ln_fwd_bf16_kernel(const uint4* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                          const uint4* __restrict__ res, uint4* __restrict__ y, float* __restrict__ mean,
                                                          float* __restrict__ rstd, long long rows, float eps) {
  constexpr int CV = NV * LANES;  // 16-byte vectors per row
  constexpr int C = CV * 8;
  constexpr int GPB = 256 / LANES;
  __shared__ float4 s_w[CV * 2], s_b[CV * 2];
  for (int i = threadIdx.x; i < C; i += 256) {
    reinterpret_cast<float*>(s_w)[i] = w[i];
    reinterpret_cast<float*>(s_b)[i] = b[i];
  }
  __syncthreads();
  const int gid = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const float invC = 1.0f / (float)C;
  const long long stride = (long long)gridDim.x * GPB;
  uint4 cur[NV], nxt[NV], rcur[NV], rnxt[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) cur[j] = nxt[j] = rcur[j] = rnxt[j] = make_uint4(0u, 0u, 0u, 0u);
  {
    const long long row = (long long)blockIdx.x * GPB + gid;
    if (row < rows) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        cur[j] = __ldg(x + row * CV + lane + j * LANES);
        if (RES) rcur[j] = __ldg(res + row * CV + lane + j * LANES);
      }
    }
  }
  // the loop bound is CTA-uniform (the shuffles below are full-warp); rows past the end are computed on zeros, not stored
  for (long long base = (long long)blockIdx.x * GPB; base < rows; base += stride) {
    const long long row = base + gid;
    const bool valid = row < rows;
    const long long nrow = row + stride;
    if (nrow < rows) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        nxt[j] = __ldg(x + nrow * CV + lane + j * LANES);
        if (RES) rnxt[j] = __ldg(res + nrow * CV + lane + j * LANES);
      }
    }
    float xv[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      unpack8f(cur[j], xv[j]);
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += xv[j][i];
    }
    const float mu = group_sum_c<LANES>(sum) * invC;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xv[j][i] -= mu;
        sq = fmaf(xv[j][i], xv[j][i], sq);
      }
    const float rs = 1.0f / sqrtf(group_sum_c<LANES>(sq) * invC + eps);
    if (valid && lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int cv = lane + j * LANES;
      const float4 w0 = s_w[2 * cv], w1 = s_w[2 * cv + 1], b0 = s_b[2 * cv], b1 = s_b[2 * cv + 1];
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(xv[j][i] * rs, wv[i], bv[i]);
      if (RES) {
        float r[8];
        unpack8f(rcur[j], r);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += r[i];
      }
      if (valid) y[row * CV + cv] = pack8f(o);
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      cur[j] = nxt[j];
      if (RES) rcur[j] = rnxt[j];
    }
  }
}

template <int VEC>
struct RawVec;
template <>
struct RawVec<8> {
  typedef uint4 type;
  static __device__ __forceinline__ void unpack(const uint4& r, float* u) { unpack8f(r, u); }
  static __device__ __forceinline__ uint4 pack(const float* v) { return pack8f(v); }
};
template <>
struct RawVec<4> {
  typedef uint2 type;
  static __device__ __forceinline__ void unpack(const uint2& r, float* u) {
    u[0] = __uint_as_float(r.x << 16); u[1] = __uint_as_float(r.x & 0xffff0000u);
    u[2] = __uint_as_float(r.y << 16); u[3] = __uint_as_float(r.y & 0xffff0000u);
  }
  static __device__ __forceinline__ uint2 pack(const float* v) { return make_uint2(pack2f(v[0], v[1]), pack2f(v[2], v[3])); }
};

// VEC = bf16 elements per thread slot (8 or 4): 4 halves the 2 * NV * VEC per-thread gamma / beta accumulators so
// that three CTAs stay resident at C = 96 / 192 / 384.
template <int NV, int LANES, int VEC>
__global__ void __launch_bounds__(256, (VEC == 4 ? 2 : 1))
    ln_bwd_bf16_kernel(const void* __restrict__ dy_, const void* __restrict__ x_, const float* __restrict__ w, const float* __restrict__ mean,
                       const float* __restrict__ rstd, const void* __restrict__ dres_, void* __restrict__ dx_, float* __restrict__ dw,
                       float* __restrict__ db, long long rows) {
  typedef typename RawVec<VEC>::type RV;
  const RV* __restrict__ dres = reinterpret_cast<const RV*>(dres_);  // optional: skip-connection gradient added to dx
  const RV* __restrict__ dy = reinterpret_cast<const RV*>(dy_);
  const RV* __restrict__ x = reinterpret_cast<const RV*>(x_);
  RV* __restrict__ dx = reinterpret_cast<RV*>(dx_);
  constexpr int CV = NV * LANES;  // vectors per row
  constexpr int C = CV * VEC;
  constexpr int GPB = 256 / LANES;
  __shared__ __align__(16) float s_w[C];
  __shared__ float s_red[2 * C];
  for (int i = threadIdx.x; i < C; i += 256) {
    s_w[i] = w[i];
    s_red[i] = 0.f;
    s_red[C + i] = 0.f;
  }
  __syncthreads();
  const int gid = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const float invC = 1.0f / (float)C;
  const long long stride = (long long)gridDim.x * GPB;
  float dw_acc[NV][VEC], db_acc[NV][VEC];
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int i = 0; i < VEC; ++i) dw_acc[j][i] = 0.f, db_acc[j][i] = 0.f;
  // two rows of lookahead: the loads of rows r + stride and r + 2 stride are in flight while row r is reduced
  RV xc[NV], gc[NV], xn[NV], gn[NV], xm[NV], gm[NV];
  float mu = 0.f, rs = 0.f, mun = 0.f, rsn = 0.f, mum = 0.f, rsm = 0.f;
  {
    const long long row = (long long)blockIdx.x * GPB + gid;
    const bool ok = row < rows;
    const long long row1 = row + stride;
    const bool ok1 = row1 < rows;
#pragma unroll
    for (int j = 0; j < NV; ++j) {  // rows past the end read row 0 and get rstd = 0: every contribution vanishes
      xc[j] = __ldg(x + (ok ? row : 0) * CV + lane + j * LANES);
      gc[j] = __ldg(dy + (ok ? row : 0) * CV + lane + j * LANES);
      xn[j] = __ldg(x + (ok1 ? row1 : 0) * CV + lane + j * LANES);
      gn[j] = __ldg(dy + (ok1 ? row1 : 0) * CV + lane + j * LANES);
    }
    if (ok) {
      mu = mean[row];
      rs = rstd[row];
    }
    if (ok1) {
      mun = mean[row1];
      rsn = rstd[row1];
    }
  }
  // the loop bound is CTA-uniform (the shuffles below are full-warp)
  for (long long base = (long long)blockIdx.x * GPB; base < rows; base += stride) {
    const long long row = base + gid;
    const bool valid = row < rows;
    const long long mrow = row + 2 * stride;
    mum = 0.f;
    rsm = 0.f;
    if (mrow < rows) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        xm[j] = __ldg(x + mrow * CV + lane + j * LANES);
        gm[j] = __ldg(dy + mrow * CV + lane + j * LANES);
      }
      mum = mean[mrow];
      rsm = rstd[mrow];
    }
    // pass 1: row statistics and the per-column gamma / beta partial sums (nothing kept but the raw vectors)
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int cv = lane + j * LANES;
      float wv[VEC], xh[VEC], g[VEC];
#pragma unroll
      for (int i = 0; i < VEC; i += 4) *reinterpret_cast<float4*>(wv + i) = *reinterpret_cast<const float4*>(s_w + cv * VEC + i);
      RawVec<VEC>::unpack(xc[j], xh);
      RawVec<VEC>::unpack(gc[j], g);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        xh[i] = (xh[i] - mu) * rs;
        const float gw = g[i] * wv[i];
        s1 += gw;
        s2 = fmaf(gw, xh[i], s2);
        dw_acc[j][i] = fmaf(g[i], xh[i], dw_acc[j][i]);
        db_acc[j][i] += valid ? g[i] : 0.f;
      }
    }
    const float c1 = group_sum_c<LANES>(s1) * invC;
    const float c2 = group_sum_c<LANES>(s2) * invC;
    // pass 2: dx, recomputing x-hat and g*gamma from the raw vectors (cheaper than keeping them live)
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int cv = lane + j * LANES;
      float wv[VEC], xh[VEC], g[VEC], o[VEC];
#pragma unroll
      for (int i = 0; i < VEC; i += 4) *reinterpret_cast<float4*>(wv + i) = *reinterpret_cast<const float4*>(s_w + cv * VEC + i);
      RawVec<VEC>::unpack(xc[j], xh);
      RawVec<VEC>::unpack(gc[j], g);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float xhat = (xh[i] - mu) * rs;
        o[i] = rs * (fmaf(g[i], wv[i], -c1) - xhat * c2);
      }
      if (dres && valid) {
        float rv[VEC];
        RawVec<VEC>::unpack(__ldg(dres + row * CV + cv), rv);
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] += rv[i];
      }
      if (valid) dx[row * CV + cv] = RawVec<VEC>::pack(o);
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      xc[j] = xn[j];
      gc[j] = gn[j];
      xn[j] = xm[j];
      gn[j] = gm[j];
    }
    mu = mun;
    rs = rsn;
    mun = mum;
    rsn = rsm;
  }
  // lanes of different groups hold partial sums of the same columns: fold through shared memory, then one global
  // atomic per column per CTA
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int cv = lane + j * LANES;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      atomicAdd(&s_red[cv * VEC + i], dw_acc[j][i]);
      atomicAdd(&s_red[C + cv * VEC + i], db_acc[j][i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) {
    atomicAdd(dw + i, s_red[i]);
    atomicAdd(db + i, s_red[C + i]);
  }
}

// NV in {3, 4, 2, 1} with Cv = NV * lanes, lanes a power of two in [4, 32]
inline bool fast_plan(int Cv, int* nv, int* lanes) {
  const int cand[4] = {3, 4, 2, 1};
  for (int k = 0; k < 4; ++k) {
    const int n = cand[k];
    if (Cv % n) continue;
    const int l = Cv / n;
    if (l == 4 || l == 8 || l == 16 || l == 32) {
      *nv = n;
      *lanes = l;
      return true;
    }
  }
  return false;
}

#define LNX_LN_FAST_DISPATCH(MACRO) \
  if (nv == 3 && lanes == 4) { MACRO(3, 4) } else if (nv == 3 && lanes == 8) { MACRO(3, 8) } else if (nv == 3 && lanes == 16) { MACRO(3, 16) } \
  else if (nv == 3 && lanes == 32) { MACRO(3, 32) } else if (nv == 4 && lanes == 4) { MACRO(4, 4) } else if (nv == 4 && lanes == 8) { MACRO(4, 8) } \
  else if (nv == 4 && lanes == 16) { MACRO(4, 16) } else if (nv == 4 && lanes == 32) { MACRO(4, 32) } else if (nv == 2 && lanes == 4) { MACRO(2, 4) } \
  else if (nv == 1 && lanes == 4) { MACRO(1, 4) } else { return LNX_ERR_UNSUPPORTED; }

int ln_fwd_fast(const void* x, const float* w, const float* b, const void* res, void* y, float* mean, float* rstd, long long rows, int C,
                float eps, cudaStream_t st) {
  int nv, lanes;
  if (C % 8 != 0 || !fast_plan(C / 8, &nv, &lanes)) return LNX_ERR_UNSUPPORTED;
  const int gpb = 256 / lanes;
  const int blocks = (int)max(1LL, min((long long)kNumSMs * 6, (rows + gpb - 1) / gpb));
#define LNX_F(NVV, LL)                                                                                                                      \
  if (res) ln_fwd_bf16_kernel<NVV, LL, true><<<blocks, 256, 0, st>>>((const uint4*)x, w, b, (const uint4*)res, (uint4*)y, mean, rstd, rows, eps); \
  else ln_fwd_bf16_kernel<NVV, LL, false><<<blocks, 256, 0, st>>>((const uint4*)x, w, b, nullptr, (uint4*)y, mean, rstd, rows, eps);
  LNX_LN_FAST_DISPATCH(LNX_F)
#undef LNX_F
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

int ln_bwd_fast(const void* dy, const void* x, const float* w, const float* mean, const float* rstd, const void* dres, void* dx, float* dw,
                float* db, long long rows, int C, cudaStream_t st) {
  int nv, lanes;
#define LNX_B(NVV, LL, VV)                                                                                                   \
  {                                                                                                                          \
    const int gpb = 256 / LL;                                                                                                \
    const int blocks = (int)max(1LL, min((long long)kNumSMs * 2, (rows + gpb - 1) / gpb));                    \
    ln_bwd_bf16_kernel<NVV, LL, VV><<<blocks, 256, 0, st>>>(dy, x, w, mean, rstd, dres, dx, dw, db, rows);                          \
  }
  if (C % 12 == 0 && (C / 12 == 8 || C / 12 == 16 || C / 12 == 32)) {  // 4-element slots, 3 per thread: C = 96 / 192 / 384
    if (C == 96) LNX_B(3, 8, 4) else if (C == 192) LNX_B(3, 16, 4) else LNX_B(3, 32, 4)
    LNX_CHECK_LAUNCH();
    return LNX_OK;
  }
  if (C % 8 != 0 || !fast_plan(C / 8, &nv, &lanes)) return LNX_ERR_UNSUPPORTED;
#define LNX_B8(NVV, LL) LNX_B(NVV, LL, 8)
  LNX_LN_FAST_DISPATCH(LNX_B8)
#undef LNX_B8
#undef LNX_B
  LNX_CHECK_LAUNCH();
  return LNX_OK;
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
int ln_bwd_launch(const void* dy, const void* x, const float* w, const float* mean, const float* rstd, const void* dres, void* dx,
                  float* dw, float* db, long long rows, int C, cudaStream_t st) {
  constexpr int V = Vec16<T>::N;
  LnPlan p;
  if (C % V != 0 || !ln_plan(C / V, &p)) return LNX_ERR_SHAPE;
  const int gpb = 256 / p.lanes;
  const int blocks = (int)max(1LL, min((long long)kNumSMs * 4, (rows + gpb - 1) / gpb));
  const size_t smem = 2 * (size_t)C * sizeof(float);
#define LNX_LN_B(NVV)                                                                                                         \
  case NVV:                                                                                                                   \
    ln_bwd_kernel<T, NVV><<<blocks, 256, smem, st>>>((const T*)dy, (const T*)x, w, mean, rstd, (const T*)dres, (T*)dx, dw, db, rows, C, p.lanes); \
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
  if (dtype == LNX_BF16) {
    const int r = ln_fwd_fast(x, w, b, residual, y, mean, rstd, rows, C, eps, (cudaStream_t)s);
    if (r != LNX_ERR_UNSUPPORTED) return r;
    return ln_fwd_launch<bf16>(x, w, b, residual, y, mean, rstd, rows, C, eps, (cudaStream_t)s);
  }
  return LNX_ERR_DTYPE;
}

extern "C" int lnx_layernorm_bwd(const void* dy, const void* x, const float* w, const float* mean, const float* rstd, const void* dres,
                                 void* dx, float* dw, float* db, int64_t rows, int C, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(dy && x && w && mean && rstd && dx && dw && db, LNX_ERR_NULL);
  LNX_REQUIRE(rows > 0 && C > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(x) && lnx_aligned16(dy) && lnx_aligned16(dx) && lnx_aligned16(dres), LNX_ERR_ALIGN);
  if (dtype == LNX_F32) return ln_bwd_launch<float>(dy, x, w, mean, rstd, dres, dx, dw, db, rows, C, (cudaStream_t)s);
  if (dtype == LNX_BF16) {
    const int r = ln_bwd_fast(dy, x, w, mean, rstd, dres, dx, dw, db, rows, C, (cudaStream_t)s);
    if (r != LNX_ERR_UNSUPPORTED) return r;
    return ln_bwd_launch<bf16>(dy, x, w, mean, rstd, dres, dx, dw, db, rows, C, (cudaStream_t)s);
  }
  return LNX_ERR_DTYPE;
}
