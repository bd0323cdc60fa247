// Flat-buffer optimizer kernels: global grad norm, clip coefficient (device side, no
// host sync) and AdamW with decoupled weight decay.  HBM-bound: 16 B read + 12 B written
// per parameter.
#include "lnx_common.cuh"

using namespace lnx;

namespace {

// ws (optional): [0] ticket counter (zero on entry, left zero), [1 .. gridDim.x] per-block partial sums.  With a workspace the
// last block to finish adds the partials in index order, so the result does not depend on block scheduling: data-parallel
// replicas compute bit-identical clip coefficients from bit-identical gradients and stay in lock step.
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out, float* __restrict__ ws) {
  __shared__ float red[32];
  __shared__ bool last;
  float acc = 0.f;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < n - n4 * 4) {
    const float v = g[n4 * 4 + threadIdx.x];
    acc += v * v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      if (!ws) {
        atomicAdd(out, t);
      } else {
        ws[1 + blockIdx.x] = t;
        __threadfence();
        const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(ws), 1u);
        last = ticket == gridDim.x - 1;
      }
    }
  }
  if (!ws) return;
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) t += __ldcg(ws + 1 + i);  // fixed order per lane, fixed shuffle tree
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      out[0] += t;
      *reinterpret_cast<unsigned*>(ws) = 0u;
    }
  }
}

__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float gscale, float clip, float* __restrict__ norm_out,
                                 float* __restrict__ coef_out) {
  const float norm = sqrtf(sumsq[0]) * fabsf(gscale);
  if (norm_out) norm_out[0] = norm;
  coef_out[0] = (clip > 0.f) ? fminf(1.0f, clip / (norm + 1e-6f)) : 1.0f;
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, float gscale,
                             const float* __restrict__ coef, const float* __restrict__ lr_dev, const float* __restrict__ step_dev) {
  const float gs = gscale * (coef ? coef[0] : 1.0f);
  if (lr_dev) lr = lr_dev[0];
  if (step_dev) {  // CUDA-graph replays: the step count lives on the device
    const float t = step_dev[0];
    bc1 = 1.0f - powf(b1, t);
    bc2 = 1.0f - powf(b2, t);
  }
  const float step = lr / bc1;
  const float inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
  const float decay = 1.0f - lr * wd;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; const float* gg = &gv.x; float* mm = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gr = gg[e] * gs;
      pp[e] *= decay;
      mm[e] = b1 * mm[e] + (1.0f - b1) * gr;
      vp[e] = b2 * vp[e] + (1.0f - b2) * gr * gr;
      pp[e] -= step * mm[e] / (sqrtf(vp[e]) * inv_sqrt_bc2 + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < n - n4 * 4) {
    const long long i = n4 * 4 + threadIdx.x;
    const float gr = g[i] * gs;
    float pe = p[i] * decay;
    const float me = b1 * m[i] + (1.0f - b1) * gr;
    const float ve = b2 * v[i] + (1.0f - b2) * gr * gr;
    pe -= step * me / (sqrtf(ve) * inv_sqrt_bc2 + eps);
    p[i] = pe; m[i] = me; v[i] = ve;
  }
}

}  // namespace

extern "C" int lnx_sumsq(const float* g, int64_t n, float* sumsq, float* workspace, lnx_stream_t s) {
  LNX_REQUIRE(g && sumsq, LNX_ERR_NULL);
  LNX_REQUIRE(n > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(g), LNX_ERR_ALIGN);
  const int blocks = (int)max(1LL, min((long long)kNumSMs * 4, ((long long)n / 4 + 255) / 256));
  sumsq_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(g, n, sumsq, workspace);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_clip_coef(const float* sumsq, float gscale, float clip, float* norm_out, float* coef_out, lnx_stream_t s) {
  LNX_REQUIRE(sumsq && coef_out, LNX_ERR_NULL);
  clip_coef_kernel<<<1, 1, 0, (cudaStream_t)s>>>(sumsq, gscale, clip, norm_out, coef_out);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                         float weight_decay, float bias_corr1, float bias_corr2, float gscale, const float* coef,
                         const float* lr_dev, const float* step_dev, lnx_stream_t s) {
  LNX_REQUIRE(p && g && m && v, LNX_ERR_NULL);
  LNX_REQUIRE(n > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(p) && lnx_aligned16(g) && lnx_aligned16(m) && lnx_aligned16(v), LNX_ERR_ALIGN);
  const int blocks = (int)max(1LL, min((long long)kNumSMs * 8, ((long long)n / 4 + 255) / 256));
  adamw_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2, gscale, coef, lr_dev, step_dev);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
