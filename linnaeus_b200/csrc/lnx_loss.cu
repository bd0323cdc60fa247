// Hierarchical multi-rank masked loss, fused across the K heads, no host syncs.
//   fwd    : one warp per (task k, sample i): log-softmax statistics + loss against the
//            target distribution T (one-hot / uniform smoothing / taxonomy soft-label row),
//            null masking (coin flips passed in) and PHASE1 masking
//   reduce : n_valid_k = #(l != 0); total = sum_k w_k sum_i l_ki / max(n_valid_k, 1e-6)
//   bwd    : dlogits = g * scale_k * active * (sum(T) softmax - T)
#include "lnx_common.cuh"

using namespace lnx;

namespace {

constexpr int MAXK = 16;
struct LossMeta {
  int off[MAXK + 1];
  const float* mats[MAXK];
};

template <typename T>
__device__ __forceinline__ float target_prob(int kind, float smoothing, const float* __restrict__ mat_row, int C, int y, int c) {
  if (kind == LNX_LOSS_CE) return c == y ? 1.f : 0.f;
  if (kind == LNX_LOSS_LABEL_SMOOTHING) return c == y ? 1.f - smoothing : smoothing / (float)(C - 1);
  return mat_row[c];
}

template <typename T>
__global__ void loss_fwd_kernel(const T* __restrict__ logits, int B, int K, int Ctot, LossMeta meta, const long long* __restrict__ targets,
                                const unsigned char* __restrict__ null_flag, const float* __restrict__ keep, int kind, float smoothing,
                                int phase1, float* __restrict__ per_sample, float* __restrict__ raw, float* __restrict__ lse_out,
                                float* __restrict__ wgt_out) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)K * B;
  for (long long w = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); w < total; w += (long long)gridDim.x * warps_per_block) {
    const int k = (int)(w / B), i = (int)(w % B);
    const int off = meta.off[k], C = meta.off[k + 1] - off;
    const T* z = logits + (long long)i * Ctot + off;
    const int y = (int)targets[w];
    const float* mrow = (kind == LNX_LOSS_TAXONOMY) ? meta.mats[k] + (long long)y * C : nullptr;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, to_f32(z[c]));
    mx = warp_max(mx);
    float se = 0.f, tz = 0.f, ts = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float zc = to_f32(z[c]);
      se += expf(zc - mx);
      const float t = target_prob<T>(kind, smoothing, mrow, C, y, c);
      tz = fmaf(t, zc, tz);
      ts += t;
    }
    se = warp_sum(se); tz = warp_sum(tz); ts = warp_sum(ts);
    if (lane == 0) {
      const float L = mx + logf(se);
      const float loss = ts * L - tz;  // -sum_c T_c (z_c - L)
      const bool is_null = null_flag ? (null_flag[w] != 0) : (y == 0);
      float r = loss;
      float wgt = keep ? keep[w] : 1.f;   // per-sample multiplier: null-mask coin flips x class weights
      if (phase1 && is_null) { wgt = 0.f; r = 0.f; }  // ignore_index = 0 zeroes the criterion output itself
      per_sample[w] = loss * wgt;
      if (raw) raw[w] = r;
      lse_out[w] = L;
      wgt_out[w] = wgt;
    }
  }
}

// single block; K <= MAXK
__global__ void loss_reduce_kernel(const float* __restrict__ per_sample, const float* __restrict__ task_w, int B, int K, int phase1,
                                   float* __restrict__ total, float* __restrict__ scale, float* __restrict__ task_sum,
                                   float* __restrict__ nvalid) {
  __shared__ float s_sum[32], s_cnt[32];
  __shared__ float s_total;
  if (threadIdx.x == 0) s_total = 0.f;
  for (int k = 0; k < K; ++k) {
    float sum = 0.f, cnt = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
      const float l = per_sample[(long long)k * B + i];
      sum += l;
      cnt += (l != 0.f) ? 1.f : 0.f;
    }
    sum = warp_sum(sum); cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = sum, s_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      float ts = 0.f, tc = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) ts += s_sum[w], tc += s_cnt[w];
      const float nv = phase1 ? (float)B : tc;
      const float wk = task_w ? task_w[k] : 1.f;
      const float sc = wk / fmaxf(nv, 1e-6f);
      scale[k] = sc;
      nvalid[k] = nv;
      task_sum[k] = ts * sc;
      s_total += ts * sc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = s_total;
}

template <typename T>
__global__ void loss_bwd_kernel(const T* __restrict__ logits, int B, int K, int Ctot, LossMeta meta, const long long* __restrict__ targets,
                                const float* __restrict__ per_sample, int kind, float smoothing, const float* __restrict__ lse,
                                const float* __restrict__ scale, const float* __restrict__ gscale, T* __restrict__ dlogits) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)K * B;
  const float gs = gscale ? gscale[0] : 1.f;
  for (long long w = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); w < total; w += (long long)gridDim.x * warps_per_block) {
    const int k = (int)(w / B), i = (int)(w % B);
    const int off = meta.off[k], C = meta.off[k + 1] - off;
    const T* z = logits + (long long)i * Ctot + off;
    T* dz = dlogits + (long long)i * Ctot + off;
    const int y = (int)targets[w];
    const float* mrow = (kind == LNX_LOSS_TAXONOMY) ? meta.mats[k] + (long long)y * C : nullptr;
    // per_sample here is the multiplier written by the forward (0 for masked samples)
    const float f = gs * scale[k] * per_sample[w];
    const float L = lse[w];
    float ts = 1.f;
    if (kind == LNX_LOSS_TAXONOMY) {
      float t = 0.f;
      for (int c = lane; c < C; c += 32) t += mrow[c];
      ts = warp_sum(t);
    }
    for (int c = lane; c < C; c += 32) {
      const float p = expf(to_f32(z[c]) - L);
      const float t = target_prob<T>(kind, smoothing, mrow, C, y, c);
      dz[c] = from_f32<T>(f * (ts * p - t));
    }
  }
}

int fill_meta(LossMeta* m, int K, const int* class_off, const float* const* soft_mats, int kind) {
  if (K <= 0 || K > MAXK) return LNX_ERR_SHAPE;
  for (int k = 0; k <= K; ++k) m->off[k] = class_off[k];
  for (int k = 0; k < K; ++k) {
    m->mats[k] = soft_mats ? soft_mats[k] : nullptr;
    if (kind == LNX_LOSS_TAXONOMY && !m->mats[k]) return LNX_ERR_NULL;
    if (m->off[k + 1] <= m->off[k]) return LNX_ERR_SHAPE;
  }
  return LNX_OK;
}

}  // namespace

extern "C" int lnx_loss_fwd(const void* logits, int dtype, int B, int K, const int* class_off, const int64_t* targets,
                            const uint8_t* null_flag, const float* keep, int kind, float smoothing, const float* const* soft_mats,
                            int phase1, float* per_sample, float* raw, float* lse, float* sample_w, lnx_stream_t s) {
  LNX_REQUIRE(logits && class_off && targets && per_sample && lse && sample_w, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && kind >= 0 && kind <= 2, LNX_ERR_SHAPE);
  LossMeta m;
  const int r = fill_meta(&m, K, class_off, soft_mats, kind);
  if (r != LNX_OK) return r;
  const int Ctot = class_off[K];
  const long long total = (long long)K * B;
  const int blocks = (int)min((long long)kNumSMs * 4, (total + 7) / 8);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_F32)
    loss_fwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)logits, B, K, Ctot, m, (const long long*)targets, null_flag, keep, kind, smoothing, phase1, per_sample, raw, lse, sample_w);
  else if (dtype == LNX_BF16)
    loss_fwd_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)logits, B, K, Ctot, m, (const long long*)targets, null_flag, keep, kind, smoothing, phase1, per_sample, raw, lse, sample_w);
  else
    return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_loss_reduce(const float* per_sample, const float* task_w, int B, int K, int phase1, float* total, float* scale,
                               float* task_sum, float* nvalid, lnx_stream_t s) {
  LNX_REQUIRE(per_sample && total && scale && task_sum && nvalid, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && K > 0 && K <= MAXK, LNX_ERR_SHAPE);
  loss_reduce_kernel<<<1, 256, 0, (cudaStream_t)s>>>(per_sample, task_w, B, K, phase1, total, scale, task_sum, nvalid);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_loss_bwd(const void* logits, int dtype, int B, int K, const int* class_off, const int64_t* targets,
                            const float* sample_w, int kind, float smoothing, const float* const* soft_mats, const float* lse,
                            const float* scale, const float* gscale, void* dlogits, lnx_stream_t s) {
  LNX_REQUIRE(logits && class_off && targets && sample_w && lse && scale && dlogits, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && kind >= 0 && kind <= 2, LNX_ERR_SHAPE);
  LossMeta m;
  const int r = fill_meta(&m, K, class_off, soft_mats, kind);
  if (r != LNX_OK) return r;
  const int Ctot = class_off[K];
  const long long total = (long long)K * B;
  const int blocks = (int)min((long long)kNumSMs * 4, (total + 7) / 8);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_F32)
    loss_bwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)logits, B, K, Ctot, m, (const long long*)targets, sample_w, kind, smoothing, lse, scale, gscale, (float*)dlogits);
  else if (dtype == LNX_BF16)
    loss_bwd_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)logits, B, K, Ctot, m, (const long long*)targets, sample_w, kind, smoothing, lse, scale, gscale, (bf16*)dlogits);
  else
    return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
