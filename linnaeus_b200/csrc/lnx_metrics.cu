// Validation metrics over the K hierarchical heads, fused and sync-free.
//
// The reference (linnaeus/utils/metrics/tracker.py:609-735, chain_accuracy.py:51-364, basic.py:79-133) runs, per
// validation batch and per task, argmax + topk(3) + equality + .sum().item() - about 4K + 6 host syncs - and then
// stacks the K equality vectors for the chain / partial-chain accuracies.  Everything it needs is the RANK of the
// ground-truth class in each logits row:
//     rank[k, i] = #{ c : z[c] > z[y] or (z[c] == z[y] and c < y) }
// (argmax returns the first maximal index, so top-1 correct <=> rank == 0; target in the top n <=> rank < n, with the
// same lowest-index-first order among exact ties).  One CTA per sample: warp w handles tasks w, w + 4, ... (a strided
// count over the task's C_k logits, one warp reduction), then thread 0 folds the K ranks into the counters
//     [0, K)    top-1 correct per task            [K, 2K)   top-3 correct per task (C_k < 3: same as top-1,
//                                                            tracker.py:722-724)
//     [2K]      chain correct (all K ranks right) [2K + 1]  partial-chain correct (ranks 0..highest non-null right)
//     [2K + 2]  samples with any non-null target  [2K + 3]  samples
//     [2K + 4, 3K + 4)  top-1 correct among the samples whose target for that task IS null (tracker.py:797-848)
//     [3K + 4, 4K + 4)  samples whose target for that task is null (non-null accuracy = the complements)
// with integer atomics (order independent => bit-reproducible).  HBM traffic = the logits once (B * sum C_k elements).
#include "lnx_common.cuh"

using namespace lnx;

namespace {

constexpr int MAXK = 16;
struct MetricsMeta {
  int off[MAXK + 1];
};

template <typename T>
__global__ void __launch_bounds__(128) hier_metrics_kernel(const T* __restrict__ logits, long long ld, int B, int K, MetricsMeta meta,
                                                           const long long* __restrict__ targets, int null_index,
                                                           int* __restrict__ ranks_out, long long* __restrict__ counters) {
  __shared__ int s_rank[MAXK];
  const int i = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int k = warp; k < K; k += nwarps) {
    const int off = meta.off[k], C = meta.off[k + 1] - off;
    const T* z = logits + (long long)i * ld + off;
    const int y = (int)targets[(long long)k * B + i];
    int cnt = 0;
    if (y >= 0 && y < C) {
      const float zy = to_f32(z[y]);
      for (int c = lane; c < C; c += 32) {
        const float zc = to_f32(z[c]);
        cnt += (zc > zy || (zc == zy && c < y)) ? 1 : 0;
      }
      cnt = (int)warp_sum((float)cnt);  // C_k < 2^24: exact in fp32
    } else {
      cnt = C;  // a target outside the head's range can never be predicted
    }
    if (lane == 0) {
      s_rank[k] = cnt;
      if (ranks_out) ranks_out[(long long)k * B + i] = cnt;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && counters) {
    bool chain = true, partial = true;
    int highest = -1;
    for (int k = 0; k < K; ++k)
      if ((int)targets[(long long)k * B + i] != null_index) highest = k;
    for (int k = 0; k < K; ++k) {
      const int C = meta.off[k + 1] - meta.off[k];
      const bool top1 = s_rank[k] == 0;
      const bool top3 = C < 3 ? top1 : s_rank[k] < 3;
      if (top1) atomicAdd(reinterpret_cast<unsigned long long*>(counters + k), 1ULL);
      if (top3) atomicAdd(reinterpret_cast<unsigned long long*>(counters + K + k), 1ULL);
      if ((int)targets[(long long)k * B + i] == null_index) {
        if (top1) atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2 * K + 4 + k), 1ULL);
        atomicAdd(reinterpret_cast<unsigned long long*>(counters + 3 * K + 4 + k), 1ULL);
      }
      chain = chain && top1;
      if (k <= highest) partial = partial && top1;
    }
    if (chain) atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2 * K), 1ULL);
    if (highest >= 0) {
      if (partial) atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2 * K + 1), 1ULL);
      atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2 * K + 2), 1ULL);
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2 * K + 3), 1ULL);
  }
}

// softmax + top-k of every head for the whole batch (inference post-processing, SURVEY.md 8(f) N4;
// R/inference/handler.py:186-214 does softmax -> topk -> 2k .item() per sample and task).  One warp per (task, sample):
// max and sum of exponentials, then kk selection passes in (value descending, index ascending) order - the row is at most a
// few KB and stays in L1.  Slots past min(kk, C_k) get index -1 / probability 0.
template <typename T>
__global__ void __launch_bounds__(128) hier_topk_kernel(const T* __restrict__ logits, long long ld, int B, int K, MetricsMeta meta, int kk,
                                                        int* __restrict__ idx_out, float* __restrict__ prob_out) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)K * B) return;
  const int k = (int)(w / B), i = (int)(w % B);
  const int off = meta.off[k], C = meta.off[k + 1] - off;
  const T* z = logits + (long long)i * ld + off;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, to_f32(z[c]));
  mx = warp_max(mx);
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(to_f32(z[c]) - mx);
  se = warp_sum(se);
  float pv = INFINITY;
  int pi = -1;
  for (int t = 0; t < kk; ++t) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    if (t < C) {
      for (int c = lane; c < C; c += 32) {
        const float v = to_f32(z[c]);
        const bool after = v < pv || (v == pv && c > pi);  // not yet selected
        if (after && (v > bv || (v == bv && c < bi))) bv = v, bi = c;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
      }
    }
    if (lane == 0) {
      const bool ok = t < C && bi != 0x7fffffff;
      idx_out[w * kk + t] = ok ? bi : -1;
      prob_out[w * kk + t] = ok ? expf(bv - mx) / se : 0.f;
    }
    pv = bv;
    pi = bi;
  }
}

}  // namespace

// Parent / child consistency of the per-rank top-1 predictions of a batch, in place on the output of hier_topk_kernel
// (R/inference/postprocessing.py:14-171, one sample at a time there).  Task 0 is the lowest rank, task K - 1 the highest; thread =
// sample, walking down from the highest rank: the highest rank is kept; below it, if the parent rank's consistent prediction is its
// null class, or the tree parent of this rank's top-1 class is not the parent rank's consistent prediction, the rank's prediction
// list becomes the single entry (null class, probability 1) -- when the rank has a null class; otherwise it is kept as is.
struct ConsistencyMeta {
  int off[MAXK + 1];
  int null_idx[MAXK];  // class index of the null taxon of each task, -1 if the task has none
};
__global__ void __launch_bounds__(128) hier_consistency_kernel(int* __restrict__ idx, float* __restrict__ prob, const int* __restrict__ parent,
                                                               ConsistencyMeta meta, unsigned char* __restrict__ changed, int B, int K, int kk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  int cons = -1;  // consistent class index of the rank above
  for (int k = K - 1; k >= 0; --k) {
    int* row = idx + ((long long)k * B + i) * kk;
    float* prow = prob + ((long long)k * B + i) * kk;
    const int cur = row[0];
    bool nullify = false;
    if (k < K - 1) {
      const bool parent_is_null = meta.null_idx[k + 1] >= 0 && cons == meta.null_idx[k + 1];
      if (parent_is_null) {
        nullify = true;
      } else {
        const int C = meta.off[k + 1] - meta.off[k];
        const int actual = (cur >= 0 && cur < C) ? parent[meta.off[k] + cur] : -1;
        nullify = actual != cons;
      }
    }
    if (nullify && meta.null_idx[k] >= 0) {
      row[0] = meta.null_idx[k];
      prow[0] = 1.0f;
      for (int t = 1; t < kk; ++t) row[t] = -1, prow[t] = 0.f;
      cons = meta.null_idx[k];
      if (changed) changed[(long long)k * B + i] = 1;
    } else {
      cons = cur;  // consistent, the highest rank, or a rank that cannot be nullified: its own prediction stands
      if (changed) changed[(long long)k * B + i] = 0;
    }
  }
}

extern "C" int lnx_hier_consistency(int* idx, float* prob, const int* parent, const int* class_off, const int* null_idx, unsigned char* changed,
                                    int B, int K, int kk, lnx_stream_t s) {
  LNX_REQUIRE(idx && prob && parent && class_off && null_idx, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && K > 0 && K <= MAXK && kk > 0, LNX_ERR_SHAPE);
  ConsistencyMeta meta;
  for (int k = 0; k <= K; ++k) meta.off[k] = class_off[k];
  for (int k = 0; k < K; ++k) {
    LNX_REQUIRE(meta.off[k + 1] > meta.off[k], LNX_ERR_SHAPE);
    LNX_REQUIRE(null_idx[k] < meta.off[k + 1] - meta.off[k], LNX_ERR_SHAPE);
    meta.null_idx[k] = null_idx[k];
  }
  hier_consistency_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)s>>>(idx, prob, parent, meta, changed, B, K, kk);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_hier_topk(const void* logits, int dtype, int64_t ld, int B, int K, const int* class_off, int kk, int* idx_out,
                             float* prob_out, lnx_stream_t s) {
  LNX_REQUIRE(logits && class_off && idx_out && prob_out, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && K > 0 && K <= MAXK && kk > 0 && kk <= 64, LNX_ERR_SHAPE);
  MetricsMeta meta;
  for (int k = 0; k <= K; ++k) meta.off[k] = class_off[k];
  for (int k = 0; k < K; ++k) LNX_REQUIRE(meta.off[k + 1] > meta.off[k], LNX_ERR_SHAPE);
  LNX_REQUIRE(ld >= meta.off[K], LNX_ERR_SHAPE);
  const long long warps = (long long)K * B;
  const unsigned grid = (unsigned)((warps + 3) / 4);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_F32)
    hier_topk_kernel<float><<<grid, 128, 0, st>>>((const float*)logits, ld, B, K, meta, kk, idx_out, prob_out);
  else if (dtype == LNX_BF16)
    hier_topk_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)logits, ld, B, K, meta, kk, idx_out, prob_out);
  else
    return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_hier_metrics(const void* logits, int dtype, int64_t ld, int B, int K, const int* class_off, const int64_t* targets,
                                int null_index, int* ranks_out, int64_t* counters, lnx_stream_t s) {
  LNX_REQUIRE(logits && class_off && targets, LNX_ERR_NULL);
  LNX_REQUIRE(ranks_out || counters, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && K > 0 && K <= MAXK, LNX_ERR_SHAPE);
  MetricsMeta meta;
  for (int k = 0; k <= K; ++k) meta.off[k] = class_off[k];
  for (int k = 0; k < K; ++k) LNX_REQUIRE(meta.off[k + 1] > meta.off[k] && meta.off[k + 1] - meta.off[k] < (1 << 24), LNX_ERR_SHAPE);
  LNX_REQUIRE(ld >= meta.off[K], LNX_ERR_SHAPE);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_F32)
    hier_metrics_kernel<float><<<B, 128, 0, st>>>((const float*)logits, ld, B, K, meta, (const long long*)targets, null_index, ranks_out, (long long*)counters);
  else if (dtype == LNX_BF16)
    hier_metrics_kernel<bf16><<<B, 128, 0, st>>>((const bf16*)logits, ld, B, K, meta, (const long long*)targets, null_index, ranks_out, (long long*)counters);
  else
    return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
