// Batch augmentation that feeds the model: the apply step of selective mixup.
//
// R/aug/gpu/selective_mixup.py:77-330 draws (a) an in-group permutation, (b) one lambda ~ Beta(alpha, alpha), (c) one
// uniform number per sample, and then
//   images / soft targets : out[i] = lam * x[i] + (1 - lam) * x[perm[i]]                                  (:150, :177)
//   metadata              : every chunk with any zero entry is zeroed and marked invalid, in place (:371-392); then per
//                           sample and chunk the original or the partner chunk is taken whole: both non-zero -> original
//                           iff pick[i] < 0.5, exactly one non-zero -> that one, both zero -> zeros / invalid (:394-560),
//                           a Python loop with two host syncs per (sample, chunk).
// Here the draws stay on the device (lam and pick are device pointers: no host sync) and the apply step is three launches.
// lnx_mix_pairs is HBM bound: 2 reads + 1 write of the batch (the partner row is a gather of whole rows, fully coalesced).
// Arithmetic is mul, mul, add in fp32 with no contraction, so results equal the reference's expression bit for bit.
#include "lnx_common.cuh"

using namespace lnx;

namespace {

constexpr int MAX_CHUNKS = 16;
struct ChunkBounds {
  int n;
  int lo[MAX_CHUNKS], hi[MAX_CHUNKS];
};

// grid (chunks of a row, samples)
template <int V>
__global__ void __launch_bounds__(256) mix_pairs_kernel(const float* __restrict__ x, const long long* __restrict__ perm, const float* __restrict__ lam_p,
                                                        float* __restrict__ out, long long row_v) {
  const float lam = *lam_p;
  const float oml = __fsub_rn(1.0f, lam);
  const long long i = blockIdx.y, j = perm[i];
  const float* a = x + i * row_v * V;
  const float* b = x + j * row_v * V;
  float* o = out + i * row_v * V;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < row_v; e += (long long)gridDim.x * blockDim.x) {
    if (V == 4) {
      const float4 va = __ldg(reinterpret_cast<const float4*>(a) + e), vb = __ldg(reinterpret_cast<const float4*>(b) + e);
      float4 vo;
      vo.x = __fadd_rn(__fmul_rn(lam, va.x), __fmul_rn(oml, vb.x));
      vo.y = __fadd_rn(__fmul_rn(lam, va.y), __fmul_rn(oml, vb.y));
      vo.z = __fadd_rn(__fmul_rn(lam, va.z), __fmul_rn(oml, vb.z));
      vo.w = __fadd_rn(__fmul_rn(lam, va.w), __fmul_rn(oml, vb.w));
      reinterpret_cast<float4*>(o)[e] = vo;
    } else {
      o[e] = __fadd_rn(__fmul_rn(lam, __ldg(a + e)), __fmul_rn(oml, __ldg(b + e)));
    }
  }
}

// one thread per (sample, chunk): zero the chunk and clear its validity when any entry is exactly zero
__global__ void meta_enforce_kernel(float* __restrict__ aux, unsigned char* __restrict__ mask, int B, int D, ChunkBounds cb) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * cb.n) return;
  const int i = t / cb.n, c = t - i * cb.n;
  float* a = aux + (long long)i * D;
  bool partial = false;
  for (int d = cb.lo[c]; d < cb.hi[c]; ++d) partial |= (a[d] == 0.0f);
  if (partial) {
    for (int d = cb.lo[c]; d < cb.hi[c]; ++d) {
      a[d] = 0.0f;
      mask[(long long)i * D + d] = 0;
    }
  }
}

// one thread per (sample, chunk): take the original or the partner chunk whole
__global__ void meta_pick_kernel(const float* __restrict__ aux, const unsigned char* __restrict__ mask, const long long* __restrict__ perm,
                                 const float* __restrict__ pick, float* __restrict__ out_aux, unsigned char* __restrict__ out_mask, int B, int D,
                                 ChunkBounds cb) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * cb.n) return;
  const int i = t / cb.n, c = t - i * cb.n;
  const long long j = perm[i];
  const float* a1 = aux + (long long)i * D;
  const float* a2 = aux + j * D;
  bool z1 = true, z2 = true;
  for (int d = cb.lo[c]; d < cb.hi[c]; ++d) {
    z1 &= (a1[d] == 0.0f);
    z2 &= (a2[d] == 0.0f);
  }
  int src;  // 0: zeros, 1: original, 2: partner
  if (!z1 && !z2) src = pick[i] < 0.5f ? 1 : 2;
  else if (!z1) src = 1;
  else if (!z2) src = 2;
  else src = 0;
  const long long s = src == 2 ? j : i;
  for (int d = cb.lo[c]; d < cb.hi[c]; ++d) {
    out_aux[(long long)i * D + d] = src ? aux[s * D + d] : 0.0f;
    out_mask[(long long)i * D + d] = src ? mask[s * D + d] : (unsigned char)0;
  }
}

// selective CutMix (R/aug/gpu/selective_cutmix.py:204-273): one box for the whole batch; inside it a grouped sample
// (group id != -1) takes its partner's pixels.  grid (chunks of an image, samples); e runs over C*H*W (float4 when W % 4 == 0)
template <int V>
__global__ void __launch_bounds__(256) cutmix_paste_kernel(const float* __restrict__ x, const long long* __restrict__ perm,
                                                           const long long* __restrict__ gids, float* __restrict__ out, long long img_v, int H, int W,
                                                           int h1, int w1, int h2, int w2) {
  const long long i = blockIdx.y;
  const long long j = gids[i] != -1 ? perm[i] : i;
  const float* a = x + i * img_v * V;
  const float* b = x + j * img_v * V;
  float* o = out + i * img_v * V;
  const int wv = W / V;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < img_v; e += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(e % wv) * V;
    const int h = (int)((e / wv) % H);
    const bool row_in = j != i && h >= h1 && h < h2;
    if (V == 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(a) + e);
      if (row_in && w + 3 >= w1 && w < w2) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(b) + e);
        if (w >= w1 && w < w2) v.x = p.x;
        if (w + 1 >= w1 && w + 1 < w2) v.y = p.y;
        if (w + 2 >= w1 && w + 2 < w2) v.z = p.z;
        if (w + 3 >= w1 && w + 3 < w2) v.w = p.w;
      }
      reinterpret_cast<float4*>(o)[e] = v;
    } else {
      o[e] = (row_in && w >= w1 && w < w2) ? __ldg(b + e) : __ldg(a + e);
    }
  }
}

// out[i] = group id != -1 ? ca * x[i] + cb * x[perm[i]] : x[i]   (soft targets under CutMix; coefficients by value)
__global__ void mix_pairs_valid_kernel(const float* __restrict__ x, const long long* __restrict__ perm, const long long* __restrict__ gids, float ca,
                                       float cb, float* __restrict__ out, int B, long long row) {
  const long long total = (long long)B * row;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / row, e = t - i * row;
    const float v = x[t];
    out[t] = gids[i] != -1 ? __fadd_rn(__fmul_rn(ca, v), __fmul_rn(cb, x[perm[i] * row + e])) : v;
  }
}

}  // namespace

extern "C" int lnx_cutmix_paste(const float* x, const int64_t* perm, const int64_t* group_ids, float* out, int B, int C, int H, int W, int h1,
                                int w1, int h2, int w2, lnx_stream_t s) {
  LNX_REQUIRE(x && perm && group_ids && out, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && B <= 65535 && C > 0 && H > 0 && W > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(x != out, LNX_ERR_UNSUPPORTED);
  cudaStream_t st = (cudaStream_t)s;
  const bool vec = W % 4 == 0 && lnx_aligned16(x) && lnx_aligned16(out);
  const long long img = (long long)C * H * W, img_v = vec ? img / 4 : img;
  const int bx = (int)max(1LL, min((img_v + 255) / 256, (long long)(8 * kNumSMs + B - 1) / B));
  dim3 grid(bx, B);
  if (vec) cutmix_paste_kernel<4><<<grid, 256, 0, st>>>(x, (const long long*)perm, (const long long*)group_ids, out, img_v, H, W, h1, w1, h2, w2);
  else cutmix_paste_kernel<1><<<grid, 256, 0, st>>>(x, (const long long*)perm, (const long long*)group_ids, out, img_v, H, W, h1, w1, h2, w2);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_mix_pairs_valid(const float* x, const int64_t* perm, const int64_t* group_ids, float coef_self, float coef_partner, float* out,
                                   int B, int64_t row, lnx_stream_t s) {
  LNX_REQUIRE(x && perm && group_ids && out, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && row > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(x != out, LNX_ERR_UNSUPPORTED);
  const long long total = (long long)B * row;
  const int blocks = (int)max(1LL, min((total + 255) / 256, (long long)kNumSMs * 16));
  mix_pairs_valid_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(x, (const long long*)perm, (const long long*)group_ids, coef_self, coef_partner, out, B,
                                                              row);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_mix_pairs(const float* x, const int64_t* perm, const float* lam, float* out, int B, int64_t row, lnx_stream_t s) {
  LNX_REQUIRE(x && perm && lam && out, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && row > 0 && B <= 65535, LNX_ERR_SHAPE);
  LNX_REQUIRE(x != out, LNX_ERR_UNSUPPORTED);  // rows are gathered from other samples: not an in-place operation
  cudaStream_t st = (cudaStream_t)s;
  const bool vec = row % 4 == 0 && lnx_aligned16(x) && lnx_aligned16(out);
  const long long row_v = vec ? row / 4 : row;
  const int bx = (int)max(1LL, min((row_v + 255) / 256, (long long)(8 * kNumSMs + B - 1) / B));
  dim3 grid(bx, B);
  if (vec) mix_pairs_kernel<4><<<grid, 256, 0, st>>>(x, (const long long*)perm, lam, out, row_v);
  else mix_pairs_kernel<1><<<grid, 256, 0, st>>>(x, (const long long*)perm, lam, out, row_v);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_mix_meta_chunks(float* aux, uint8_t* mask, const int64_t* perm, const float* pick, const int* chunk_bounds, int n_chunks,
                                   float* out_aux, uint8_t* out_mask, int B, int D, lnx_stream_t s) {
  LNX_REQUIRE(aux && mask && perm && pick && chunk_bounds && out_aux && out_mask, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && D > 0 && n_chunks > 0 && n_chunks <= MAX_CHUNKS, LNX_ERR_SHAPE);
  LNX_REQUIRE(aux != out_aux && mask != out_mask, LNX_ERR_UNSUPPORTED);
  ChunkBounds cb;
  cb.n = n_chunks;
  for (int c = 0; c < n_chunks; ++c) {
    cb.lo[c] = chunk_bounds[2 * c];
    cb.hi[c] = chunk_bounds[2 * c + 1];
    LNX_REQUIRE(cb.lo[c] >= 0 && cb.lo[c] <= cb.hi[c] && cb.hi[c] <= D, LNX_ERR_SHAPE);
  }
  cudaStream_t st = (cudaStream_t)s;
  const int threads = 128, blocks = (B * n_chunks + threads - 1) / threads;
  meta_enforce_kernel<<<blocks, threads, 0, st>>>(aux, mask, B, D, cb);
  LNX_CHECK_LAUNCH();
  // entries outside every chunk keep the reference's torch.empty_like semantics: they are left to the caller (the host
  // mirror copies them through); inside chunks every element is written
  meta_pick_kernel<<<blocks, threads, 0, st>>>(aux, mask, (const long long*)perm, pick, out_aux, out_mask, B, D, cb);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
