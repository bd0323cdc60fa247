// fp32-accumulating CUDA-core GEMM with the fused epilogue of lnx_gemm().
// This is the 1e-4 parity path (LNX_F32) and the small/odd-shape path of the
// bf16 mode (K pitch not TMA-able, e.g. the metadata Linear with K = 2/3/10).
// 128x128x16 tiles, 256 threads, 8x8 register micro-tiles, fp32 shared tiles.
#include "lnx_common.cuh"
#include "lnx_gemm.cuh"

using namespace lnx;

namespace {

constexpr int BM = 128, BN = 128, BK = 16;

// load 8 consecutive elements starting at p[0] (n_valid of them in range), as fp32
template <typename T>
__device__ __forceinline__ void load8(const T* p, int n_valid, bool vec_ok, float* out) {
  if (n_valid >= 8 && vec_ok) {
    if (sizeof(T) == 4) {
      const float4 a = *reinterpret_cast<const float4*>(p);
      const float4 b = *reinterpret_cast<const float4*>(p + 4);
      out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
      out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
    } else {
      const uint4 raw = *reinterpret_cast<const uint4*>(p);
      const bf16* h = reinterpret_cast<const bf16*>(&raw);
#pragma unroll
      for (int i = 0; i < 8; ++i) out[i] = __bfloat162float(h[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = (i < n_valid) ? to_f32(p[i]) : 0.f;
  }
}

// Fill S[BK][BMN] (fp32) with operand(mn0 + i, k0 + kk).
// TRANS = false: element (mn,k) at base[mn*ld + k];  true: at base[k*ld + mn]
template <typename T, bool TRANS>
__device__ __forceinline__ void load_operand(float (*S)[BM + 4], const T* __restrict__ base, long long ld, int mn0, int k0, int MN,
                                             int Kend, bool vec_ok) {
  const int t = threadIdx.x;
  float v[8];
  if (!TRANS) {
    const int r = t & 127, kofs = (t >> 7) * 8;
    const int mn = mn0 + r, k = k0 + kofs;
    const int nv = (mn < MN) ? max(0, min(8, Kend - k)) : 0;
    if (nv > 0) load8(base + (long long)mn * ld + k, nv, vec_ok, v);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) S[kofs + i][r] = v[i];
  } else {
    const int kk = t >> 4, m8 = (t & 15) * 8;
    const int k = k0 + kk, mn = mn0 + m8;
    const int nv = (k < Kend) ? max(0, min(8, MN - mn)) : 0;
    if (nv > 0) load8(base + (long long)k * ld + mn, nv, vec_ok, v);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    *reinterpret_cast<float4*>(&S[kk][m8]) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(&S[kk][m8 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

template <typename TA, typename TC, bool A_TRANS, bool B_TRANS>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs g, int k_per_split, bool a_vec, bool b_vec) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const TA* A = reinterpret_cast<const TA*>(g.A);
  const TA* Bm = reinterpret_cast<const TA*>(g.B);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(g.K, kbeg + k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    load_operand<TA, A_TRANS>(As, A, g.lda, m0, k0, g.M, kend, a_vec);
    load_operand<TA, B_TRANS>(Bs, Bm, g.ldb, n0, k0, g.N, kend, b_vec);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  TC* C = reinterpret_cast<TC*>(g.C);
  TC* aux = reinterpret_cast<TC*>(g.aux_out);
  const TC* agi = reinterpret_cast<const TC*>(g.act_grad_in);
  const TC* res = reinterpret_cast<const TC*>(g.residual);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      const long long idx = (long long)m * g.N + n;
      float v = acc[i][j];
      if (g.accumulate) {
        atomicAdd(reinterpret_cast<float*>(g.C) + idx, v);
        continue;
      }
      v = gemm_epilogue_scalar<TC>(v, m, n, idx, g, aux, agi, res);
      C[idx] = from_f32<TC>(v);
    }
  }
}


// ---- skinny shapes (the metadata heads: Linear(2 / 3 / 10 -> D) on a [B, 15] fp32 tensor).  The 128 x 128 tile above
// spends ~45 us on them (one serial k-step, a 97 % empty tile, 64 dependent bias-load/store pairs per thread); these two
// kernels are plain per-output loops.
// C[m, n..n+3] = epilogue(sum_k A[m, k] * B[n, k]), K <= 16, no transposes
template <typename TA, typename TC>
__global__ void __launch_bounds__(256) gemm_small_k_kernel(const GemmArgs g) {
  const int n4 = (g.N + 3) / 4;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)g.M * n4) return;
  const int m = (int)(t / n4), n0 = (int)(t % n4) * 4;
  const TA* a = reinterpret_cast<const TA*>(g.A) + (long long)m * g.lda;
  const TA* b = reinterpret_cast<const TA*>(g.B) + (long long)n0 * g.ldb;
  float av[16], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 16; ++k) av[k] = k < g.K ? to_f32(a[k]) : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (n0 + j < g.N) {
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < g.K) acc[j] = fmaf(av[k], to_f32(b[(long long)j * g.ldb + k]), acc[j]);
    }
  }
  TC* C = reinterpret_cast<TC*>(g.C);
  TC* aux = reinterpret_cast<TC*>(g.aux_out);
  const TC* agi = reinterpret_cast<const TC*>(g.act_grad_in);
  const TC* res = reinterpret_cast<const TC*>(g.residual);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (n0 + j < g.N) {
      const long long idx = (long long)m * g.N + n0 + j;
      C[idx] = from_f32<TC>(gemm_epilogue_scalar<TC>(acc[j], m, n0 + j, idx, g, aux, agi, res));
    }
  }
}

// C[m, n] (+)= sum_k A[k, m] * B[k, n], N <= 16, both operands "transposed" (the weight gradient dy^T x of a skinny Linear):
// a thread owns one m (coalesced reads of A across the block), the K range is split over blockIdx.y and added with atomics
template <typename TA>
__global__ void __launch_bounds__(128) gemm_skinny_n_kernel(const GemmArgs g, int k_per_split, int atomic) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int kbeg = blockIdx.y * k_per_split, kend = min(g.K, kbeg + k_per_split);
  if (m >= g.M) return;
  const TA* A = reinterpret_cast<const TA*>(g.A) + m;
  const TA* B = reinterpret_cast<const TA*>(g.B);
  float acc[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) acc[n] = 0.f;
#pragma unroll 4
  for (int k = kbeg; k < kend; ++k) {
    const float a = to_f32(A[(long long)k * g.lda]);
#pragma unroll
    for (int n = 0; n < 16; ++n)
      if (n < g.N) acc[n] = fmaf(a, to_f32(B[(long long)k * g.ldb + n]), acc[n]);
  }
  float* C = reinterpret_cast<float*>(g.C) + (long long)m * g.N;
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    if (n < g.N) {
      if (atomic) atomicAdd(C + n, acc[n]);
      else C[n] = acc[n];
    }
  }
}

template <typename TA, typename TC>
int launch(const GemmArgs& g, cudaStream_t st) {
  if (!g.a_trans && !g.b_trans && g.K <= 16 && !g.accumulate) {
    const long long total = (long long)g.M * ((g.N + 3) / 4);
    gemm_small_k_kernel<TA, TC><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g);
    LNX_CHECK_LAUNCH();
    return LNX_OK;
  }
  if (g.a_trans && g.b_trans && g.N <= 16 && sizeof(TC) == 4 && !g.bias && !g.act && !g.aux_out && !g.act_grad_in && !g.residual &&
      !g.col_scale && !g.row_scale) {
    const int bx = (g.M + 127) / 128;
    int splits = g.accumulate ? max(1, min((g.K + 31) / 32, (kNumSMs + bx - 1) / bx)) : 1;
    const int kps = (g.K + splits - 1) / splits;
    splits = (g.K + kps - 1) / kps;
    gemm_skinny_n_kernel<TA><<<dim3(bx, splits), 128, 0, st>>>(g, kps, g.accumulate);
    LNX_CHECK_LAUNCH();
    return LNX_OK;
  }
  const int gx = (g.N + BN - 1) / BN, gy = (g.M + BM - 1) / BM;
  int splits = 1;
  if (g.accumulate) {
    const int tiles = gx * gy;
    splits = max(1, min((g.K + 511) / 512, (kNumSMs * 2 + tiles - 1) / tiles));
  }
  int kps = (g.K + splits - 1) / splits;
  kps = ((kps + BK - 1) / BK) * BK;
  splits = (g.K + kps - 1) / kps;
  dim3 grid(gx, gy, splits);
  const int V = sizeof(TA) == 4 ? 4 : 8;
  const bool a_vec = (g.lda % V == 0) && lnx_aligned16(g.A);
  const bool b_vec = (g.ldb % V == 0) && lnx_aligned16(g.B);
#define LNX_SG(AT, BT) gemm_simt_kernel<TA, TC, AT, BT><<<grid, 256, 0, st>>>(g, kps, a_vec, b_vec)
  if (!g.a_trans && !g.b_trans) LNX_SG(false, false);
  else if (!g.a_trans && g.b_trans) LNX_SG(false, true);
  else if (g.a_trans && !g.b_trans) LNX_SG(true, false);
  else LNX_SG(true, true);
#undef LNX_SG
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

}  // namespace

int lnx_gemm_simt(const GemmArgs& g, int ab_dtype, int c_dtype, cudaStream_t st) {
  if (ab_dtype == LNX_F32 && c_dtype == LNX_F32) return launch<float, float>(g, st);
  if (ab_dtype == LNX_F32 && c_dtype == LNX_BF16) return launch<float, bf16>(g, st);
  if (ab_dtype == LNX_BF16 && c_dtype == LNX_BF16) return launch<bf16, bf16>(g, st);
  if (ab_dtype == LNX_BF16 && c_dtype == LNX_F32) return launch<bf16, float>(g, st);
  return LNX_ERR_DTYPE;
}
