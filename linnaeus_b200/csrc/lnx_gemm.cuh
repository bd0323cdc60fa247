// Shared GEMM argument block and the scalar epilogue used by both GEMM kernels.
#pragma once
#include "lnx_common.cuh"

struct GemmArgs {
  const void* A;
  const void* B;
  void* C;
  long long lda, ldb;
  int M, N, K;
  int a_trans, b_trans;
  const float* bias;
  int act;
  void* aux_out;
  const void* act_grad_in;
  const void* residual;
  const float* col_scale;
  const float* row_scale;  // per row-group multiplier (DropPath): row_scale[m / rows_per_group]
  int rows_per_group;
  int accumulate;
  // Token-position factors (the reference's cos "RoPE", rope_2d_mhsa.py:432-501, folded into the qkv projection; tcgen05 path only):
  // columns n < 2 tok_dim of row m are multiplied by cos(tx fx[p] + ty fy[p]), p = (n % tok_dim) / 2, fx = tok_scale[p],
  // fy = tok_scale[tok_dim / 2 + p], (tx, ty) = (pos % tok_w, pos / tok_w), pos = m % tok_period - tok_extra (factor 1 for the
  // tok_extra leading tokens of every sequence); columns n < tok_dim also by tok_qscale.
  const float* tok_scale = nullptr;  // the learnable frequencies [2, tok_dim / 2]
  int tok_period = 1, tok_extra = 0, tok_dim = 0, tok_w = 1;
  float tok_qscale = 1.f;
};

// v = acc -> epilogue value for element (m, n) at flat index idx (pitch N)
template <typename TC>
__device__ __forceinline__ float gemm_epilogue_scalar(float v, int m, int n, long long idx, const GemmArgs& g, TC* aux, const TC* agi,
                                                      const TC* res) {
  using namespace lnx;
  if (g.bias) v += g.bias[n];
  if (aux) aux[idx] = from_f32<TC>(g.act == LNX_ACT_GELU_DG ? gelu_grad_f(v) : v);
  if (agi) {
    const float u = to_f32(agi[idx]);
    if (g.act == LNX_ACT_GELU) v *= gelu_grad_f(u);
    else if (g.act == LNX_ACT_RELU) v = (u > 0.f) ? v : 0.f;
    else if (g.act == LNX_ACT_MUL) v *= u;
  } else {
    if (g.act == LNX_ACT_GELU || g.act == LNX_ACT_GELU_DG) v = gelu_f(v);
    else if (g.act == LNX_ACT_RELU) v = fmaxf(v, 0.f);
    else if (g.act == LNX_ACT_SWISH) v = v / (1.f + expf(-v));
  }
  if (g.col_scale) v *= g.col_scale[n];
  if (g.row_scale) v *= g.row_scale[m / g.rows_per_group];
  if (res) v += to_f32(res[idx]);
  return v;
}

int lnx_gemm_simt(const GemmArgs& g, int ab_dtype, int c_dtype, cudaStream_t st);
// returns LNX_ERR_UNSUPPORTED when the shape cannot run on the tcgen05 kernel
int lnx_gemm_tc(const GemmArgs& g, int c_dtype, cudaStream_t st);
