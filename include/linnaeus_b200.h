/* linnaeus_b200 -- C ABI of the B200 (sm_100a) mFormer hot-path kernels.
 *
 * The reference (polli-labs/linnaeus) is pure Python/PyTorch and has no native
 * interface of its own; each entry point below replaces the chain of PyTorch
 * library calls at the cited reference location (R/ = /root/reference/linnaeus).
 * The Python host layer (linnaeus_b200/functional.py) binds these with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Contract ("C-ABI layer"):
 *  - plain pointers and sizes only; every pointer is DEVICE memory owned by the
 *    caller; the library never allocates, frees or retains pointers;
 *  - every launch goes to the passed stream, no hidden synchronisation, CUDA-graph
 *    capturable;
 *  - returns 0 (LNX_OK) or a negative error code, never aborts or throws.
 *  - dtype: LNX_F32 = 0, LNX_BF16 = 1 (activations); parameters, statistics,
 *    gradients of parameters are always float32.
 *  - "+=" in a comment means the kernel atomically accumulates into the buffer.
 */
#ifndef LINNAEUS_B200_H
#define LINNAEUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* lnx_stream_t; /* cudaStream_t */

enum { LNX_F32 = 0, LNX_BF16 = 1 };
enum {
  LNX_OK = 0,
  LNX_ERR_SHAPE = -1,
  LNX_ERR_DTYPE = -2,
  LNX_ERR_ALIGN = -3,
  LNX_ERR_CUDA = -4,
  LNX_ERR_UNSUPPORTED = -5,
  LNX_ERR_NULL = -6
};
/* LNX_ACT_GELU_DG: forward variant that stores gelu'(pre) (not pre) in aux_out, so the backward needs no
 * transcendental: LNX_ACT_MUL multiplies the accumulator by act_grad_in (the stored derivative). */
enum { LNX_ACT_NONE = 0, LNX_ACT_GELU = 1, LNX_ACT_RELU = 2, LNX_ACT_GELU_DG = 3, LNX_ACT_MUL = 4, LNX_ACT_SWISH = 5 };
enum { LNX_LOSS_CE = 0, LNX_LOSS_LABEL_SMOOTHING = 1, LNX_LOSS_TAXONOMY = 2 };

int lnx_version(void);
const char* lnx_strerror(int code);

/* ---- layout ------------------------------------------------------------ */
/* Stem im2col: x f32 NCHW [B,Cin,H,W] -> out [B*(H/p)*(W/p), Kpad], column order
 * (c,kh,kw) = Conv2d weight.flatten(1), zero padded to Kpad.
 * Replaces the input side of nn.Conv2d(k=4,s=4), R/models/mFormerV1.py:145-148. */
int lnx_patchify_nchw(const float* x, void* out, int B, int Cin, int H, int W, int p, int Kpad, int out_dtype, lnx_stream_t s);

/* NHWC [B,H,W,C] <-> [B*(H/2)*(W/2), 4C] with column order (kh,kw,c) (inverse=1
 * is the backward scatter).  Input side of the 2x2/s2 downsample conv,
 * R/models/blocks/convnext.py:104-115. */
int lnx_space_to_depth(const void* x, void* out, int B, int H, int W, int C, int inverse, int dtype, lnx_stream_t s);

/* tokens[b] = cat(cls[b or broadcast], extras[b], patches[b]) (torch.cat,
 * R/models/mFormerV1.py:446-465,479-504).  cls_stride = 0 broadcasts one row;
 * extras may be NULL with n_meta > 0 (rows are zero filled; used by the
 * backward of lnx_tokens_split). */
int lnx_tokens_assemble(const void* cls, int64_t cls_stride, const void* extras, const void* patches, void* tokens,
                        int B, int n_meta, int n_patch, int D, int dtype, lnx_stream_t s);
/* inverse: any of cls_out [B,D] / extras_out [B,n_meta,D] / patches_out [B,n_patch,D] may be NULL */
int lnx_tokens_split(const void* tokens, void* cls_out, void* extras_out, void* patches_out,
                     int B, int n_meta, int n_patch, int D, int dtype, lnx_stream_t s);

/* out[n] += sum_m x[m,n]  (bias / broadcast-parameter gradients) */
int lnx_colsum(const void* x, float* out, int64_t M, int N, int dtype, lnx_stream_t s);
int lnx_cast_f32_to_bf16(const float* in, void* out, int64_t n, lnx_stream_t s);
/* out[m,:] = x[m,:] * s[m / rows_per_group]  (DropPath backward: per-sample mask on the branch gradient) */
int lnx_rowscale(const void* x, const float* s, void* out, int64_t M, int N, int rows_per_group, int dtype, lnx_stream_t st);
/* ConvNeXt layer scale y = gamma * (h W2^T + b2) (R/models/blocks/convnext.py:84): out[n,:] = bf16(w[n,:] * cs[n]) is the weight
 * with gamma folded in for the data-gradient GEMM; lnx_layerscale_bwd turns the UN-scaled gradients into the parameter gradients
 * dw[N,K] += cs dw_raw, db[N] += cs db_raw, dcs[N] += rowsum(dw_raw * w) + b db_raw (db_raw / b / db may be NULL). */
int lnx_rowscale_cast_bf16(const float* w, const float* cs, void* out, int N, int K, lnx_stream_t s);
int lnx_layerscale_bwd(const float* dw_raw, const float* db_raw, const float* w, const float* b, const float* cs, float* dw, float* db,
                       float* dcs, int N, int K, lnx_stream_t s);
/* out = dy * act'(pre)  (single-layer Linear+activation backward; act = LNX_ACT_GELU | LNX_ACT_RELU) */
int lnx_act_bwd(const void* dy, const void* pre, void* out, int64_t n, int act, int dtype, lnx_stream_t s);

/* ---- normalisation ----------------------------------------------------- */
/* y = LN(x) * w + b (+ residual); biased variance, fp32 statistics saved in
 * mean/rstd [rows].  nn.LayerNorm (R/models/blocks/convnext.py:77-78,
 * rope_2d_mhsa.py:546-547,605,634) and LayerNormChannelsFirst (convnext.py:32-43:
 * the conv trunk is kept NHWC so channels-first LN is a row LN). */
int lnx_layernorm_fwd(const void* x, const float* w, const float* b, const void* residual, void* y, float* mean, float* rstd,
                      int64_t rows, int C, float eps, int dtype, lnx_stream_t s);
/* dx = LN backward; dw[C] += , db[C] += */
int lnx_layernorm_bwd(const void* dy, const void* x, const float* w, const float* mean, const float* rstd, const void* dres, void* dx,
                      float* dw, float* db, int64_t rows, int C, int dtype, lnx_stream_t s);

/* ---- depthwise 7x7 (pad 3) on NHWC ------------------------------------- */
/* w_layout: LNX_DW_W_TAP_MAJOR = [49, C]; LNX_DW_W_NATIVE = the Conv2d weight itself, [C,1,7,7] = [C, 49]; LNX_DW_W_NATIVE_FLIPPED =
 * the same memory read with the taps reversed.  bias may be NULL.  The data gradient is the same call on dY with the taps reversed
 * and bias = NULL; `residual` (nullable) is added to the result (the skip-connection gradient when run as the data gradient).
 * bf16: the default kernels run on the tensor pipe (banded matrix products, lnx_dwconv_mma.cu): the weights are rounded to bf16 for
 * the multiply (as the reference's autocast Conv2d does), accumulation is fp32; a non-finite input spreads over the 16-column K
 * window it falls in (never to another channel or image).  fp32 and lnx_dwconv7_set_impl(0): fp32 weights, exact 7 x 7 support.
 * R/models/blocks/convnext.py:56-58,76. */
enum { LNX_DW_W_TAP_MAJOR = 0, LNX_DW_W_NATIVE = 1, LNX_DW_W_NATIVE_FLIPPED = 2 };
int lnx_dwconv7_fwd(const void* x, const float* w, int w_layout, const float* bias, const void* residual, void* y, int B, int H, int W, int C,
                    int dtype, lnx_stream_t s);
/* dw += (layout LNX_DW_W_TAP_MAJOR or LNX_DW_W_NATIVE: the native layout lets the kernel accumulate straight into the parameter's
 * gradient buffer), dbias[C] += */
int lnx_dwconv7_wgrad(const void* x, const void* dy, float* dw, int w_layout, float* dbias, int B, int H, int W, int C, int dtype,
                      lnx_stream_t s);
/* Which bf16 kernels the two calls above run: 1 = tensor pipe (banded matrix products on mma.sync, lnx_dwconv_mma.cu; the default),
 * 0 = fp32x2 FMA kernels (lnx_dwconv_bf16.cu; also the fallback for shapes the tensor-pipe kernels do not cover), -1 = back to the
 * default / the LNX_DWCONV_MMA environment variable.  Returns the previous setting.  For A/B measurements and parity tests of one
 * implementation against the other; no reference counterpart. */
int lnx_dwconv7_set_impl(int impl);

/* ---- GEMM with fused epilogue ------------------------------------------ */
/* Weight (+ bias) gradient of y = x W^T + b on the tensor cores, reduction over the M rows (tokens):
 *   dw[N,K] += dy[M,N]^T x[M,K],   db[N] += sum_m dy[m,:]   (db may be NULL).
 * dy / x are bf16 row-major with pitches ldy / ldx (elements); split-K over all SMs, db from an extra
 * N = 16 MMA against a tile of ones.  Replaces autograd's Linear weight/bias backward
 * (torch.nn.functional.linear backward; R/models/blocks/{convnext.py:79-86, mlp.py:61-65, rope_2d_mhsa.py:292-294}). */
int lnx_wgrad(const void* dy, int64_t ldy, const void* x, int64_t ldx, float* dw, float* db, int64_t M, int N, int K, int dtype,
              lnx_stream_t s);

/* ---- fused ConvNeXt pointwise pair ---------------------------------------- */
/* y[M,C] = residual + row_scale[m / rows_per_group] * gamma * (gelu(x W1^T + b1) W2^T + b2), all activations bf16.
 * One tcgen05 kernel: the [M, H = 4C] hidden tensor stays in TMEM (the GELU output is written back into the
 * accumulator's columns and read by the second MMA as its A operand).  gamma, row_scale, residual, b1, b2 may be
 * NULL.  C in {96, 192}; anything else returns LNX_ERR_UNSUPPORTED (the caller then runs two lnx_gemm calls).
 * Replaces pwconv1 -> GELU -> pwconv2 -> gamma -> DropPath -> + input, R/models/blocks/convnext.py:79-86. */
int lnx_mlp_fused_fwd(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const float* gamma,
                      const float* row_scale, int rows_per_group, const void* residual, void* y, int64_t M, int C, int H,
                      lnx_stream_t s);
/* Data path of its backward, with the pre-activation recomputed from x (the forward saves only its input):
 *   h[M,H] = gelu(x W1^T + b1),  dpre[M,H] = (dy W2e) * gelu'(x W1^T + b1),  dx[M,C] = dpre W1,
 * W2e = W2 * gamma[:, None] folded by the caller; dy already carries the DropPath row scale.  h and dpre feed the two
 * weight-gradient GEMMs (lnx_wgrad).  C = 96 only (weights resident in shared memory); else LNX_ERR_UNSUPPORTED.
 * Replaces autograd through R/models/blocks/convnext.py:79-86. */
int lnx_mlp_fused_bwd(const void* x, const void* dy, const void* w1, const float* b1, const void* w2e, void* h, void* dpre, void* dx,
                      int64_t M, int C, int H, lnx_stream_t s);
/* The weight gradients of the same pair without any 4C-wide tensor in HBM (pass h = dpre = NULL to lnx_mlp_fused_bwd then):
 * every CTA owns one third of the hidden units, recomputes its slice of the pre-activation and of dY W2e per row tile, and
 * accumulates dw1[H,C] += dpre^T x and dw2_raw[C,H] += dy^T h in TMEM across all its row tiles (one atomic flush at the end);
 * db1[H] += colsum(dpre) (extra MMA columns against a tile of ones).  "raw" = before the layer scale (lnx_layerscale_bwd finishes
 * it; db2_raw = lnx_colsum(dy)).  C = 96. */
int lnx_mlp_fused_wgrad(const void* x, const void* dy, const void* w1, const float* b1, const void* w2e, float* dw1, float* db1,
                        float* dw2_raw, int64_t M, int C, int H, lnx_stream_t s);

/* acc[m,n] = sum_k A(m,k) * B(n,k)
 *   a_trans = 0: A stored [M,K] (row pitch lda); 1: stored [K,M]
 *   b_trans = 0: B stored [N,K] (row pitch ldb); 1: stored [K,N]
 * v = acc + bias[n]; aux_out[m,n] = v (pre-activation, optional); v = act(v);
 * if act_grad_in: v = acc * act'(act_grad_in[m,n]);   (backward through act)
 * v *= col_scale[n]; v *= row_scale[m / rows_per_group] (DropPath mask, optional);
 * v += residual[m,n];  C[m,n] = v   (C, aux, residual pitch = N)
 * colsum_out (nullable, float[N]) += column sums of C (the bias gradient of the layer
 * below, fused into the persistent tensor-core kernel's epilogue).
 * accumulate = 1: C is float32 and receives atomic += (split-K weight gradients).
 * ab_dtype LNX_BF16 runs the tcgen05/TMEM/TMA kernel when the shape allows
 * (K-pitch multiple of 8, 16-byte aligned bases), else the SIMT kernel;
 * LNX_F32 always runs the fp32 SIMT kernel (1e-4 parity mode).
 * Replaces nn.Linear / patchify Conv2d / their autograd (cuBLASLt, cuDNN):
 * R/models/blocks/convnext.py:60-64,79-86,110; rope_2d_mhsa.py:292,294,432,502;
 * mlp.py:37-39,61-65; mFormerV1.py:146,316-323; heads/linear_head.py:33-46. */
int lnx_gemm(int ab_dtype, const void* A, int64_t lda, int a_trans, const void* B, int64_t ldb, int b_trans,
             void* C, int c_dtype, int M, int N, int K,
             const float* bias, int act, void* aux_out, const void* act_grad_in,
             const void* residual, const float* col_scale, const float* row_scale, int rows_per_group, float* colsum_out,
             int accumulate, int force_simt, lnx_stream_t s);

/* ---- 2-D "RoPE" (cos scaling) and attention ------------------ */
/* theta[n,h,j] = tx[n]*freqs[0,h,j] + ty[n]*freqs[1,h,j]; cos/sin tables [H*W,heads,half].
 * R/models/blocks/rope_2d_mhsa.py:56-73,114-155,404-408. */
int lnx_rope_table(const float* freqs, float* cos_out, float* sin_out, int H, int W, int heads, int half, lnx_stream_t s);
/* qkv [B,N,3,heads,hd] -> q,k,v [B,heads,N,hd]; image-token (n >= n_extra) pairs of
 * q,k scaled by cos; q additionally by q_scale.  rope_2d_mhsa.py:432-456. */
int lnx_rope_qk_fwd(const void* qkv, const float* cos_tab, void* q, void* k, void* v,
                    int B, int N, int heads, int hd, int n_extra, float q_scale, int dtype, lnx_stream_t s);
/* dq,dk,dv [B,heads,N,hd] -> dqkv [B,N,3,heads,hd]; dtheta[N_img,heads,hd/2] += */
int lnx_rope_qk_bwd(const void* dq, const void* dk, const void* dv, const void* qkv, const float* cos_tab, const float* sin_tab,
                    void* dqkv, float* dtheta, int B, int N, int heads, int hd, int n_extra, float q_scale, int dtype, lnx_stream_t s);
/* same, when qkv holds the already scaled q cos s / k cos (written by lnx_qkv_rope_gemm; the un-scaled q / k were never stored) */
int lnx_rope_qk_bwd_scaled(const void* dq, const void* dk, const void* dv, const void* qkv_scaled, const float* cos_tab, const float* sin_tab,
                           void* dqkv, float* dtheta, int B, int N, int heads, int hd, int n_extra, float q_scale, int dtype, lnx_stream_t s);
/* dfreqs[2,heads,half] += sum_n (tx[n], ty[n]) * dtheta[n,h,j] */
int lnx_rope_freq_grad(const float* dtheta, float* dfreqs, int H, int W, int heads, int half, lnx_stream_t s);

/* softmax(q k^T) v per (batch, head), fp32 softmax; out [B,N,heads*hd]; lse [B,heads,N].
 * rope_2d_mhsa.py:492-501 (standard path, scale already folded into q). */
int lnx_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse,
                 int B, int heads, int N, int hd, int dtype, int force_simt, lnx_stream_t s);
/* delta_ws: 16-byte aligned float scratch of B*heads*N*(hd+1) + 4 elements (row deltas + fp32 dQ accumulator) */
int lnx_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                 void* dq, void* dk, void* dv, float* delta_ws, int B, int heads, int N, int hd, int dtype, int force_simt, lnx_stream_t s);

/* bf16 GEMM on CTA pairs (tcgen05.mma.cta_group::2, 256 x block_n tiles, each CTA stages half of the B tile): c[M, N] = a[M, K] b[N, K]^T
 * (+ bias).  N % 128 == 0, K % 8 == 0.  The 2-SM building block for the transformer-stage Linear layers (R/models/blocks/mlp.py:37-39,
 * rope_2d_mhsa.py:292-294). */
int lnx_gemm_pair(const void* a, const void* b, const float* bias, void* c, int64_t M, int N, int K, lnx_stream_t s);

/* The qkv projection with the cos factors and the softmax scale applied in its epilogue (replaces self.qkv(x) + the q / k scaling of
 * rope_2d_mhsa.py:432-501): qkv[M, 3 D] bf16 = x[M, K] w[3 D, K]^T + bias, q / k columns of the image tokens times
 * cos(tx fx + ty fy) computed from freqs [2, D / 2] (the learnable frequencies, float32) in the epilogue, q columns times q_scale.
 * M = B * tokens; the image tokens follow the n_extra leading tokens of each sequence and form a grid_w-wide grid.
 * tcgen05 only (bf16): LNX_ERR_UNSUPPORTED otherwise. */
int lnx_qkv_rope_gemm(const void* x, const void* w, const float* bias, const float* freqs, void* qkv, int64_t M, int D, int K,
                      int tokens, int n_extra, int grid_w, float q_scale, lnx_stream_t s);
/* attention reading q / k / v straight from that [B, N, 3, heads, hd] matrix (bf16, hd = 64, N <= 240); out [B,N,heads*hd], lse [B,heads,N] */
int lnx_attn_qkv_fwd(const void* qkv, void* out, float* lse, int B, int heads, int N, int hd, int dtype, lnx_stream_t s);
/* its backward: dq / dk / dv head-major [B,heads,N,hd] */
int lnx_attn_qkv_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dq, void* dk, void* dv,
                     int B, int heads, int N, int hd, int dtype, lnx_stream_t s);

/* ---- tail ---------------------------------------------------------------- */
/* out[b,d] = w[0]*a[b,d] + w[1]*c[b,d] + bias[0]   (Conv1d(2->1,k=1), mFormerV1.py:512-524) */
int lnx_aggregate2_fwd(const void* a, const void* c, const float* w2, const float* bias1, void* out, int64_t B, int D, int dtype, lnx_stream_t s);
/* da, dc; dw2[2] += ; dbias1[1] += */
int lnx_aggregate2_bwd(const void* dout, const void* a, const void* c, const float* w2, void* da, void* dc, float* dw2, float* dbias1,
                       int64_t B, int D, int dtype, lnx_stream_t s);

/* ---- hierarchical masked loss (R/loss/*) ---------------------------------- */
/* logits [B,Ctot] (all K heads concatenated, class_off[K+1] column offsets, host
 * array); targets int64 [K,B]; null_flag uint8 [K,B] or NULL (then null = target==0);
 * keep float [K,B] or NULL: per-sample multiplier (0 drops a null sample = the
 * reference's coin flips; class weights are folded in here by the host);
 * soft_mats: K device pointers ([C_k,C_k] rows) for LNX_LOSS_TAXONOMY, host array.
 * phase1 != 0: every null sample is zeroed (PHASE1_MASK_NULL_LOSS / ignore_index 0).
 * Outputs: per_sample [K,B] (masked), raw [K,B] (unmasked, for logging), lse [K,B],
 * sample_w [K,B] (effective multiplier, consumed by lnx_loss_bwd).
 * basic_loss.py:15-228, taxonomy_label_smoothing.py:219-408, masking.py:19-466. */
int lnx_loss_fwd(const void* logits, int dtype, int B, int K, const int* class_off, const int64_t* targets,
                 const uint8_t* null_flag, const float* keep, int kind, float smoothing, const float* const* soft_mats,
                 int phase1, float* per_sample, float* raw, float* lse, float* sample_w, lnx_stream_t s);
/* total = sum_k w[k] * sum_i l[k,i] / max(nvalid_k, 1e-6); nvalid_k = #(l != 0), or B
 * when phase1 (hierarchical_loss.py:241-276,337-340; gradient_weighting.py:301-358).
 * scale[k] = w[k]/max(nvalid_k,1e-6); task_sum[k] = weighted task loss; nvalid float [K]. */
int lnx_loss_reduce(const float* per_sample, const float* task_w, int B, int K, int phase1,
                    float* total, float* scale, float* task_sum, float* nvalid, lnx_stream_t s);
/* dlogits[i, off_k + c] = gscale[0] * scale[k] * sample_w[k,i] * (sum(T) * softmax_c - T_c) */
int lnx_loss_bwd(const void* logits, int dtype, int B, int K, const int* class_off, const int64_t* targets,
                 const float* sample_w, int kind, float smoothing, const float* const* soft_mats,
                 const float* lse, const float* scale, const float* gscale, void* dlogits, lnx_stream_t s);

/* ---- optimizer (R/train.py:282-313, R/optimizers/build.py:67-106) -------- */
/* sumsq[0] += sum g^2  (call once per flat buffer, zero sumsq first).  workspace: NULL (fp32 atomics: the last bits depend on
 * block scheduling) or LNX_SUMSQ_WORKSPACE floats, zero before the first call: per-block partials added in index order by the last
 * block, i.e. a deterministic norm -> data-parallel replicas compute bit-identical clip coefficients and stay in lock step. */
#define LNX_SUMSQ_WORKSPACE 1024
int lnx_sumsq(const float* g, int64_t n, float* sumsq, float* workspace, lnx_stream_t s);
/* norm_out[0] = sqrt(sumsq * gscale^2); coef_out[0] = clip > 0 ? min(1, clip/(norm+1e-6)) : 1 */
int lnx_clip_coef(const float* sumsq, float gscale, float clip, float* norm_out, float* coef_out, lnx_stream_t s);
/* AdamW (decoupled decay) on a flat buffer; g is multiplied by gscale*coef[0] first.
 * bias_corr{1,2} = 1 - beta^t from the caller, unless step_dev (device float, the
 * 1-based step count) is given; lr_dev (device float) overrides lr.  The device
 * variants keep a captured CUDA graph valid across steps. */
int lnx_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
              float weight_decay, float bias_corr1, float bias_corr2, float gscale, const float* coef,
              const float* lr_dev, const float* step_dev, lnx_stream_t s);

/* ---- mFormerV0 (RelativeAttention variant) inference --------------------- */
/* Gather for dense 3x3 convs (pad 1): out[b*Ho*Wo + ho*Wo + wo, (kh*3+kw)*C + c] = x[b, ho*stride-1+kh, wo*stride-1+kw, c]
 * (zero outside).  x is NHWC in `dtype` (Kpad = 9C), or the fp32 NCHW input image when x_is_nchw_f32 (any C, columns
 * zero padded to Kpad).  Input side of nn.Conv2d(k=3) of the stem (R/models/mFormerV0.py:166-190) and of
 * OverlapPatchEmbed.proj (R/models/blocks/relative_mhsa.py:57-66); the contraction runs on lnx_gemm. */
int lnx_im2col3x3(const void* x, int x_is_nchw_f32, void* out, int B, int H, int W, int C, int stride, int Ho, int Wo, int Kpad, int dtype,
                  lnx_stream_t s);
/* Dense 3x3 convolution, stride 1, pad 1, as an implicit GEMM on the tensor cores (no im2col buffer): x [B,H,W,C] bf16
 * (C <= 64, C % 8 == 0), w9 [N, 9*64] bf16 (per output channel: 9 taps x 64 input channels, zero padded), bias float[N]
 * (nullable; the folded BatchNorm shift), optional ReLU, y [B,H,W,N] bf16 (N % 16 == 0; the 9 weight tiles stay resident in
 * shared memory, which bounds N at 96).  The stem convs 2 and 3
 * of mFormerV0 (R/models/mFormerV0.py:175-190).  Returns LNX_ERR_UNSUPPORTED for other shapes / dtypes. */
int lnx_conv3x3_s1(const void* x, const void* w9, const float* bias, void* y, int B, int H, int W, int C, int N, int relu, int dtype,
                   lnx_stream_t s);
/* nn.MaxPool2d(3, 2, 1) on NHWC (R/models/mFormerV0.py:193). */
int lnx_maxpool3s2(const void* x, void* y, int B, int H, int W, int C, int dtype, lnx_stream_t s);
/* Depthwise 3x3 with explicit top/left padding (TF "same" static padding of Conv2dStaticSamePadding,
 * R/models/blocks/mb_conv.py:46-99), y = act(conv(x) * scale[c] + shift[c]) (folded BatchNorm; act 1 = swish), and
 * pool_sum[b, c] += sum over pixels of y (nullable; the squeeze-excite average pool, mb_conv.py:239). */
int lnx_dwconv3_fwd(const void* x, const float* w9c, const float* scale, const float* shift, void* y, float* pool_sum, int B, int H, int W,
                    int C, int stride, int pad_t, int pad_l, int Ho, int Wo, int act, int dtype, lnx_stream_t s);
/* y[b, p, c] = x[b, p, c] * sigmoid(gate[b, c])  (squeeze-excite gating, mb_conv.py:242). */
int lnx_se_scale(const void* x, const float* gate, void* y, int B, int HW, int C, int dtype, lnx_stream_t s);
/* out[b, n, h*hd + d] = softmax_j(scale * q k^T + bias[h, n, j]) v with q/k/v read from qkv [B, N, 3, heads, hd]
 * (RelativeAttention.forward, R/models/blocks/relative_mhsa.py:201-236); bias float [heads, N, N] or NULL. */
int lnx_attn_bias_fwd(const void* qkv, const float* bias, void* out, int B, int heads, int N, int hd, float scale, int dtype, lnx_stream_t s);


/* ---- validation metrics and inference post-processing ---- */
/* Rank of the ground-truth class in each head's logits row and the per-batch counters the reference's MetricsTracker
 * accumulates (R/utils/metrics/tracker.py:609-735: per-task top-1 / top-3; R/utils/metrics/chain_accuracy.py:51-364:
 * chain and partial-chain accuracy; R/utils/metrics/basic.py:79-133: accuracy(topk) is `rank < k`), with no host sync.
 * logits [B, ld] (`dtype`; head k occupies columns class_off[k] .. class_off[k+1], class_off a HOST array of K+1 ints,
 * K <= 16); targets int64 [K, B] (class indices; one-hot / soft targets are arg-maxed by the caller as the reference does).
 * rank[k,i] = #{c : z[c] > z[y] or (z[c] == z[y] and c < y)}  (0 <=> argmax == y; first index wins ties like torch.argmax).
 * ranks_out int32 [K, B] (nullable).  counters int64 [4K+4] (nullable), ADDED to: [0,K) top-1 correct per task; [K,2K) top-3
 * correct per task (top-1 when C_k < 3, tracker.py:722-724); [2K] samples with every task right; [2K+1] samples right on
 * tasks 0..highest task whose target != null_index; [2K+2] samples with any target != null_index; [2K+3] samples;
 * [2K+4,3K+4) top-1 correct among samples whose target of that task == null_index, [3K+4,4K+4) their number
 * (null / non-null accuracy split, tracker.py:786-912). */
int lnx_hier_metrics(const void* logits, int dtype, int64_t ld, int B, int K, const int* class_off, const int64_t* targets,
                     int null_index, int* ranks_out, int64_t* counters, lnx_stream_t s);
/* softmax + top-kk per head for the whole batch (R/inference/handler.py:186-214: softmax -> topk -> .item() per sample and
 * task).  idx_out int32 [K, B, kk], prob_out float [K, B, kk], ordered by (probability descending, class index ascending);
 * slots past min(kk, C_k) hold -1 / 0.  kk <= 64. */
int lnx_hier_topk(const void* logits, int dtype, int64_t ld, int B, int K, const int* class_off, int kk, int* idx_out, float* prob_out,
                  lnx_stream_t s);

/* Parent / child consistency of the per-rank top-1 predictions, in place on lnx_hier_topk's output (R/inference/postprocessing.py:14-171,
 * one sample at a time there).  Task 0 = lowest rank, K - 1 = highest.  parent int32 [sum C_k]: parent[class_off[k] + c] = class index,
 * in task k + 1, of the tree parent of class c of task k (-1: none); null_idx int [K] (host): null class of each task (-1: none).
 * Walking down from the highest rank, a rank whose parent rank is null, or whose top-1 class is not a child of the parent rank's
 * (consistent) prediction, becomes the single entry (null class, 1.0): idx[k, b, :] = {null, -1, ...}, prob = {1, 0, ...};
 * changed uint8 [K, B] (nullable) marks those rows. */
int lnx_hier_consistency(int* idx, float* prob, const int* parent, const int* class_off, const int* null_idx, unsigned char* changed,
                         int B, int K, int kk, lnx_stream_t s);

/* ---- batch augmentation feeding the model: the apply step of selective mixup ---- */
/* out[i, :] = lam * x[i, :] + (1 - lam) * x[perm[i], :]  (fp32 mul, mul, add: bit-equal to the reference expression
 * `lam * v + (1 - lam) * v[perm]`, R/aug/gpu/selective_mixup.py:150,177).  x, out float [B, row] (out != x); perm int64 [B];
 * lam a DEVICE scalar (the Beta sample never visits the host). */
int lnx_mix_pairs(const float* x, const int64_t* perm, const float* lam, float* out, int B, int64_t row, lnx_stream_t s);
/* Metadata side of selective mixup (R/aug/gpu/selective_mixup.py:320-328,371-392,394-560).  aux float [B, D] and mask
 * uint8 (torch.bool) [B, D] are first "all-or-nothing" enforced IN PLACE: a chunk [lo, hi) with any entry == 0 is zeroed
 * and its mask cleared.  Then per sample i and chunk: both the original and the partner (perm[i]) chunk non-zero -> the
 * original iff pick[i] < 0.5 else the partner; exactly one non-zero -> that one; both zero -> zeros / mask 0; written to
 * out_aux / out_mask (distinct buffers; entries outside every chunk are not written).  chunk_bounds: HOST array of
 * 2 * n_chunks ints (lo, hi), n_chunks <= 16; pick float [B] uniform numbers (device). */
int lnx_mix_meta_chunks(float* aux, uint8_t* mask, const int64_t* perm, const float* pick, const int* chunk_bounds, int n_chunks,
                        float* out_aux, uint8_t* out_mask, int B, int D, lnx_stream_t s);
/* Selective CutMix, image side (R/aug/gpu/selective_cutmix.py:204-236): out = x, except that inside the box
 * [h1, h2) x [w1, w2) of dims 2 / 3 every sample with group_ids[i] != -1 takes the pixels of x[perm[i]].  x, out float
 * [B, C, H, W] (out != x); perm, group_ids int64 [B]. */
int lnx_cutmix_paste(const float* x, const int64_t* perm, const int64_t* group_ids, float* out, int B, int C, int H, int W, int h1, int w1,
                     int h2, int w2, lnx_stream_t s);
/* Selective CutMix, target side (selective_cutmix.py:266-269): out[i] = group_ids[i] != -1 ?
 * coef_self * x[i] + coef_partner * x[perm[i]] : x[i]   (fp32 mul, mul, add; the caller passes (float)lam_adjusted and
 * (float)(1.0 - lam_adjusted), the two Python-float scalars of the reference expression). */
int lnx_mix_pairs_valid(const float* x, const int64_t* perm, const int64_t* group_ids, float coef_self, float coef_partner, float* out, int B,
                        int64_t row, lnx_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* LINNAEUS_B200_H */
