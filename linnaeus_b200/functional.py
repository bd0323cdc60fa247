"""Autograd bindings of the sm_100a kernels (one ``torch.autograd.Function`` per fused op).

Every forward/backward here is a sequence of C-ABI calls (``_lib.call``); PyTorch only
allocates the tensors.  Activations are float32 (1e-4 parity mode) or bfloat16;
parameters stay float32 (``weight``) with a compute-dtype shadow (``weight_c``) in
bf16 mode; parameter gradients are always float32.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_GELU_DG, ACT_MUL, ACT_NONE, ACT_RELU, call, dt, ptr

# set by tests to force the CUDA-core kernels instead of tcgen05 ones
FORCE_SIMT = False

_ACT = {None: ACT_NONE, "none": ACT_NONE, "gelu": ACT_GELU, "relu": ACT_RELU, "swish": _lib.ACT_SWISH}


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def cast_bf16(src: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """float32 -> bfloat16 with the library's cast kernel."""
    src = _c(src)
    if out is None:
        out = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    call("lnx_cast_f32_to_bf16", src.data_ptr(), out.data_ptr(), src.numel())
    return out


def compute_copy(p: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Parameter in the compute dtype (detached)."""
    p = p.detach()
    if dtype == torch.float32:
        return _c(p)
    return cast_bf16(p)


def gemm(a, b, M, N, K, *, a_trans=False, b_trans=False, lda=None, ldb=None, out=None, out_dtype=None, bias=None, act=ACT_NONE,
         aux_out=None, act_grad_in=None, residual=None, col_scale=None, row_scale=None, rows_per_group=0, colsum_out=None,
         accumulate=False):
    """Raw lnx_gemm call.  ``a``/``b`` share one dtype (f32 or bf16)."""
    if lda is None:
        lda = M if a_trans else K
    if ldb is None:
        ldb = N if b_trans else K
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype or a.dtype, device=a.device)
    call("lnx_gemm", dt(a), a.data_ptr(), lda, int(a_trans), b.data_ptr(), ldb, int(b_trans), out.data_ptr(), dt(out), M, N, K,
         ptr(bias), act, ptr(aux_out), ptr(act_grad_in), ptr(residual), ptr(col_scale), ptr(row_scale), int(rows_per_group), ptr(colsum_out), int(accumulate), int(FORCE_SIMT))
    return out


# Parameter gradients are accumulated straight into an existing float32 ``param.grad`` (e.g. the views of
# FlatAdamW's flat buffer) instead of being returned to autograd: no per-tensor zero-fill + add launches.
DIRECT_GRAD = True


def _sink(param):
    if (DIRECT_GRAD and isinstance(param, torch.nn.Parameter) and param.grad is not None and param.grad.dtype == torch.float32
            and param.grad.is_contiguous()):
        return param.grad
    return None


def _grad_done(param) -> None:
    """Tell a data-parallel wrapper that this parameter's gradient is final (replaces the autograd hook)."""
    h = getattr(param, "_lnx_grad_ready", None)
    if h is not None:
        h(param)


def colsum(x2d: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    M, N = x2d.shape
    if out is None:
        out = torch.zeros(N, dtype=torch.float32, device=x2d.device)
    call("lnx_colsum", x2d.data_ptr(), out.data_ptr(), M, N, dt(x2d))
    return out


def rowscale(x2d: torch.Tensor, s: torch.Tensor, rows_per_group: int) -> torch.Tensor:
    M, N = x2d.shape
    out = torch.empty_like(x2d)
    call("lnx_rowscale", x2d.data_ptr(), s.data_ptr(), out.data_ptr(), M, N, rows_per_group, dt(x2d))
    return out


def wgrad(dy2d: torch.Tensor, x2d: torch.Tensor, x_ld: int | None = None, out: torch.Tensor | None = None,
          db_out: torch.Tensor | None = None) -> torch.Tensor:
    """dW[N,K] (+)= dy[M,N]^T x[M,K] and, when ``db_out`` is given, db[N] += colsum(dy); float32, split over M with
    atomic accumulation (into ``out`` when given).  bf16 operands run the dedicated tcgen05 kernel (lnx_wgrad: the
    bias gradient is an extra MMA against a tile of ones); anything else the generic GEMM + a column-sum pass."""
    M, N = dy2d.shape
    K = x2d.shape[1]
    dw = out if out is not None else torch.zeros((N, K), dtype=torch.float32, device=dy2d.device)
    if (dy2d.dtype == torch.bfloat16 and x2d.dtype == torch.bfloat16 and not FORCE_SIMT and N % 8 == 0 and K % 4 == 0
            and (x_ld or K) % 8 == 0 and dw.data_ptr() % 16 == 0):
        call("lnx_wgrad", dy2d.data_ptr(), N, x2d.data_ptr(), x_ld or K, dw.data_ptr(), ptr(db_out), M, N, K, _lib.BF16)
        return dw
    gemm(dy2d, x2d, N, K, M, a_trans=True, b_trans=True, lda=N, ldb=x_ld or K, out=dw, accumulate=True)
    if db_out is not None:
        colsum(dy2d, out=db_out)
    return dw


# --------------------------------------------------------------------------- Linear
def _linear_param_grads(weight, bias, dpre, dpre_w, x2, lda, N, K, want_dw, want_db):
    """dW = dpre^T x and db = colsum(dpre) of a Linear layer, accumulated straight into the parameters' gradient sinks when they
    exist (then (None, None) is returned); the bias gradient rides on the weight-gradient GEMM when the dtypes allow."""
    dw = db = None
    sb = _sink(bias) if want_db else None
    db_buf = None
    if want_db:
        db_buf = sb if sb is not None else torch.zeros(N, dtype=torch.float32, device=dpre.device)
    db_done = False
    if want_dw:
        sw = _sink(weight)
        fuse_db = want_db and dpre_w is dpre  # the bias gradient rides on the weight-gradient GEMM
        if sw is not None:
            wgrad(dpre_w, x2, x_ld=lda, out=sw.view(N, K), db_out=db_buf if fuse_db else None)
            _grad_done(weight)
        else:
            dw = wgrad(dpre_w, x2, x_ld=lda, db_out=db_buf if fuse_db else None)
        db_done = fuse_db
    if want_db:
        if not db_done:
            colsum(dpre, out=db_buf)
        if sb is not None:
            _grad_done(bias)
        else:
            db = db_buf
    return dw, db


class _Linear(torch.autograd.Function):
    """y = act(x W^T + b) (+ residual).  nn.Linear (+GELU/ReLU) of the reference."""

    @staticmethod
    def forward(ctx, x, weight, bias, weight_c, act, residual, x_ld, out_dtype, row_scale, rpg, grad_on=True):
        K = weight.shape[1]
        N = weight.shape[0]
        lead = x.shape[:-1]
        if x_ld is None:
            x2 = _c(x).view(-1, K)
            lda = K
        else:  # a column slice of a wider row-major matrix (metadata components)
            x2 = x
            lda = x_ld
        M = x2.shape[0]
        wc = weight_c if weight_c is not None else compute_copy(weight, x2.dtype)
        if out_dtype is None:
            out_dtype = wc.dtype if x2.dtype != wc.dtype else x2.dtype
        if x2.dtype != wc.dtype:  # f32 metadata into a bf16 model: run the f32 kernel, emit bf16
            wc_use = _c(weight.detach())
        else:
            wc_use = wc
        # ctx.needs_input_grad ignores torch.no_grad(): the wrapper passes the grad mode so inference saves nothing
        need_grad = grad_on and any(ctx.needs_input_grad)
        aux = torch.empty((M, N), dtype=out_dtype, device=x2.device) if (act != ACT_NONE and need_grad) else None
        res2 = _c(residual).view(M, N) if residual is not None else None
        y = gemm(x2, wc_use, M, N, K, lda=lda, out_dtype=out_dtype, bias=bias, act=act, aux_out=aux, residual=res2,
                 row_scale=row_scale, rows_per_group=rpg)
        ctx.save_for_backward(x2, wc_use, aux, row_scale)
        ctx.rpg = rpg
        ctx.params = (weight, bias)
        ctx.meta = (act, lead, K, N, lda, bias is not None, residual is not None, x.shape)
        return y.view(*lead, N) if x_ld is None else y

    @staticmethod
    def backward(ctx, dy):
        x2, wc, aux, row_scale = ctx.saved_tensors
        act, lead, K, N, lda, has_bias, has_res, xshape = ctx.meta
        M = x2.shape[0]
        dy2 = _c(dy).view(M, N)
        d_res = dy if has_res else None
        if row_scale is not None:  # DropPath: the branch gradient carries the per-sample mask
            dy2 = rowscale(dy2, row_scale, ctx.rpg)
        if act == ACT_NONE and dy2.dtype != x2.dtype:  # e.g. float32 logits out of a bf16 trunk
            dy2 = cast_bf16(dy2) if x2.dtype == torch.bfloat16 else dy2.float()
        if act != ACT_NONE:
            dpre = torch.empty_like(dy2)
            call("lnx_act_bwd", dy2.data_ptr(), aux.data_ptr(), dpre.data_ptr(), dy2.numel(), act, dt(dy2))
        else:
            dpre = dy2
        dx = dw = db = None
        if dpre.dtype != x2.dtype:  # mixed (f32 x, bf16 activations): gradients in f32
            dpre_w = dpre.float()
        else:
            dpre_w = dpre
        if ctx.needs_input_grad[0]:
            dx = gemm(dpre_w, wc, M, K, N, b_trans=True, ldb=K).view(xshape)
        weight, bias = ctx.params
        dw, db = _linear_param_grads(weight, bias if has_bias else None, dpre, dpre_w, x2, lda, N, K, ctx.needs_input_grad[1],
                                     has_bias and ctx.needs_input_grad[2])
        return dx, dw, db, None, None, d_res, None, None, None, None, None


def linear(x, weight, bias=None, weight_c=None, act=None, residual=None, x_ld=None, out_dtype=None, row_scale=None, rows_per_group=0):
    return _Linear.apply(x, weight, bias, weight_c, _ACT[act], residual, x_ld, out_dtype, row_scale, rows_per_group, torch.is_grad_enabled())


# --------------------------------------------------------------------------- fused ConvNeXt pointwise pair
FUSED_MLP = True  # tests flip this to compare against the two-GEMM path
_FUSED_C = (96, 192)


def fused_mlp_ok(x2: torch.Tensor, K: int, Hd: int, N: int, act) -> bool:
    """The single-kernel pointwise pair (lnx_mlp_fused_*) covers bf16, C in {96, 192}, hidden = 4 C, GELU."""
    return (FUSED_MLP and not FORCE_SIMT and x2.dtype == torch.bfloat16 and act == ACT_GELU and K == N and K in _FUSED_C
            and Hd == 4 * K)


def mlp_fused_fwd(x2, w1c, b1, w2c, b2, gamma=None, row_scale=None, rows_per_group=0, residual=None, out=None):
    """y = residual + row_scale * gamma * (gelu(x W1^T + b1) W2^T + b2): one tcgen05 kernel, hidden tile kept in TMEM."""
    M, C = x2.shape
    Hd = w1c.shape[0]
    if out is None:
        out = torch.empty((M, C), dtype=torch.bfloat16, device=x2.device)
    call("lnx_mlp_fused_fwd", x2.data_ptr(), w1c.data_ptr(), ptr(b1), w2c.data_ptr(), ptr(b2), ptr(gamma), ptr(row_scale),
         int(rows_per_group), ptr(residual), out.data_ptr(), M, C, Hd)
    return out


def mlp_fused_bwd(x2, dy2, w1c, b1, w2e, store_hidden: bool = True):
    """-> (h, dpre, dx): recomputes the pre-activation from x, one kernel (see lnx_mlp_fused_bwd).  ``store_hidden=False``: only
    dx is produced (h = dpre = None); the weight gradients then come from :func:`mlp_fused_wgrad`."""
    M, C = x2.shape
    Hd = w1c.shape[0]
    h = torch.empty((M, Hd), dtype=torch.bfloat16, device=x2.device) if store_hidden else None
    dpre = torch.empty((M, Hd), dtype=torch.bfloat16, device=x2.device) if store_hidden else None
    dx = torch.empty((M, C), dtype=torch.bfloat16, device=x2.device)
    call("lnx_mlp_fused_bwd", x2.data_ptr(), dy2.data_ptr(), w1c.data_ptr(), ptr(b1), w2e.data_ptr(), ptr(h), ptr(dpre), dx.data_ptr(), M, C, Hd)
    return h, dpre, dx


def mlp_fused_wgrad(x2, dy2, w1c, b1, w2e, dw1, db1, dw2_raw, db2_raw):
    """dw1 += dpre^T x, db1 += colsum(dpre), dw2_raw += dy^T h (accumulated on chip, nothing 4C-wide in HBM); db2_raw += colsum(dy)."""
    M, C = x2.shape
    call("lnx_mlp_fused_wgrad", x2.data_ptr(), dy2.data_ptr(), w1c.data_ptr(), ptr(b1), w2e.data_ptr(), dw1.data_ptr(), ptr(db1),
         dw2_raw.data_ptr(), M, C, w1c.shape[0])
    if db2_raw is not None:
        colsum(dy2, out=db2_raw)


# --------------------------------------------------------------------------- two-layer MLP
class _Mlp2(torch.autograd.Function):
    """y = [residual +] [col_scale *] (act(x W1^T + b1) W2^T + b2).

    ConvNeXt pointwise pair with layer scale (convnext.py:79-86), transformer Mlp
    (mlp.py:61-65) and cl_1_fc's Mlp.  The backward fuses act' into the epilogue of
    the dH GEMM and derives d(col_scale) from the un-scaled weight gradient:
    dgamma = rowsum(dW2_raw * W2) + b2 * db2_raw  (no extra pass over activations)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w1c, w2c, act, residual, col_scale, row_scale, rpg, grad_on=True):
        K = w1.shape[1]
        Hd = w1.shape[0]
        N = w2.shape[0]
        lead = x.shape[:-1]
        x2 = _c(x).view(-1, K)
        M = x2.shape[0]
        w1c = w1c if w1c is not None else compute_copy(w1, x2.dtype)
        w2c = w2c if w2c is not None else compute_copy(w2, x2.dtype)
        need_grad = grad_on and any(ctx.needs_input_grad)  # grad_on: see _Linear.forward
        ctx.fused = False
        if fused_mlp_ok(x2, K, Hd, N, act) and (not need_grad or K == 96):
            # ConvNeXt pointwise pair in ONE kernel (hidden tile kept in TMEM); nothing but x is saved: the backward
            # recomputes the pre-activation on the tensor cores (lnx_mlp_fused_bwd)
            res2 = _c(residual).view(M, N) if residual is not None else None
            y = mlp_fused_fwd(x2, w1c, b1, w2c, b2, gamma=col_scale, row_scale=row_scale, rows_per_group=rpg, residual=res2)
            if need_grad:
                ctx.fused = True
                ctx.save_for_backward(x2, w1c, w2c, None, None, col_scale, w2, b2, row_scale)
                ctx.rpg = rpg
                ctx.params = (w1, b1, w2, b2, col_scale)
                ctx.meta = (act, lead, K, Hd, N, residual is not None, x.shape)
            return y.view(*lead, N)
        # GELU: the forward epilogue stores gelu'(pre) (one tanh serves both), so the backward epilogue is a multiply
        pre = torch.empty((M, Hd), dtype=x2.dtype, device=x2.device) if need_grad else None
        save_dg = need_grad and act == ACT_GELU
        h = gemm(x2, w1c, M, Hd, K, bias=b1, act=ACT_GELU_DG if save_dg else act, aux_out=pre)
        ctx.save_dg = save_dg
        res2 = _c(residual).view(M, N) if residual is not None else None
        y = gemm(h, w2c, M, N, Hd, bias=b2, residual=res2, col_scale=col_scale, row_scale=row_scale, rows_per_group=rpg)
        ctx.save_for_backward(x2, w1c, w2c, pre, h, col_scale, w2, b2, row_scale)
        ctx.rpg = rpg
        ctx.params = (w1, b1, w2, b2, col_scale)
        ctx.meta = (act, lead, K, Hd, N, residual is not None, x.shape)
        return y.view(*lead, N)

    @staticmethod
    def backward(ctx, dy):
        x2, w1c, w2c, pre, h, col_scale, w2, b2, row_scale = ctx.saved_tensors
        act, lead, K, Hd, N, has_res, xshape = ctx.meta
        M = x2.shape[0]
        dy2 = _c(dy).view(M, N)
        d_res = dy if has_res else None
        if row_scale is not None:  # DropPath: the branch gradient carries the per-sample mask
            dy2 = rowscale(dy2, row_scale, ctx.rpg)
        if col_scale is not None:
            # fold the layer scale into the weight seen by the data-gradient GEMM
            if dy2.dtype == torch.bfloat16 and w2.is_contiguous():
                w2_eff = torch.empty(w2.shape, dtype=torch.bfloat16, device=w2.device)
                call("lnx_rowscale_cast_bf16", w2.data_ptr(), col_scale.data_ptr(), w2_eff.data_ptr(), N, Hd)
            else:
                w2_eff = compute_copy(w2.detach() * col_scale.detach()[:, None], dy2.dtype)
        else:
            w2_eff = w2c
        # dPre = (dy W2_eff) * act'(pre)   [M, Hd]
        p_w1, p_b1, p_w2, p_b2, p_cs = ctx.params
        s_w1, s_b1, s_w2, s_b2 = _sink(p_w1), _sink(p_b1), _sink(p_w2), _sink(p_b2)
        db1 = s_b1 if s_b1 is not None else torch.zeros(Hd, dtype=torch.float32, device=dy2.device)
        dx_fused = None
        if ctx.fused:
            h, dpre, dx_fused = mlp_fused_bwd(x2, dy2, w1c, p_b1.detach() if p_b1 is not None else None, w2_eff)
        else:
            dpre = gemm(dy2, w2_eff, M, Hd, N, b_trans=True, ldb=Hd, act=ACT_MUL if ctx.save_dg else act, act_grad_in=pre)
        d_cs = dw2 = db2 = None
        if col_scale is not None:
            raw = torch.zeros(N * Hd + N, dtype=torch.float32, device=dy2.device)  # one fill: dW2_raw | db2_raw
            dw2_raw, db2_raw = raw[:N * Hd].view(N, Hd), raw[N * Hd:]
            wgrad(dy2, h, out=dw2_raw, db_out=db2_raw)
            cs = col_scale.detach()
            s_cs = _sink(p_cs)
            if s_w2 is not None and s_b2 is not None and s_cs is not None and w2.is_contiguous():
                # one launch: dW2 += cs dW2_raw, db2 += cs db2_raw, dgamma += rowsum(dW2_raw * W2) + b2 db2_raw
                call("lnx_layerscale_bwd", dw2_raw.data_ptr(), db2_raw.data_ptr(), w2.data_ptr(), b2.data_ptr(), cs.data_ptr(),
                     s_w2.data_ptr(), s_b2.data_ptr(), s_cs.data_ptr(), N, Hd)
                _grad_done(p_cs)
            else:
                d_cs = (dw2_raw * w2.detach()).sum(1) + b2.detach() * db2_raw
                if s_w2 is not None:
                    s_w2.view(N, Hd).addcmul_(dw2_raw, cs[:, None])
                    s_b2.addcmul_(db2_raw, cs)
                else:
                    dw2, db2 = dw2_raw * cs[:, None], db2_raw * cs
                if s_cs is not None:
                    s_cs.add_(d_cs)
                    _grad_done(p_cs)
                    d_cs = None
        elif s_w2 is not None:
            wgrad(dy2, h, out=s_w2.view(N, Hd), db_out=s_b2)
        else:
            db2 = torch.zeros(N, dtype=torch.float32, device=dy2.device)
            dw2 = wgrad(dy2, h, db_out=db2)
        if dx_fused is not None:
            dx = dx_fused.view(xshape)
        else:
            dx = gemm(dpre, w1c, M, K, Hd, b_trans=True, ldb=K).view(xshape) if ctx.needs_input_grad[0] else None
        dw1 = None
        if s_w1 is not None:
            wgrad(dpre, x2, out=s_w1.view(Hd, K), db_out=db1)  # db1 = colsum(dpre) comes out of the same GEMM
        else:
            dw1 = wgrad(dpre, x2, db_out=db1)
        if s_b1 is not None:
            db1 = None
        for prm, snk in ((p_w1, s_w1), (p_b1, s_b1), (p_w2, s_w2), (p_b2, s_b2)):
            if snk is not None:
                _grad_done(prm)
        return dx, dw1, db1, dw2, db2, None, None, None, d_res, d_cs, None, None, None


def mlp2(x, w1, b1, w2, b2, w1c=None, w2c=None, act="gelu", residual=None, col_scale=None, row_scale=None, rows_per_group=0):
    return _Mlp2.apply(x, w1, b1, w2, b2, w1c, w2c, _ACT[act], residual, col_scale, row_scale, rows_per_group, torch.is_grad_enabled())


# --------------------------------------------------------------------------- LayerNorm
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps, residual, fork):
        C = x.shape[-1]
        x2 = _c(x).view(-1, C)
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        res2 = _c(residual).view(rows, C) if residual is not None else None
        call("lnx_layernorm_fwd", x2.data_ptr(), w.data_ptr(), b.data_ptr(), ptr(res2), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
             rows, C, float(eps), dt(x2))
        ctx.save_for_backward(x2, w, mean, rstd)
        ctx.params = (w, b)
        ctx.has_res = residual is not None
        ctx.xshape = x.shape
        if fork:  # second output = the input itself (skip connection): its gradient is added inside the backward kernel
            return y.view(x.shape), x.view_as(x)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy, dskip=None):
        x2, w, mean, rstd = ctx.saved_tensors
        rows, C = x2.shape
        dy2 = _c(dy).view(rows, C)
        dres = _c(dskip).view(rows, C) if dskip is not None else None
        dx = torch.empty_like(x2)
        p_w, p_b = ctx.params
        s_w, s_b = _sink(p_w), _sink(p_b)
        direct = s_w is not None and s_b is not None
        dw = s_w if direct else torch.zeros(C, dtype=torch.float32, device=x2.device)
        db = s_b if direct else torch.zeros(C, dtype=torch.float32, device=x2.device)
        call("lnx_layernorm_bwd", dy2.data_ptr(), x2.data_ptr(), w.data_ptr(), mean.data_ptr(), rstd.data_ptr(), ptr(dres), dx.data_ptr(),
             dw.data_ptr(), db.data_ptr(), rows, C, dt(x2))
        if direct:
            _grad_done(p_w)
            _grad_done(p_b)
            dw = db = None
        return dx.view(ctx.xshape), dw, db, None, (dy if ctx.has_res else None), None


def layernorm(x, w, b, eps=1e-5, residual=None):
    return _LayerNorm.apply(x, w, b, eps, residual, False)


def layernorm_fork(x, w, b, eps=1e-5):
    """-> (LN(x), x): use the second output as the skip connection of a pre-norm block (see _LayerNorm.forward)."""
    return _LayerNorm.apply(x, w, b, eps, None, True)


# --------------------------------------------------------------------------- depthwise 7x7
W_TAP_MAJOR, W_NATIVE, W_NATIVE_FLIPPED = 0, 1, 2  # LNX_DW_W_*


class _DwConv7(torch.autograd.Function):
    """Depthwise 7x7.  With ``fork`` the input is also returned as a second output (the skip connection of the
    ConvNeXt block): its gradient is then added inside the data-gradient kernel instead of by a separate
    autograd add over the whole activation.  The kernels read the Conv2d weight [C,1,7,7] in place (the data gradient
    reads it with the taps reversed) and the weight-gradient kernel accumulates in that layout straight into the
    parameter's gradient buffer: no transposed / flipped copies, no zero-fill + add launches."""

    @staticmethod
    def forward(ctx, x, weight, bias, fork):
        B, H, W, C = x.shape
        x = _c(x)
        w = _c(weight.detach())
        y = torch.empty_like(x)
        call("lnx_dwconv7_fwd", x.data_ptr(), w.data_ptr(), W_NATIVE, ptr(bias), None, y.data_ptr(), B, H, W, C, dt(x))
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.params = (weight, bias)
        if fork:
            return y, x.view_as(x)
        return y

    @staticmethod
    def backward(ctx, dy, dskip=None):
        x, w = ctx.saved_tensors
        B, H, W, C = x.shape
        dy = _c(dy)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            res = _c(dskip) if dskip is not None else None
            call("lnx_dwconv7_fwd", dy.data_ptr(), w.data_ptr(), W_NATIVE_FLIPPED, None, ptr(res), dx.data_ptr(), B, H, W, C, dt(x))
        p_w, p_b = ctx.params
        s_w = _sink(p_w)
        s_b = _sink(p_b) if p_b is not None else None
        direct = s_w is not None and (p_b is None or s_b is not None)
        dw = s_w if direct else torch.zeros((C, 1, 7, 7), dtype=torch.float32, device=x.device)
        db = s_b if direct else torch.zeros(C, dtype=torch.float32, device=x.device)
        call("lnx_dwconv7_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), W_NATIVE, ptr(db), B, H, W, C, dt(x))
        if direct:
            _grad_done(p_w)
            if p_b is not None:
                _grad_done(p_b)
            return dx, None, None, None
        return dx, dw, (db if ctx.has_bias else None), None


def dwconv7(x_nhwc, weight, bias):
    return _DwConv7.apply(x_nhwc, weight, bias, False)


def dwconv7_fork(x_nhwc, weight, bias):
    """-> (conv(x), x): use the second output as the block's skip connection (see _DwConv7)."""
    return _DwConv7.apply(x_nhwc, weight, bias, True)


# --------------------------------------------------------------------------- layout
def patchify(x_nchw: torch.Tensor, p: int, kpad: int, out_dtype: torch.dtype) -> torch.Tensor:
    """Stem im2col (no gradient: the image is a leaf input)."""
    B, Cin, H, W = x_nchw.shape
    x = _c(x_nchw.detach().float())
    out = torch.empty((B * (H // p) * (W // p), kpad), dtype=out_dtype, device=x.device)
    call("lnx_patchify_nchw", x.data_ptr(), out.data_ptr(), B, Cin, H, W, p, kpad, dt(out))
    return out


class _SpaceToDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        B, H, W, C = x.shape
        x = _c(x)
        out = torch.empty((B * (H // 2) * (W // 2), 4 * C), dtype=x.dtype, device=x.device)
        call("lnx_space_to_depth", x.data_ptr(), out.data_ptr(), B, H, W, C, 0, dt(x))
        ctx.shape = (B, H, W, C)
        return out

    @staticmethod
    def backward(ctx, dy):
        B, H, W, C = ctx.shape
        dy = _c(dy)
        dx = torch.empty((B, H, W, C), dtype=dy.dtype, device=dy.device)
        call("lnx_space_to_depth", dy.data_ptr(), dx.data_ptr(), B, H, W, C, 1, dt(dy))
        return dx


def space_to_depth(x_nhwc):
    return _SpaceToDepth.apply(x_nhwc)


class _TokensAssemble(torch.autograd.Function):
    """cat([cls.expand(B), extras, patches], dim=1)."""

    @staticmethod
    def forward(ctx, cls_param, extras, patches):
        B, n_patch, D = patches.shape
        n_meta = 0 if extras is None else extras.shape[1]
        patches = _c(patches)
        cls_c = compute_copy(cls_param.reshape(D), patches.dtype)
        ex = _c(extras) if extras is not None else None
        tokens = torch.empty((B, 1 + n_meta + n_patch, D), dtype=patches.dtype, device=patches.device)
        call("lnx_tokens_assemble", cls_c.data_ptr(), 0, ptr(ex), patches.data_ptr(), tokens.data_ptr(), B, n_meta, n_patch, D, dt(patches))
        ctx.dims = (B, n_meta, n_patch, D, cls_param.shape)
        return tokens

    @staticmethod
    def backward(ctx, dtok):
        B, n_meta, n_patch, D, cls_shape = ctx.dims
        dtok = _c(dtok)
        d_cls_rows = torch.empty((B, D), dtype=dtok.dtype, device=dtok.device)
        d_extras = torch.empty((B, n_meta, D), dtype=dtok.dtype, device=dtok.device) if n_meta else None
        d_patches = torch.empty((B, n_patch, D), dtype=dtok.dtype, device=dtok.device)
        call("lnx_tokens_split", dtok.data_ptr(), d_cls_rows.data_ptr(), ptr(d_extras), d_patches.data_ptr(), B, n_meta, n_patch, D, dt(dtok))
        d_cls = colsum(d_cls_rows).view(cls_shape)
        return d_cls, d_extras, d_patches


def tokens_assemble(cls_param, extras, patches):
    return _TokensAssemble.apply(cls_param, extras, patches)


class _TokensSplit(torch.autograd.Function):
    """tokens -> (cls rows [B,D], patch tokens [B,n_patch,D]); the extra (meta) rows are dropped."""

    @staticmethod
    def forward(ctx, tokens, n_meta):
        B, N, D = tokens.shape
        n_patch = N - 1 - n_meta
        tokens = _c(tokens)
        cls_rows = torch.empty((B, D), dtype=tokens.dtype, device=tokens.device)
        patches = torch.empty((B, n_patch, D), dtype=tokens.dtype, device=tokens.device)
        call("lnx_tokens_split", tokens.data_ptr(), cls_rows.data_ptr(), None, patches.data_ptr(), B, n_meta, n_patch, D, dt(tokens))
        ctx.dims = (B, n_meta, n_patch, D)
        return cls_rows, patches

    @staticmethod
    def backward(ctx, d_cls, d_patches):
        B, n_meta, n_patch, D = ctx.dims
        ref = d_cls if d_cls is not None else d_patches
        if d_cls is None:
            d_cls = torch.zeros((B, D), dtype=ref.dtype, device=ref.device)
        if d_patches is None:
            d_patches = torch.zeros((B, n_patch, D), dtype=ref.dtype, device=ref.device)
        d_cls, d_patches = _c(d_cls), _c(d_patches)
        dtok = torch.empty((B, 1 + n_meta + n_patch, D), dtype=ref.dtype, device=ref.device)
        call("lnx_tokens_assemble", d_cls.data_ptr(), D, None, d_patches.data_ptr(), dtok.data_ptr(), B, n_meta, n_patch, D, dt(dtok))
        return dtok, None


def tokens_split(tokens, n_meta):
    return _TokensSplit.apply(tokens, n_meta)


# --------------------------------------------------------------------------- RoPE attention
class _RopeAttention(torch.autograd.Function):
    """qkv [B,N,3*D] -> softmax((q*cos*s)(k*cos)^T) v -> [B,N,D]   (rope_2d_mhsa.py:432-501)."""

    @staticmethod
    def forward(ctx, qkv, freqs, H, W, heads, n_extra):
        B, N, D3 = qkv.shape
        D = D3 // 3
        hd = D // heads
        half = hd // 2
        qkv = _c(qkv)
        dev = qkv.device
        fr = _c(freqs.detach().float())
        cos = torch.empty((H * W, heads, half), dtype=torch.float32, device=dev)
        sin = torch.empty_like(cos)
        call("lnx_rope_table", fr.data_ptr(), cos.data_ptr(), sin.data_ptr(), H, W, heads, half)
        qkv_h = torch.empty((3, B, heads, N, hd), dtype=qkv.dtype, device=dev)
        scale = float(hd) ** -0.5
        call("lnx_rope_qk_fwd", qkv.data_ptr(), cos.data_ptr(), qkv_h[0].data_ptr(), qkv_h[1].data_ptr(), qkv_h[2].data_ptr(),
             B, N, heads, hd, n_extra, scale, dt(qkv))
        out = torch.empty((B, N, D), dtype=qkv.dtype, device=dev)
        lse = torch.empty((B, heads, N), dtype=torch.float32, device=dev)
        call("lnx_attn_fwd", qkv_h[0].data_ptr(), qkv_h[1].data_ptr(), qkv_h[2].data_ptr(), out.data_ptr(), lse.data_ptr(),
             B, heads, N, hd, dt(qkv), int(FORCE_SIMT))
        ctx.save_for_backward(qkv, qkv_h, out, lse, cos, sin)
        ctx.freqs_param = freqs
        ctx.dims = (B, N, D, heads, hd, half, H, W, n_extra, scale, freqs.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, qkv_h, out, lse, cos, sin = ctx.saved_tensors
        B, N, D, heads, hd, half, H, W, n_extra, scale, fshape = ctx.dims
        dout = _c(dout)
        dev = qkv.device
        dqkv_h = torch.empty_like(qkv_h)
        delta = torch.empty(B * heads * N * (hd + 1) + 4, dtype=torch.float32, device=dev)
        call("lnx_attn_bwd", qkv_h[0].data_ptr(), qkv_h[1].data_ptr(), qkv_h[2].data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
             dqkv_h[0].data_ptr(), dqkv_h[1].data_ptr(), dqkv_h[2].data_ptr(), delta.data_ptr(), B, heads, N, hd, dt(qkv), int(FORCE_SIMT))
        dqkv = torch.empty_like(qkv)
        dtheta = torch.zeros((H * W, heads, half), dtype=torch.float32, device=dev)
        call("lnx_rope_qk_bwd", dqkv_h[0].data_ptr(), dqkv_h[1].data_ptr(), dqkv_h[2].data_ptr(), qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(),
             dqkv.data_ptr(), dtheta.data_ptr(), B, N, heads, hd, n_extra, scale, dt(qkv))
        s_f = _sink(ctx.freqs_param)
        dfreqs = s_f if s_f is not None else torch.zeros(fshape, dtype=torch.float32, device=dev)
        call("lnx_rope_freq_grad", dtheta.data_ptr(), dfreqs.data_ptr(), H, W, heads, half)  # dfreqs +=
        if s_f is not None:
            _grad_done(ctx.freqs_param)
            dfreqs = None
        return dqkv, dfreqs, None, None, None, None


def rope_attention(qkv, freqs, H, W, heads, n_extra):
    return _RopeAttention.apply(qkv, freqs, H, W, heads, n_extra)


FUSED_QKV_ROPE = True  # tests flip this to compare against the split path (Linear -> rope_qk_fwd -> attention)


def fused_qkv_rope_ok(x: torch.Tensor, dim: int, heads: int, n_tokens: int) -> bool:
    """The fused projection + attention path covers bf16, head_dim 64, <= 240 tokens (the tcgen05 attention kernels)."""
    return (FUSED_QKV_ROPE and not FORCE_SIMT and x.dtype == torch.bfloat16 and dim == heads * 64 and n_tokens <= 240
            and x.shape[-1] % 8 == 0)


class _QkvRopeAttention(torch.autograd.Function):
    """x [B,N,K] -> softmax((q cos s)(k cos)^T) v [B,N,D] with q/k/v = x W^T + b   (rope_2d_mhsa.py:432-501, self.qkv included).

    The cos factors and the softmax scale are applied in the epilogue of the projection GEMM (lnx_qkv_rope_gemm) and the attention
    kernels read q / k / v straight from its [B,N,3,heads,64] output through 4-D tensor maps: no separate scaling pass, no head-major
    copies.  Backward: attention -> dq/dk/dv (head-major) -> lnx_rope_qk_bwd_scaled (factors, d theta from the scaled q / k,
    token-major dqkv) -> the Linear gradients."""

    @staticmethod
    def forward(ctx, x, weight, bias, weight_c, freqs, H, W, heads, n_extra, grad_on=True):
        B, N, K = x.shape
        D = weight.shape[0] // 3
        hd = D // heads
        half = hd // 2
        x2 = _c(x).view(-1, K)
        dev = x.device
        wc = weight_c if weight_c is not None else compute_copy(weight, x2.dtype)
        need_grad = grad_on and any(ctx.needs_input_grad)
        fr = _c(freqs.detach().float())
        scale = float(hd) ** -0.5
        qkv = torch.empty((B, N, 3 * D), dtype=x2.dtype, device=dev)
        call("lnx_qkv_rope_gemm", x2.data_ptr(), wc.data_ptr(), ptr(bias), fr.data_ptr(), qkv.data_ptr(), B * N, D, K, N, n_extra, W, scale)
        out = torch.empty((B, N, D), dtype=x2.dtype, device=dev)
        lse = torch.empty((B, heads, N), dtype=torch.float32, device=dev)
        call("lnx_attn_qkv_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, heads, N, hd, dt(qkv))
        if need_grad:
            ctx.save_for_backward(x2, wc, qkv, out, lse, fr)
            ctx.params = (weight, bias, freqs)
            ctx.dims = (B, N, K, D, heads, hd, half, H, W, n_extra, scale, freqs.shape, x.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        x2, wc, qkv, out, lse, fr = ctx.saved_tensors
        B, N, K, D, heads, hd, half, H, W, n_extra, scale, fshape, xshape = ctx.dims
        weight, bias, freqs = ctx.params
        dev = qkv.device
        dout = _c(dout)
        cos = torch.empty((H * W, heads, half), dtype=torch.float32, device=dev)  # [position][pair] tables of the backward kernel
        sin = torch.empty_like(cos)
        call("lnx_rope_table", fr.data_ptr(), cos.data_ptr(), sin.data_ptr(), H, W, heads, half)
        dqkv_h = torch.empty((3, B, heads, N, hd), dtype=qkv.dtype, device=dev)
        call("lnx_attn_qkv_bwd", qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv_h[0].data_ptr(), dqkv_h[1].data_ptr(),
             dqkv_h[2].data_ptr(), B, heads, N, hd, dt(qkv))
        dqkv = torch.empty_like(qkv)
        dtheta = torch.zeros((H * W, heads, half), dtype=torch.float32, device=dev)
        call("lnx_rope_qk_bwd_scaled", dqkv_h[0].data_ptr(), dqkv_h[1].data_ptr(), dqkv_h[2].data_ptr(), qkv.data_ptr(), cos.data_ptr(),
             sin.data_ptr(), dqkv.data_ptr(), dtheta.data_ptr(), B, N, heads, hd, n_extra, scale, dt(qkv))
        dfreqs = None
        if ctx.needs_input_grad[4]:
            s_f = _sink(freqs)
            dfreqs = s_f if s_f is not None else torch.zeros(fshape, dtype=torch.float32, device=dev)
            call("lnx_rope_freq_grad", dtheta.data_ptr(), dfreqs.data_ptr(), H, W, heads, half)  # dfreqs +=
            if s_f is not None:
                _grad_done(freqs)
                dfreqs = None
        M = B * N
        dqkv2 = dqkv.view(M, 3 * D)
        dx = gemm(dqkv2, wc, M, K, 3 * D, b_trans=True, ldb=K).view(xshape) if ctx.needs_input_grad[0] else None
        dw, db = _linear_param_grads(weight, bias, dqkv2, dqkv2, x2, K, 3 * D, K, ctx.needs_input_grad[1],
                                     bias is not None and ctx.needs_input_grad[2])
        return dx, dw, db, None, dfreqs, None, None, None, None, None


def qkv_rope_attention(x, weight, bias, weight_c, freqs, H, W, heads, n_extra):
    return _QkvRopeAttention.apply(x, weight, bias, weight_c, freqs, H, W, heads, n_extra, torch.is_grad_enabled())


# --------------------------------------------------------------------------- aggregate
class _Aggregate2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, c, w, b):
        a, c = _c(a), _c(c)
        B, D = a.shape
        w2 = _c(w.detach().reshape(2).float())
        out = torch.empty_like(a)
        call("lnx_aggregate2_fwd", a.data_ptr(), c.data_ptr(), w2.data_ptr(), b.data_ptr(), out.data_ptr(), B, D, dt(a))
        ctx.save_for_backward(a, c, w2)
        ctx.wshape = w.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        a, c, w2 = ctx.saved_tensors
        B, D = a.shape
        dout = _c(dout)
        da, dc = torch.empty_like(a), torch.empty_like(c)
        dw = torch.zeros(2, dtype=torch.float32, device=a.device)
        db = torch.zeros(1, dtype=torch.float32, device=a.device)
        call("lnx_aggregate2_bwd", dout.data_ptr(), a.data_ptr(), c.data_ptr(), w2.data_ptr(), da.data_ptr(), dc.data_ptr(), dw.data_ptr(),
             db.data_ptr(), B, D, dt(a))
        return da, dc, dw.view(ctx.wshape), db


def aggregate2(a, c, w, b):
    return _Aggregate2.apply(a, c, w, b)


# --------------------------------------------------------------------------- loss
class _HierLoss(torch.autograd.Function):
    """Fused K-task masked loss on the concatenated logits [B, sum C_k] -> scalar."""

    @staticmethod
    def forward(ctx, logits, targets, class_off, kind, smoothing, soft_mats, task_w, keep, null_flag, phase1, stats, count_all=None):
        # phase1: zero the loss of null samples (PHASE1 masking / criteria built with ignore_index = 0);
        # count_all: divide by the batch size instead of #(loss != 0) (the PHASE1 training branch, hierarchical_loss.py:241-276)
        if count_all is None:
            count_all = phase1
        logits = _c(logits)
        B, Ctot = logits.shape
        K = len(class_off) - 1
        dev = logits.device
        offs = (ctypes.c_int * (K + 1))(*class_off)
        mats = None
        if soft_mats is not None:
            mats = (ctypes.c_void_p * K)(*[m.data_ptr() for m in soft_mats])
        per = torch.empty((K, B), dtype=torch.float32, device=dev)
        raw = torch.empty((K, B), dtype=torch.float32, device=dev)
        lse = torch.empty((K, B), dtype=torch.float32, device=dev)
        wgt = torch.empty((K, B), dtype=torch.float32, device=dev)
        call("lnx_loss_fwd", logits.data_ptr(), dt(logits), B, K, offs, targets.data_ptr(), ptr(null_flag), ptr(keep), kind, float(smoothing),
             mats, int(phase1), per.data_ptr(), raw.data_ptr(), lse.data_ptr(), wgt.data_ptr())
        red = torch.empty(1 + 3 * K, dtype=torch.float32, device=dev)  # total | scale[K] | task_sum[K] | nvalid[K]
        call("lnx_loss_reduce", per.data_ptr(), ptr(task_w), B, K, int(count_all), red.data_ptr(), red[1:].data_ptr(), red[1 + K:].data_ptr(),
             red[1 + 2 * K:].data_ptr())
        ctx.save_for_backward(logits, targets, wgt, lse, red)
        ctx.meta = (class_off, kind, smoothing, soft_mats, K, B)
        stats["per_sample"] = per
        stats["raw"] = raw
        stats["task_sum"] = red[1 + K: 1 + 2 * K]
        stats["nvalid"] = red[1 + 2 * K:]
        return red[0].clone()

    @staticmethod
    def backward(ctx, g):
        logits, targets, wgt, lse, red = ctx.saved_tensors
        class_off, kind, smoothing, soft_mats, K, B = ctx.meta
        offs = (ctypes.c_int * (K + 1))(*class_off)
        mats = None
        if soft_mats is not None:
            mats = (ctypes.c_void_p * K)(*[m.data_ptr() for m in soft_mats])
        g32 = _c(g.detach().float().reshape(1))
        dlogits = torch.empty_like(logits)
        call("lnx_loss_bwd", logits.data_ptr(), dt(logits), B, K, offs, targets.data_ptr(), wgt.data_ptr(), kind, float(smoothing), mats,
             lse.data_ptr(), red[1:].data_ptr(), g32.data_ptr(), dlogits.data_ptr())
        return (dlogits,) + (None,) * 11


def hier_loss(logits_cat, targets_kb, class_off, kind=0, smoothing=0.1, soft_mats=None, task_w=None, keep=None, null_flag=None,
              phase1=False, stats=None, count_all=None):
    if stats is None:
        stats = {}
    return _HierLoss.apply(logits_cat, targets_kb, tuple(class_off), kind, smoothing, soft_mats, task_w, keep, null_flag, phase1, stats,
                           count_all)


class _PerSampleLoss(torch.autograd.Function):
    """One task, per-sample loss vector [B] (criterion API of the reference, reduction='none')."""

    @staticmethod
    def forward(ctx, logits, target, kind, smoothing, soft_mats, ignore_null):
        logits = _c(logits)
        B, C = logits.shape
        dev = logits.device
        offs = (ctypes.c_int * 2)(0, C)
        mats = (ctypes.c_void_p * 1)(soft_mats[0].data_ptr()) if soft_mats is not None else None
        tg = _c(target.long()).view(1, B)
        per = torch.empty((1, B), dtype=torch.float32, device=dev)
        lse = torch.empty((1, B), dtype=torch.float32, device=dev)
        wgt = torch.empty((1, B), dtype=torch.float32, device=dev)
        call("lnx_loss_fwd", logits.data_ptr(), dt(logits), B, 1, offs, tg.data_ptr(), None, None, kind, float(smoothing), mats,
             int(ignore_null), per.data_ptr(), None, lse.data_ptr(), wgt.data_ptr())
        ctx.save_for_backward(logits, tg, lse, wgt)
        ctx.meta = (kind, smoothing, soft_mats, C)
        return per[0]

    @staticmethod
    def backward(ctx, dper):
        logits, tg, lse, wgt = ctx.saved_tensors
        kind, smoothing, soft_mats, C = ctx.meta
        B = logits.shape[0]
        offs = (ctypes.c_int * 2)(0, C)
        mats = (ctypes.c_void_p * 1)(soft_mats[0].data_ptr()) if soft_mats is not None else None
        w = _c((wgt[0] * dper.float()).view(1, B))  # upstream per-sample gradient folded into the sample weight
        one = torch.ones(1, dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        call("lnx_loss_bwd", logits.data_ptr(), dt(logits), B, 1, offs, tg.data_ptr(), w.data_ptr(), kind, float(smoothing), mats,
             lse.data_ptr(), one.data_ptr(), one.data_ptr(), dlogits.data_ptr())
        return dlogits, None, None, None, None, None


def per_sample_loss(logits, target, kind=0, smoothing=0.1, soft_mats=None, ignore_null=False):
    return _PerSampleLoss.apply(logits, target, kind, smoothing, soft_mats, ignore_null)
