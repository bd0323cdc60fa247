"""GPU: the tensor-pipe depthwise 7x7 kernels (csrc/lnx_dwconv_mma.cu: banded matrix products on mma.sync) through the C ABI, against
a plain PyTorch fp32 convolution and against the fp32x2-FMA kernels they replace (lnx_dwconv7_set_impl switches between the two).

Shapes cover the bench stages (56 x 56 x 96, 28 x 28 x 192), several column tiles (W = 96, 36), widths that are not a multiple of
4 or 8, heights that are not a multiple of the 16-row tile / 8-row band, and a single-pixel-scale image.  The tensor-pipe kernels
multiply bf16 weights (what the reference's autocast Conv2d does); the reference convolution below therefore uses the
bf16-rounded weights, so the remaining error is the bf16 rounding of the outputs (forward / data gradient) or fp32 summation
order (weight gradient).  R/models/blocks/convnext.py:56-58,76."""
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu

DEV = "cuda"
SHAPES = [(2, 56, 56, 96), (3, 28, 28, 192), (1, 20, 23, 64), (2, 14, 14, 256), (1, 96, 96, 64), (2, 40, 36, 32), (1, 7, 9, 32), (2, 17, 30, 64)]
W_TAP_MAJOR, W_NATIVE, W_NATIVE_FLIPPED = 0, 1, 2


@pytest.fixture(autouse=True)
def _seed_and_impl():
    from linnaeus_b200 import _lib

    torch.manual_seed(0)
    lib = _lib.load()
    lib.lnx_dwconv7_set_impl(1)
    yield
    lib.lnx_dwconv7_set_impl(-1)


def _rel_max(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _rel_l2(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _data(B, H, W, C):
    x = torch.randn(B, H, W, C, device=DEV).to(torch.bfloat16)
    g = torch.randn(B, H, W, C, device=DEV).to(torch.bfloat16)
    w = 0.2 * torch.randn(C, 1, 7, 7, device=DEV)
    b = 0.1 * torch.randn(C, device=DEV)
    return x, g, w, b


def _fwd(x, w, layout, bias, res):
    from linnaeus_b200 import _lib

    B, H, W, C = x.shape
    y = torch.empty_like(x)
    _lib.call("lnx_dwconv7_fwd", x.data_ptr(), w.data_ptr(), layout, _lib.ptr(bias), _lib.ptr(res), y.data_ptr(), B, H, W, C, _lib.BF16)
    return y


def _wgrad(x, g, layout, dw, db):
    from linnaeus_b200 import _lib

    B, H, W, C = x.shape
    _lib.call("lnx_dwconv7_wgrad", x.data_ptr(), g.data_ptr(), dw.data_ptr(), layout, _lib.ptr(db), B, H, W, C, _lib.BF16)


@pytest.mark.parametrize("B,H,W,C", SHAPES)
def test_forward_and_data_gradient(B, H, W, C):
    x, g, w, b = _data(B, H, W, C)
    wq = w.bfloat16().float()
    xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    yr = TF.conv2d(xr, wq, b, padding=3, groups=C)
    yr.backward(g.float().permute(0, 3, 1, 2))
    y = _fwd(x, w.reshape(C, 49).contiguous(), W_NATIVE, b, None)
    assert _rel_max(y, yr.permute(0, 2, 3, 1)) < 6e-3
    # tap-major weights give the same bits
    y2 = _fwd(x, w.reshape(C, 49).t().contiguous(), W_TAP_MAJOR, b, None)
    assert torch.equal(y, y2)
    # data gradient = the same kernel on dY with the taps reversed, plus the fused skip-connection gradient
    skip = torch.randn_like(x)
    dx = _fwd(g, w.reshape(C, 49).contiguous(), W_NATIVE_FLIPPED, None, skip)
    ref = xr.grad.permute(0, 2, 3, 1) + skip.float()
    assert _rel_max(dx, ref) < 6e-3


@pytest.mark.parametrize("B,H,W,C", SHAPES)
def test_weight_and_bias_gradient(B, H, W, C):
    x, g, w, b = _data(B, H, W, C)
    xr = x.float().permute(0, 3, 1, 2).contiguous()
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    TF.conv2d(xr, wr, br, padding=3, groups=C).backward(g.float().permute(0, 3, 1, 2))
    # accumulates (+=) into whatever the buffers hold, in either weight layout
    dw0 = torch.randn(C, 49, device=DEV)
    db0 = torch.randn(C, device=DEV)
    dw, db = dw0.clone(), db0.clone()
    _wgrad(x, g, W_NATIVE, dw, db)
    assert _rel_l2(dw - dw0, wr.grad.reshape(C, 49)) < 1e-3
    assert _rel_l2(db - db0, br.grad) < 1e-3
    dwt = torch.zeros(49, C, device=DEV)
    _wgrad(x, g, W_TAP_MAJOR, dwt, None)
    assert _rel_l2(dwt.t(), wr.grad.reshape(C, 49)) < 1e-3


@pytest.mark.parametrize("B,H,W,C", [(2, 56, 56, 96), (1, 20, 23, 64), (1, 96, 96, 64)])
def test_agrees_with_the_fma_kernels(B, H, W, C):
    """Same call, both implementations: outputs differ only by the bf16 rounding of the weights (forward / data gradient) and by
    summation order (weight gradient)."""
    from linnaeus_b200 import _lib

    lib = _lib.load()
    x, g, w, b = _data(B, H, W, C)
    wn = w.reshape(C, 49).contiguous()
    skip = torch.randn_like(x)
    out = {}
    for impl in (1, 0):
        lib.lnx_dwconv7_set_impl(impl)
        dw = torch.zeros(C, 49, device=DEV)
        db = torch.zeros(C, device=DEV)
        _wgrad(x, g, W_NATIVE, dw, db)
        out[impl] = (_fwd(x, wn, W_NATIVE, b, None), _fwd(g, wn, W_NATIVE_FLIPPED, None, skip), dw, db)
    assert lib.lnx_dwconv7_set_impl(-1) == 0
    assert _rel_max(out[1][0], out[0][0]) < 1.2e-2 and _rel_max(out[1][1], out[0][1]) < 1.2e-2
    assert _rel_l2(out[1][2], out[0][2]) < 1e-4 and _rel_l2(out[1][3], out[0][3]) < 1e-4


def test_non_finite_inputs_stay_in_their_channel_and_image():
    """A banded matrix product multiplies every pixel of a 16-column K window by a (mostly zero) band entry, and 0 x NaN = NaN: a
    non-finite input reaches every output whose K window holds it (up to 9 columns beyond the true 7 x 7 window, and through the
    window overhang of the last column block the row above).  It must cover the true receptive field and never leave its channel
    or image (each channel is its own matrix product)."""
    B, H, W, C = 2, 24, 28, 32
    x, g, w, b = _data(B, H, W, C)
    x[0, 10, 12, 5] = float("nan")
    y = _fwd(x, w.reshape(C, 49).contiguous(), W_NATIVE, b, None).float()
    bad = torch.isnan(y)
    assert bad[0, 7:14, 9:16, 5].all()
    assert not bad[1].any() and not bad[0, :, :, :5].any() and not bad[0, :, :, 6:].any()
    assert not bad[0, :6, :, 5].any() and not bad[0, 15:, :, 5].any()
