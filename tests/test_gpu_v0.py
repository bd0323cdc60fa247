"""mFormerV0 (inference path): state_dict surface on CPU, CUDA parity against the committed reference logits
(tests/golden/v0_*.npz) and the CPU oracle."""
import pytest
import torch

from oracle import mformer_v0_oracle as V
from tests.support.golden_v0 import CASES, load_case

DEV = "cuda"


def test_v0_state_dict_surface_matches_reference_names():
    """CPU: every key / shape of the reference state_dict (restated by the oracle, pinned in test_oracle_v0_vs_reference)."""
    import linnaeus_b200 as L

    cfg, nc = L.make_synthetic_config_v0("sm", 224)
    model = L.build_model(cfg, nc)
    sd = model.state_dict()
    shapes = V.param_shapes(V.arch_from_config(cfg, nc))
    assert sorted(sd.keys()) == sorted(shapes.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    model.load_state_dict(V.synth_state_dict(V.arch_from_config(cfg, nc), 0))
    with pytest.raises(RuntimeError):
        model.eval()(torch.zeros(1, 3, 224, 224))  # CPU tensors raise: there is no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_v0_cuda_matches_reference_golden(name, dtype, tol):
    import linnaeus_b200 as L

    cfg, nc, batch, wseed, dseed, z = load_case(name)
    a = V.arch_from_config(cfg, nc)
    P = V.synth_state_dict(a, wseed)
    x, m = V.synth_batch(a, batch, dseed)
    model = L.build_model(cfg, nc)
    model.load_state_dict(P)
    model = model.to(DEV).eval().set_compute_dtype(dtype)
    with torch.no_grad():
        out = model(x.to(DEV), m.to(DEV))
        out2 = model(x.to(DEV), m.to(DEV))  # second call: cached folded weights
    assert list(out.keys()) == [t for t, _ in a.tasks]
    for t, _ in a.tasks:
        ref = torch.from_numpy(z[f"logits.{t}"])
        got = out[t].float().cpu()
        err = float((got - ref).abs().max() / ref.abs().max())
        assert err <= tol, (t, err)
        torch.testing.assert_close(out[t], out2[t], rtol=1e-3, atol=1e-3)  # SE pool sums are atomic: not bitwise stable
        if dtype == torch.float32:
            assert torch.equal(got.argmax(1), ref.argmax(1)), t
    with pytest.raises(NotImplementedError):
        model.train()(x.to(DEV), m.to(DEV))


@pytest.mark.gpu
def test_v0_kernels_against_torch():
    """lnx_im2col3x3 + GEMM, maxpool, depthwise 3x3 (+affine, swish, pool sums), SE gate, biased attention vs torch ops."""
    import torch.nn.functional as TF

    from linnaeus_b200._lib import call

    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, W, C = 2, 14, 14, 32
    x = torch.randn(B, H, W, C, device=DEV)
    for stride in (1, 2):
        Ho = (H + 2 - 3) // stride + 1
        a = torch.empty(B * Ho * Ho, 9 * C, device=DEV)
        call("lnx_im2col3x3", x.data_ptr(), 0, a.data_ptr(), B, H, W, C, stride, Ho, Ho, 9 * C, 0)
        w = torch.randn(24, C, 3, 3, device=DEV)
        ref = TF.conv2d(x.permute(0, 3, 1, 2), w, None, stride, 1).permute(0, 2, 3, 1).reshape(-1, 24)
        got = a @ w.permute(0, 2, 3, 1).reshape(24, -1).t()
        assert float((got - ref).abs().max()) < 1e-3
    img = torch.randn(B, 3, 20, 20, device=DEV)
    a = torch.empty(B * 10 * 10, 32, device=DEV)
    call("lnx_im2col3x3", img.data_ptr(), 1, a.data_ptr(), B, 20, 20, 3, 2, 10, 10, 32, 0)
    w = torch.randn(8, 3, 3, 3, device=DEV)
    ref = TF.conv2d(img, w, None, 2, 1).permute(0, 2, 3, 1).reshape(-1, 8)
    got = a[:, :27] @ w.permute(0, 2, 3, 1).reshape(8, -1).t()
    assert float((got - ref).abs().max()) < 1e-3 and float(a[:, 27:].abs().max()) == 0.0
    y = torch.empty(B, 7, 7, C, device=DEV)
    call("lnx_maxpool3s2", x.data_ptr(), y.data_ptr(), B, H, W, C, 0)
    assert torch.equal(y, TF.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1))
    for stride, (lo, hi) in ((1, (1, 1)), (2, (0, 1))):
        wd = torch.randn(C, 1, 3, 3, device=DEV)
        sc, sh = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
        xp = TF.pad(x.permute(0, 3, 1, 2), (lo, hi, lo, hi))
        ref = TF.conv2d(xp, wd, None, stride, 0, 1, C) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
        ref = (ref * torch.sigmoid(ref)).permute(0, 2, 3, 1)
        Ho = ref.shape[1]
        out = torch.empty(B, Ho, Ho, C, device=DEV)
        pool = torch.zeros(B, C, device=DEV)
        call("lnx_dwconv3_fwd", x.data_ptr(), wd.reshape(C, 9).t().contiguous().data_ptr(), sc.data_ptr(), sh.data_ptr(), out.data_ptr(),
             pool.data_ptr(), B, H, W, C, stride, lo, lo, Ho, Ho, 1, 0)
        assert float((out - ref).abs().max()) < 1e-4
        assert float((pool - ref.sum((1, 2))).abs().max()) < 1e-3
    gate = torch.randn(B, C, device=DEV)
    z = torch.empty_like(x)
    call("lnx_se_scale", x.data_ptr(), gate.data_ptr(), z.data_ptr(), B, H * W, C, 0)
    assert float((z - x * torch.sigmoid(gate)[:, None, None, :]).abs().max()) < 1e-5
    # (96, 8, 53), (128, 3, 64), (32, 5, 1), (64, 2, 17): the short-sequence shared-memory kernel (fp32; bf16 when head_dim > 64)
    for hd, heads, N in ((48, 8, 53), (96, 8, 200), (64, 2, 17), (48, 8, 200), (32, 4, 132), (16, 2, 240), (96, 8, 53), (128, 3, 64), (32, 5, 1)):
        qkv = torch.randn(B, N, 3, heads, hd, device=DEV)
        bias = torch.randn(heads, N, N, device=DEV)
        o = torch.empty(B, N, heads * hd, device=DEV)
        call("lnx_attn_bias_fwd", qkv.data_ptr(), bias.data_ptr(), o.data_ptr(), B, heads, N, hd, hd ** -0.5, 0)
        q, k, v = qkv.permute(2, 0, 3, 1, 4)
        ref = (torch.softmax(q * hd ** -0.5 @ k.transpose(-2, -1) + bias, -1) @ v).transpose(1, 2).reshape(B, N, heads * hd)
        assert float((o - ref).abs().max()) < 2e-4
        # bf16: head_dim <= 64 runs the tcgen05 kernel (TMA zero-fills the head dim up to 64), others the CUDA-core kernel
        qb, ob = qkv.bfloat16(), torch.empty(B, N, heads * hd, device=DEV, dtype=torch.bfloat16)
        call("lnx_attn_bias_fwd", qb.data_ptr(), bias.data_ptr(), ob.data_ptr(), B, heads, N, hd, hd ** -0.5, 1)
        q, k, v = qb.float().permute(2, 0, 3, 1, 4)
        refb = (torch.softmax(q * hd ** -0.5 @ k.transpose(-2, -1) + bias, -1) @ v).transpose(1, 2).reshape(B, N, heads * hd)
        assert float((ob.float() - refb).abs().max() / refb.abs().max()) < 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,C,N,relu", [(2, 16, 32, 64, 64, 1), (3, 112, 112, 48, 64, 1), (1, 13, 21, 64, 96, 0), (2, 8, 16, 16, 32, 1),
                                            (1, 24, 40, 32, 48, 0)])
def test_conv3x3_implicit_gemm(B, H, W, C, N, relu):
    """lnx_conv3x3_s1 (TMA-shifted windows + tcgen05, zero-filled borders / channels) against F.conv2d of the same bf16 data."""
    import torch.nn.functional as TF

    from linnaeus_b200._lib import call

    torch.manual_seed(C + N)
    torch.backends.cudnn.allow_tf32 = False
    x = torch.randn(B, H, W, C, device=DEV).bfloat16()
    w = (torch.randn(N, C, 3, 3, device=DEV) / (3 * C ** 0.5)).bfloat16()
    bias = torch.randn(N, device=DEV)
    w9 = TF.pad(w.permute(0, 2, 3, 1), (0, 64 - C)).reshape(N, 9 * 64).contiguous()
    y = torch.empty(B, H, W, N, device=DEV, dtype=torch.bfloat16)
    call("lnx_conv3x3_s1", x.data_ptr(), w9.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W, C, N, relu, 1)
    ref = TF.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, 1, 1)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    assert float((y.float() - ref).abs().max() / ref.abs().max()) < 1e-2
