"""GPU: the tcgen05/TMEM/TMA GEMM (all operand-major combinations, fused epilogues,
split-K accumulation) against fp32 matmul of the same bf16 inputs."""
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _ref(a, b, a_trans, b_trans):
    A = a.float().t() if a_trans else a.float()
    Bm = b.float().t() if b_trans else b.float()
    return A @ Bm.t()


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(0)


SHAPES = [(128, 128, 64), (256, 96, 128), (300, 200, 192), (1000, 384, 96), (4096, 96, 384), (130, 1576, 768), (77, 64, 64),
          (512, 768, 3072), (640, 2304, 768), (128, 256, 48)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_gemm_kmajor(M, N, K):
    import linnaeus_b200.functional as F

    kp = (K + 7) // 8 * 8
    a = torch.randn(M, kp, device=DEV).to(torch.bfloat16)[:, :K]
    b = torch.randn(N, kp, device=DEV).to(torch.bfloat16)[:, :K]
    out = F.gemm(a, b, M, N, K, lda=kp, ldb=kp, out_dtype=torch.float32)
    assert rel_err(out, a.float() @ b.float().t()) < 1e-5


@pytest.mark.parametrize("M,N,K", [(256, 128, 128), (300, 96, 200), (1000, 384, 1536), (200, 192, 768), (128, 64, 64), (512, 1024, 256)])
@pytest.mark.parametrize("a_trans,b_trans", [(False, True), (True, True), (True, False)])
def test_tc_gemm_mn_major(M, N, K, a_trans, b_trans):
    import linnaeus_b200.functional as F

    a = torch.randn((K, M) if a_trans else (M, K), device=DEV).to(torch.bfloat16)
    b = torch.randn((K, N) if b_trans else (N, K), device=DEV).to(torch.bfloat16)
    out = F.gemm(a, b, M, N, K, a_trans=a_trans, b_trans=b_trans, out_dtype=torch.float32)
    assert rel_err(out, _ref(a, b, a_trans, b_trans)) < 1e-5


def test_tc_gemm_epilogues():
    import linnaeus_b200.functional as F

    M, N, K = 1000, 384, 96
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    b = (torch.randn(N, K, device=DEV) / 10).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    cs = torch.rand(N, device=DEV) + 0.5
    aux = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    out = F.gemm(a, b, M, N, K, bias=bias, act=1, aux_out=aux, residual=res, col_scale=cs)
    pre = a.float() @ b.float().t() + bias
    assert rel_err(aux, pre) < 1e-2
    assert rel_err(out, TF.gelu(pre) * cs + res.float()) < 1e-2
    out2 = F.gemm(a, b, M, N, K, act=1, act_grad_in=aux)  # aux / C share one dtype by contract
    u = aux.float().requires_grad_(True)
    TF.gelu(u).sum().backward()
    assert rel_err(out2, (a.float() @ b.float().t()) * u.grad) < 1e-2
    out3 = F.gemm(a, b, M, N, K, bias=bias, act=2, out_dtype=torch.float32)
    assert rel_err(out3, TF.relu(pre)) < 1e-5


@pytest.mark.parametrize("M,N,K", [(50000, 384, 96), (20000, 96, 384), (3000, 1576, 768), (256, 768, 768)])
def test_tc_wgrad_split_k(M, N, K):
    import linnaeus_b200.functional as F

    dy = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    x = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    dw = F.wgrad(dy, x)
    assert rel_err(dw, dy.float().t() @ x.float()) < 2e-5


def test_tc_matches_simt_on_same_inputs():
    import linnaeus_b200.functional as F

    a = torch.randn(700, 192, device=DEV).to(torch.bfloat16)
    b = torch.randn(768, 192, device=DEV).to(torch.bfloat16)
    o1 = F.gemm(a, b, 700, 768, 192, out_dtype=torch.float32)
    F.FORCE_SIMT = True
    try:
        o2 = F.gemm(a, b, 700, 768, 192, out_dtype=torch.float32)
    finally:
        F.FORCE_SIMT = False
    assert rel_err(o1, o2) < 1e-5


@pytest.mark.parametrize("M,N,K", [(4096, 384, 96), (5000, 96, 384), (3000, 768, 192), (2048, 192, 768), (1000, 1536, 384), (777, 1576, 768),
                                   (256, 768, 768), (130, 64, 48), (64, 8, 4), (9000, 2304, 768)])
@pytest.mark.parametrize("with_db", [True, False])
def test_wgrad_kernel(M, N, K, with_db):
    """lnx_wgrad: dW += dy^T x and db += colsum(dy) (ones-tile MMA), split-K atomics, accumulate semantics."""
    import linnaeus_b200.functional as F

    dy = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    ldx = (K + 7) // 8 * 8 + 8  # a column slice of a wider matrix
    xw = torch.randn(M, ldx, device=DEV).to(torch.bfloat16)
    x = xw[:, :K]
    dw = torch.full((N, K), 0.5, device=DEV)
    db = torch.full((N,), -1.0, device=DEV) if with_db else None
    F.wgrad(dy, x, x_ld=ldx, out=dw, db_out=db)
    ref = dy.float().t() @ x.float() + 0.5
    assert rel_err(dw, ref) < 2e-5
    if with_db:
        assert rel_err(db, dy.float().sum(0) - 1.0) < 2e-5


@pytest.mark.parametrize("M,N,K,bias", [(256, 128, 64, True), (1000, 256, 192, True), (4097, 384, 384, False), (12544, 768, 768, True), (300, 1152, 72, True)])
def test_gemm_on_cta_pairs_matches_fp32_torch(M, N, K, bias):
    """lnx_gemm_pair: tcgen05.mma.cta_group::2 on 256-row pair tiles (each CTA stages half of the B tile), ragged M, K not a multiple
    of the 64-wide k-block, block_n 256 and 128."""
    from linnaeus_b200._lib import call

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    call("lnx_gemm_pair", a.data_ptr(), w.data_ptr(), b.data_ptr() if bias else None, out.data_ptr(), M, N, K)
    ref = a.float() @ w.float().t() + (b if bias else 0.0)
    assert float((out.float() - ref).abs().max() / ref.abs().max()) < 8e-3
