"""GPU: tcgen05 attention (forward + backward) against the fp32 torch reference and the
CUDA-core kernel on the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _ref(q, k, v):
    att = torch.softmax(q.float() @ k.float().transpose(-2, -1), -1)
    return att @ v.float(), att


@pytest.mark.parametrize("B,heads,N", [(2, 6, 200), (3, 12, 53), (1, 2, 256), (2, 3, 128), (1, 4, 17), (1, 2, 129), (1, 1, 64), (1, 2, 300), (2, 3, 580),
                                        (1, 2, 1000), (1, 1, 241)])
def test_attn_tc_forward_backward(B, heads, N):
    from linnaeus_b200._lib import call

    torch.manual_seed(N)
    hd = 64
    q = (torch.randn(B, heads, N, hd, device=DEV) * 0.35).to(torch.bfloat16)
    k = torch.randn(B, heads, N, hd, device=DEV).to(torch.bfloat16)
    v = torch.randn(B, heads, N, hd, device=DEV).to(torch.bfloat16)
    do = torch.randn(B, N, heads * hd, device=DEV).to(torch.bfloat16)

    def run(force_simt):
        out = torch.empty(B, N, heads * hd, device=DEV, dtype=torch.bfloat16)
        lse = torch.empty(B, heads, N, device=DEV)
        call("lnx_attn_fwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(), B, heads, N, hd, 1, int(force_simt))
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ws = torch.empty(B * heads * N * (hd + 1) + 4, device=DEV)
        call("lnx_attn_bwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(),
             dk.data_ptr(), dv.data_ptr(), ws.data_ptr(), B, heads, N, hd, 1, int(force_simt))
        return out, lse, dq, dk, dv

    out, lse, dq, dk, dv = run(False)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    ref, att = _ref(qr, kr, vr)
    ref_o = ref.transpose(1, 2).reshape(B, N, heads * hd)
    ref_o.backward(do.float())
    ref_lse = torch.logsumexp(qr.detach() @ kr.detach().transpose(-2, -1), -1)
    assert rel_err(out, ref_o) < 1.5e-2
    assert rel_err(lse, ref_lse) < 1e-4
    assert rel_err(dv, vr.grad) < 2e-2
    assert rel_err(dk, kr.grad) < 2e-2
    assert rel_err(dq, qr.grad) < 2e-2
    o2, l2, dq2, dk2, dv2 = run(True)
    assert rel_err(out, o2) < 1.5e-2 and rel_err(dq, dq2) < 2e-2 and rel_err(dk, dk2) < 2e-2 and rel_err(dv, dv2) < 2e-2


def test_attn_long_sequence_falls_back_forward_but_tc_backward():
    """N = 580 (xl @ 384^2): the forward tile (<= 256 keys) does not apply; results must still be right."""
    from linnaeus_b200._lib import call

    B, heads, N, hd = 1, 2, 580, 64
    torch.manual_seed(1)
    q = (torch.randn(B, heads, N, hd, device=DEV) * 0.3).to(torch.bfloat16)
    k = torch.randn(B, heads, N, hd, device=DEV).to(torch.bfloat16)
    v = torch.randn(B, heads, N, hd, device=DEV).to(torch.bfloat16)
    do = torch.randn(B, N, heads * hd, device=DEV).to(torch.bfloat16)
    out = torch.empty(B, N, heads * hd, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, heads, N, device=DEV)
    call("lnx_attn_fwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(), B, heads, N, hd, 1, 0)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ws = torch.empty(B * heads * N * (hd + 1) + 4, device=DEV)
    call("lnx_attn_bwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(),
         dk.data_ptr(), dv.data_ptr(), ws.data_ptr(), B, heads, N, hd, 1, 0)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    ref, _ = _ref(qr, kr, vr)
    ref.transpose(1, 2).reshape(B, N, heads * hd).backward(do.float())
    assert rel_err(out, ref.transpose(1, 2).reshape(B, N, heads * hd)) < 1.5e-2
    assert rel_err(dq, qr.grad) < 2e-2 and rel_err(dk, kr.grad) < 2e-2 and rel_err(dv, vr.grad) < 2e-2
