"""GPU: every C-ABI kernel against a plain PyTorch fp32 reference of the same op
(forward and backward), in float32 (tight) and bfloat16 (2e-2) modes."""
import math

import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _F():
    import linnaeus_b200.functional as F

    return F


def rel_err(a, b):
    a, b = a.float(), b.float()
    a, b = a.detach(), b.detach()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def tol(dtype):
    return 2e-2 if dtype == torch.bfloat16 else 2e-5


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,C", [(1000, 96), (777, 192), (513, 384), (300, 768), (64, 32), (50, 2048), (33, 1024)])
def test_layernorm_fwd_bwd(dtype, rows, C):
    F = _F()
    x = torch.randn(rows, C, device=DEV).to(dtype).requires_grad_(True)
    res = torch.randn(rows, C, device=DEV).to(dtype).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(C, device=DEV)).requires_grad_(True)
    b = (0.1 * torch.randn(C, device=DEV)).requires_grad_(True)
    y = F.layernorm(x, w, b, 1e-6, residual=res)
    g = torch.randn_like(y)
    y.backward(g)
    xr = x.detach().float().requires_grad_(True)
    rr = res.detach().float().requires_grad_(True)
    wr, br = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = TF.layer_norm(xr, (C,), wr, br, 1e-6) + rr
    yr.backward(g.float())
    t = tol(dtype)
    assert rel_err(y, yr) < t
    assert rel_err(x.grad, xr.grad) < t
    assert rel_err(res.grad, rr.grad) < t
    assert rel_err(w.grad, wr.grad) < max(t, 1e-4)
    assert rel_err(b.grad, br.grad) < max(t, 1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,C", [(2, 56, 56, 96), (3, 28, 28, 192), (2, 16, 16, 32), (1, 20, 23, 64), (2, 14, 14, 256)])
def test_dwconv7_fwd_bwd(dtype, B, H, W, C):
    F = _F()
    x = torch.randn(B, H, W, C, device=DEV).to(dtype).requires_grad_(True)
    w = (0.2 * torch.randn(C, 1, 7, 7, device=DEV)).requires_grad_(True)
    b = (0.1 * torch.randn(C, device=DEV)).requires_grad_(True)
    y = F.dwconv7(x, w, b)
    g = torch.randn_like(y)
    y.backward(g)
    xr = x.detach().float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wr, br = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = TF.conv2d(xr, wr, br, padding=3, groups=C)
    yr.backward(g.float().permute(0, 3, 1, 2))
    t = tol(dtype)
    assert rel_err(y, yr.permute(0, 2, 3, 1)) < t
    assert rel_err(x.grad, xr.grad.permute(0, 2, 3, 1)) < t
    assert rel_err(w.grad, wr.grad) < max(t, 2e-4)
    assert rel_err(b.grad, br.grad) < max(t, 2e-4)


def _gemm_ref(a, b, a_trans, b_trans):
    A = a.float().t() if a_trans else a.float()
    Bm = b.float().t() if b_trans else b.float()
    return A @ Bm.t()


@pytest.mark.parametrize("ab", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(300, 200, 100), (128, 128, 64), (1000, 96, 384), (257, 1576, 768), (64, 384, 2), (5, 7, 3)])
@pytest.mark.parametrize("a_trans,b_trans", [(False, False), (False, True), (True, True), (True, False)])
def test_gemm_simt_layouts(ab, M, N, K, a_trans, b_trans):
    F = _F()
    F.FORCE_SIMT = True
    try:
        a = torch.randn((K, M) if a_trans else (M, K), device=DEV).to(ab)
        b = torch.randn((K, N) if b_trans else (N, K), device=DEV).to(ab)
        out = F.gemm(a, b, M, N, K, a_trans=a_trans, b_trans=b_trans, out_dtype=torch.float32)
        ref = _gemm_ref(a, b, a_trans, b_trans)
        assert rel_err(out, ref) < 1e-5
    finally:
        F.FORCE_SIMT = False


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_simt_epilogue_and_accumulate(dtype):
    F = _F()
    F.FORCE_SIMT = True
    try:
        M, N, K = 300, 136, 72
        a = torch.randn(M, K, device=DEV).to(dtype)
        b = torch.randn(N, K, device=DEV).to(dtype)
        bias = torch.randn(N, device=DEV)
        res = torch.randn(M, N, device=DEV).to(dtype)
        cs = torch.rand(N, device=DEV) + 0.5
        aux = torch.empty(M, N, device=DEV, dtype=dtype)
        out = F.gemm(a, b, M, N, K, bias=bias, act=1, aux_out=aux, residual=res, col_scale=cs)
        pre = a.float() @ b.float().t() + bias
        ref = TF.gelu(pre) * cs + res.float()
        t = tol(dtype)
        assert rel_err(aux, pre) < t
        assert rel_err(out, ref) < t
        # backward-through-activation epilogue
        out2 = F.gemm(a, b, M, N, K, act=1, act_grad_in=aux)
        u = aux.float().requires_grad_(True)
        TF.gelu(u).sum().backward()
        assert rel_err(out2, (a.float() @ b.float().t()) * u.grad) < t
        # split-K atomic accumulation (weight-gradient form)
        dy = torch.randn(5000, 40, device=DEV).to(dtype)
        x = torch.randn(5000, 24, device=DEV).to(dtype)
        dw = F.wgrad(dy, x)
        assert rel_err(dw, dy.float().t() @ x.float()) < 1e-4
    finally:
        F.FORCE_SIMT = False


@pytest.mark.parametrize("K,N", [(2, 384), (3, 96), (10, 768), (16, 50)])
def test_skinny_linear_paths(K, N):
    """The metadata-head Linear(K <= 16): fp32 columns of a pitched [B, 15] tensor, bias + ReLU epilogue, bf16 output, and its
    weight gradient dy^T x accumulated into an existing dW (R/models/utils.py metadata heads)."""
    F = _F()
    Bsz, pitch = 257, 17
    meta = torch.randn(Bsz, pitch, device=DEV)
    x = meta[:, 1:1 + K]
    w = torch.randn(N, K, device=DEV)
    bias = torch.randn(N, device=DEV)
    ref = torch.relu(x @ w.t() + bias)
    for od, t in ((torch.float32, 1e-6), (torch.bfloat16, 1e-2)):
        out = F.gemm(x, w, Bsz, N, K, lda=pitch, out_dtype=od, bias=bias, act=2)
        assert rel_err(out, ref) < t
    dy = torch.randn(Bsz, N, device=DEV)
    dw0 = torch.randn(N, K, device=DEV)
    dw = dw0.clone()
    F.wgrad(dy, x, x_ld=pitch, out=dw)
    torch.testing.assert_close(dw, dw0 + dy.t() @ x, rtol=1e-5, atol=1e-4)
    plain = F.gemm(dy, x, N, K, Bsz, a_trans=True, b_trans=True, lda=N, ldb=pitch, out_dtype=torch.float32)
    torch.testing.assert_close(plain, dy.t() @ x, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layout_kernels(dtype):
    F = _F()
    x = torch.randn(3, 3, 32, 48, device=DEV)
    a = F.patchify(x, 4, 64, dtype)
    ref = TF.unfold(x, kernel_size=4, stride=4).transpose(1, 2).reshape(-1, 48)
    assert torch.equal(a[:, :48].float(), ref.to(dtype).float())
    assert float(a[:, 48:].abs().max()) == 0.0
    y = torch.randn(2, 8, 12, 64, device=DEV).to(dtype).requires_grad_(True)
    s = F.space_to_depth(y)
    ref = y.detach().view(2, 4, 2, 6, 2, 64).permute(0, 1, 3, 2, 4, 5).reshape(2 * 4 * 6, 4 * 64)
    assert torch.equal(s, ref)
    g = torch.randn_like(s)
    s.backward(g)
    assert torch.equal(y.grad.view(2, 4, 2, 6, 2, 64).permute(0, 1, 3, 2, 4, 5).reshape(2 * 4 * 6, 4 * 64), g)
    # tokens
    cls = torch.randn(1, 1, 64, device=DEV, requires_grad=True)
    ex = torch.randn(2, 3, 64, device=DEV).to(dtype).requires_grad_(True)
    pt = torch.randn(2, 10, 64, device=DEV).to(dtype).requires_grad_(True)
    tok = F.tokens_assemble(cls, ex, pt)
    ref = torch.cat([cls.detach().to(dtype).expand(2, -1, -1), ex.detach(), pt.detach()], 1)
    assert torch.equal(tok, ref)
    g = torch.randn_like(tok)
    tok.backward(g)
    assert torch.equal(pt.grad, g[:, 4:])
    assert torch.equal(ex.grad, g[:, 1:4])
    assert rel_err(cls.grad.view(-1), g[:, 0].float().sum(0)) < 1e-6
    t2 = tok.detach().clone().requires_grad_(True)
    c, p = F.tokens_split(t2, 3)
    assert torch.equal(c, t2[:, 0]) and torch.equal(p, t2[:, 4:])
    gc, gp = torch.randn_like(c), torch.randn_like(p)
    (c * gc).sum().backward(retain_graph=True)
    assert torch.equal(t2.grad[:, 0], gc) and float(t2.grad[:, 1:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,heads,H,W,n_extra", [(2, 6, 14, 14, 4), (3, 2, 4, 4, 1), (1, 12, 7, 7, 4), (1, 4, 10, 13, 2)])
def test_rope_attention_fwd_bwd(dtype, B, heads, H, W, n_extra):
    F = _F()
    F.FORCE_SIMT = True
    try:
        hd = 64
        D = heads * hd
        N = H * W + n_extra
        qkv = (0.5 * torch.randn(B, N, 3 * D, device=DEV)).to(dtype).requires_grad_(True)
        freqs = (0.3 * torch.randn(2, heads, hd // 2, device=DEV)).requires_grad_(True)
        out = F.rope_attention(qkv, freqs, H, W, heads, n_extra)
        g = torch.randn_like(out)
        out.backward(g)

        qr = qkv.detach().float().requires_grad_(True)
        fr = freqs.detach().clone().requires_grad_(True)
        q, k, v = qr.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
        t = torch.arange(H * W, device=DEV, dtype=torch.float32)
        theta = (t % W)[:, None, None] * fr[0][None] + torch.div(t, W, rounding_mode="floor")[:, None, None] * fr[1][None]
        cos = torch.cos(theta).permute(1, 0, 2).repeat_interleave(2, dim=-1)
        fac = torch.cat([torch.ones(heads, n_extra, hd, device=DEV), cos], 1)[None]
        att = torch.softmax((q * fac * hd ** -0.5) @ (k * fac).transpose(-2, -1), -1)
        ref = (att @ v).transpose(1, 2).reshape(B, N, D)
        ref.backward(g.float())
        tl = tol(dtype) if dtype == torch.bfloat16 else 1e-4
        assert rel_err(out, ref) < tl
        assert rel_err(qkv.grad, qr.grad) < tl
        assert rel_err(freqs.grad, fr.grad) < max(tl, 1e-3)
    finally:
        F.FORCE_SIMT = False


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act,res,cs", [("gelu", True, True), ("gelu", True, False), ("gelu", False, False)])
def test_mlp2_and_linear_autograd(dtype, act, res, cs):
    F = _F()
    F.FORCE_SIMT = True
    try:
        M, K, Hd, N = 200, 64, 256, 64
        x = torch.randn(M, K, device=DEV).to(dtype).requires_grad_(True)
        w1 = (torch.randn(Hd, K, device=DEV) / math.sqrt(K)).requires_grad_(True)
        b1 = (0.1 * torch.randn(Hd, device=DEV)).requires_grad_(True)
        w2 = (torch.randn(N, Hd, device=DEV) / math.sqrt(Hd)).requires_grad_(True)
        b2 = (0.1 * torch.randn(N, device=DEV)).requires_grad_(True)
        r = torch.randn(M, N, device=DEV).to(dtype).requires_grad_(True) if res else None
        gamma = (torch.rand(N, device=DEV) + 0.5).requires_grad_(True) if cs else None
        y = F.mlp2(x, w1, b1, w2, b2, act=act, residual=r, col_scale=gamma)
        g = torch.randn_like(y)
        y.backward(g)
        P = [t.detach().float().clone().requires_grad_(True) if t is not None else None for t in (x, w1, b1, w2, b2, r, gamma)]
        xr, w1r, b1r, w2r, b2r, rr, gr = P
        yr = TF.linear(TF.gelu(TF.linear(xr, w1r, b1r)), w2r, b2r)
        if gr is not None:
            yr = yr * gr
        if rr is not None:
            yr = yr + rr
        yr.backward(g.float())
        t = tol(dtype) if dtype == torch.bfloat16 else 1e-4
        assert rel_err(y, yr) < t
        for a, b_ in zip((x, w1, b1, w2, b2, r, gamma), P):
            if a is not None:
                assert rel_err(a.grad, b_.grad) < t
        # single linear with relu + residual
        x2 = torch.randn(M, K, device=DEV).to(dtype).requires_grad_(True)
        rr2 = torch.randn(M, Hd, device=DEV).to(dtype).requires_grad_(True)
        y2 = F.linear(x2, w1, b1, act="relu", residual=rr2)
        g2 = torch.randn_like(y2)
        w1.grad = None
        y2.backward(g2)
        xr2 = x2.detach().float().requires_grad_(True)
        # ReLU masks flip on ~0 pre-activations, so the reference uses the same (rounded) weights
        w1q = w1.detach().to(dtype).float().requires_grad_(True)
        yr2 = TF.relu(TF.linear(xr2, w1q, b1.detach())) + rr2.detach().float()
        yr2.backward(g2.float())
        assert rel_err(y2, yr2) < t and rel_err(x2.grad, xr2.grad) < t and rel_err(w1.grad, w1q.grad) < t
    finally:
        F.FORCE_SIMT = False


@pytest.mark.parametrize("kind", ["ce", "ls", "taxonomy"])
@pytest.mark.parametrize("phase1,null_p", [(False, 1.0), (True, 1.0), (False, 0.0)])
def test_fused_loss_matches_oracle(kind, phase1, null_p):
    from oracle import mformer_oracle as O
    import linnaeus_b200.loss as LL
    from linnaeus_b200.config import get_default_config

    tasks = [("taxa_L10", 100), ("taxa_L20", 40), ("taxa_L30", 12), ("taxa_L40", 4)]
    B = 37
    g = torch.Generator().manual_seed(1)
    logits = {t: torch.randn(B, c, generator=g) * 2 for t, c in tasks}
    targets = {t: torch.randint(0, c, (B,), generator=g) for t, c in tasks}
    targets["taxa_L10"][:5] = 0
    weights = {t: 0.5 + 0.3 * i for i, (t, _) in enumerate(tasks)}
    mats = O.synthetic_taxonomy_smoothing(tasks) if kind == "taxonomy" else None
    lo = {t: v.clone().requires_grad_(True) for t, v in logits.items()}
    tot_o, comp_o = O.hierarchical_loss(lo, targets, kind=kind, task_weights=weights, null_mask_prob=null_p, phase1_mask_null=phase1,
                                        soft_matrices=mats, coin_flips={t: torch.zeros(B, dtype=torch.bool) for t, _ in tasks})
    tot_o.backward()

    cfg = get_default_config()
    cfg.TRAIN.PHASE1_MASK_NULL_LOSS = phase1
    lg = {t: v.clone().to(DEV).requires_grad_(True) for t, v in logits.items()}
    tg = {t: v.to(DEV) for t, v in targets.items()}
    ign = 0 if phase1 else None
    if kind == "ce":
        crit = {t: LL.CrossEntropyLoss(ignore_index=ign) for t, _ in tasks}
    elif kind == "ls":
        crit = {t: LL.LabelSmoothingCrossEntropy(smoothing=0.1, ignore_index=ign) for t, _ in tasks}
    else:
        crit = {t: LL.TaxonomyAwareLabelSmoothingCE(mats[t], ignore_index=ign).to(DEV) for t, _ in tasks}
    tw = LL.StaticTaskWeighting([t for t, _ in tasks], weights)

    class Sched:
        def get_null_mask_prob(self, step):
            return null_p

    total, comps, _ = LL.weighted_hierarchical_loss(lg, tg, crit, tw, Sched(), 0, config=cfg)
    total.backward()
    assert abs(float(total) - float(tot_o)) <= 1e-5 * abs(float(tot_o))
    for t, _ in tasks:
        assert rel_err(lg[t].grad.cpu(), lo[t].grad) < 1e-5
        assert abs(float(comps["tasks"][t]) - comp_o["tasks"][t]) < 1e-4
        assert int(comps["null_masking"]["num_valid_samples_per_task"][t]) == comp_o["num_valid_samples_per_task"][t]
    # criterion API (per-sample vector)
    per = crit["taxa_L20"](lg["taxa_L20"].detach(), tg["taxa_L20"])
    ref = O.per_sample_losses({"taxa_L20": logits["taxa_L20"]}, {"taxa_L20": targets["taxa_L20"]}, kind, 0.1, mats, ign)["taxa_L20"]
    assert rel_err(per.cpu(), ref) < 1e-5


def test_flat_adamw_matches_torch():
    from linnaeus_b200.optim import FlatAdamW

    torch.manual_seed(3)
    shapes = {"a.weight": (33, 17), "a.bias": (33,), "b.weight": (5, 33), "norm.weight": (33,), "tok": (1, 1, 7)}
    ps = {n: torch.nn.Parameter(torch.randn(s, device=DEV)) for n, s in shapes.items()}
    ref = {n: torch.nn.Parameter(p.detach().clone()) for n, p in ps.items()}
    opt = FlatAdamW(ps.items(), lr=3e-3, weight_decay=0.05, clip_grad=1.0)
    decay = [ref[n] for n in ("a.weight", "b.weight", "tok")]
    nodecay = [ref[n] for n in ("a.bias", "norm.weight")]
    ropt = torch.optim.AdamW([{"params": decay}, {"params": nodecay, "weight_decay": 0.0}], lr=3e-3, weight_decay=0.05)
    for step in range(3):
        opt.zero_grad()
        for n in shapes:
            gr = torch.randn(shapes[n], device=DEV) * (step + 1)
            ps[n].grad.add_(gr)
            ref[n].grad = gr.clone()
        nr = torch.nn.utils.clip_grad_norm_(list(ref.values()), 1.0)
        ropt.step()
        opt.step()
        assert abs(float(opt.grad_norm) - float(nr)) < 1e-4 * float(nr)
    for n in shapes:
        assert rel_err(ps[n], ref[n]) < 1e-5


def test_flat_adamw_state_dict_interchanges_with_torch_adamw(tmp_path):
    """FlatAdamW.state_dict() has torch.optim.AdamW's layout (R/utils/checkpoint.py saves / loads that): after two steps each
    side loads the OTHER side's state dict (through a checkpoint file), takes two more steps, and the parameters still agree."""
    from linnaeus_b200 import checkpoint as C
    from linnaeus_b200.optim import FlatAdamW

    torch.manual_seed(4)
    shapes = {"a.weight": (33, 17), "a.bias": (33,), "b.weight": (5, 33), "norm.weight": (33,), "tok": (1, 1, 7)}
    grads = [{n: torch.randn(s, device=DEV) for n, s in shapes.items()} for _ in range(4)]

    class Holder(torch.nn.Module):
        def __init__(self, init):
            super().__init__()
            self.ps = torch.nn.ParameterDict({n.replace(".", "_"): torch.nn.Parameter(v.clone()) for n, v in init.items()})

    init = {n: torch.randn(s, device=DEV) for n, s in shapes.items()}

    def make_flat(h):
        return FlatAdamW([(n, h.ps[n.replace(".", "_")]) for n in shapes], lr=3e-3, weight_decay=0.05)

    def make_torch(h):
        decay = [h.ps[n.replace(".", "_")] for n in ("a.weight", "b.weight", "tok")]
        nodecay = [h.ps[n.replace(".", "_")] for n in ("a.bias", "norm.weight")]
        return torch.optim.AdamW([{"params": decay}, {"params": nodecay, "weight_decay": 0.0}], lr=3e-3, weight_decay=0.05)

    def run(h, opt, steps):
        for g in steps:
            opt.zero_grad()
            for n in shapes:
                p = h.ps[n.replace(".", "_")]
                if p.grad is None:
                    p.grad = g[n].clone()
                else:
                    p.grad.add_(g[n])
            opt.step()

    hf, ht = Holder(init), Holder(init)
    of, ot = make_flat(hf), make_torch(ht)
    assert of.state_dict()["state"] == {}  # like torch before the first step
    run(hf, of, grads[:2])
    run(ht, ot, grads[:2])
    sdf = of.state_dict()
    assert set(sdf) == {"state", "param_groups"} and set(sdf["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert [len(g["params"]) for g in sdf["param_groups"]] == [3, 2] and float(sdf["state"][4]["step"]) == 2.0
    C.save_checkpoint(str(tmp_path / "flat"), hf, of, epoch=1)
    C.save_checkpoint(str(tmp_path / "torch"), ht, ot, epoch=1)
    # cross-load: the torch-written file resumes a FlatAdamW run and vice versa
    hf2, ht2 = Holder(init), Holder(init)
    of2, ot2 = make_flat(hf2), make_torch(ht2)
    C.load_checkpoint(str(tmp_path / "torch" / "latest.pth"), hf2, of2, map_location=DEV)
    C.load_checkpoint(str(tmp_path / "flat" / "latest.pth"), ht2, ot2, map_location=DEV)
    run(hf2, of2, grads[2:])
    run(ht2, ot2, grads[2:])
    run(hf, of, grads[2:])
    for n in shapes:
        k = n.replace(".", "_")
        assert rel_err(hf2.ps[k], ht2.ps[k]) < 1e-5
        assert rel_err(hf2.ps[k], hf.ps[k]) < 1e-5  # resumed == uninterrupted


def test_aggregate_and_colsum():
    F = _F()
    a = torch.randn(9, 40, device=DEV, requires_grad=True)
    c = torch.randn(9, 40, device=DEV, requires_grad=True)
    w = torch.tensor([0.3, -0.7], device=DEV).view(1, 2, 1).requires_grad_(True)
    b = torch.tensor([0.2], device=DEV, requires_grad=True)
    out = F.aggregate2(a, c, w, b)
    g = torch.randn_like(out)
    out.backward(g)
    # CPU reference: cuDNN's conv1d would run in TF32
    ar, cr = a.detach().cpu().requires_grad_(True), c.detach().cpu().requires_grad_(True)
    wr, br = w.detach().cpu().requires_grad_(True), b.detach().cpu().requires_grad_(True)
    ref = TF.conv1d(torch.stack([ar, cr], 1), wr, br).squeeze(1)
    ref.backward(g.cpu())
    assert rel_err(out.cpu(), ref) < 1e-6 and rel_err(a.grad.cpu(), ar.grad) < 1e-6 and rel_err(c.grad.cpu(), cr.grad) < 1e-6
    assert rel_err(w.grad.cpu(), wr.grad) < 1e-5 and rel_err(b.grad.cpu(), br.grad) < 1e-5
    x = torch.randn(1234, 77, device=DEV)
    assert rel_err(F.colsum(x), x.sum(0)) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fork_ops_fuse_the_skip_gradient(dtype):
    """dwconv7_fork / layernorm_fork return (op(x), x); the gradient arriving on the second output must be added to
    the op's input gradient inside the backward kernel (same result as autograd's separate add)."""
    F = _F()
    torch.manual_seed(3)
    B, H, W, C = 2, 28, 28, 96
    x0 = torch.randn(B, H, W, C, device=DEV).to(dtype)
    w = (0.2 * torch.randn(C, 1, 7, 7, device=DEV)).requires_grad_(True)
    b = (0.1 * torch.randn(C, device=DEV)).requires_grad_(True)
    g1, g2 = torch.randn_like(x0), torch.randn_like(x0)
    xa = x0.clone().requires_grad_(True)
    y, skip = F.dwconv7_fork(xa, w, b)
    (y * g1).sum().backward(retain_graph=True, inputs=[xa]) if False else torch.autograd.backward([y, skip], [g1, g2])
    xb = x0.clone().requires_grad_(True)
    yb = F.dwconv7(xb, w, b)
    torch.autograd.backward([yb], [g1])
    ref = xb.grad.float() + g2.float()
    assert rel_err(y, yb) == 0.0
    assert rel_err(xa.grad, ref) < tol(dtype)

    lw = torch.randn(C, device=DEV, requires_grad=True)
    lb = torch.randn(C, device=DEV, requires_grad=True)
    xa = x0.view(-1, C).clone().requires_grad_(True)
    y, skip = F.layernorm_fork(xa, lw, lb, 1e-6)
    torch.autograd.backward([y, skip], [g1.view(-1, C), g2.view(-1, C)])
    xb = x0.view(-1, C).clone().requires_grad_(True)
    yb = F.layernorm(xb, lw, lb, 1e-6)
    torch.autograd.backward([yb], [g1.view(-1, C)])
    ref = xb.grad.float() + g2.view(-1, C).float()
    assert rel_err(y, yb) == 0.0
    assert rel_err(xa.grad, ref) < tol(dtype)
