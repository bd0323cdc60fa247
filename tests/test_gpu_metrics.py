"""GPU: lnx_hier_metrics / lnx_hier_topk through the host mirror (linnaeus_b200.metrics) against the CPU oracle
(oracle/metrics_oracle.py) and the committed reference outputs (tests/golden/metrics_*.npz).  Counting metrics: exact."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO
from tests.support.golden_metrics import CASES, load_golden, make_case

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _M():
    import linnaeus_b200.metrics as M

    return M


def _dev(d, dtype=None):
    return {k: torch.from_numpy(v).to(DEV) if dtype is None else torch.from_numpy(v).to(DEV).to(dtype) for k, v in d.items()}


@pytest.mark.parametrize("name", list(CASES))
def test_phase_accumulator_matches_reference_golden(name):
    M = _M()
    keys, batches = make_case(name)
    g = load_golden(name)
    acc = M.HierMetricsAccumulator()
    for bi, (outputs, targets) in enumerate(batches):
        o, t = _dev(outputs), _dev(targets)
        acc.update(o, t)
        ol, tl = [o[k] for k in keys], [t[k] for k in keys]
        assert M.compute_chain_accuracy_vectorized(ol, tl) == pytest.approx(float(g[f"b{bi}.chain"]), abs=1e-12)
        assert M.compute_chain_accuracy_vectorized(ol, tl, ignore_index=0) == 0.0
        assert M.compute_partial_chain_accuracy_vectorized(ol, tl) == pytest.approx(float(g[f"b{bi}.partial"]), abs=1e-12)
        for k in keys:
            C = outputs[k].shape[1]
            ks = tuple(x for x in (1, 3, 5) if x <= C)
            np.testing.assert_allclose(M.accuracy(o[k], t[k], ks), g[f"b{bi}.{k}.acc"], rtol=0, atol=1e-4)
            np.testing.assert_allclose(M.accuracy(o[k], t[k], ks, ignore_index=0), g[f"b{bi}.{k}.acc_ignore0"], rtol=0, atol=1e-4)
    m = acc.compute()
    np.testing.assert_allclose([m["acc1"][k] for k in keys], g["acc1"], rtol=0, atol=1e-9)
    np.testing.assert_allclose([m["acc3"][k] for k in keys], g["acc3"], rtol=0, atol=1e-9)
    assert m["chain_accuracy"] == pytest.approx(float(g["chain_accuracy"]), abs=1e-12)
    assert m["partial_chain_accuracy"] == pytest.approx(float(g["partial_chain_accuracy"]), abs=1e-12)
    assert m["samples"] == sum(CASES[name][0])
    for f in ("null_acc1", "non_null_acc1"):  # the tracker's null / non-null split (restated inline in make_golden_metrics.py)
        ref = {k: float(v) for k, v in zip(keys, g[f]) if not np.isnan(v)}
        assert m[f] == pytest.approx(ref, abs=1e-9)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,classes", [(1, (3,)), (257, (1000, 400, 120, 40, 12, 4)), (64, (2, 1, 70000)), (5, tuple([4] * 16))])
def test_counters_and_ranks_bit_exact_vs_oracle(dtype, B, classes):
    """Heavy ties (bf16 logits quantised to a handful of values), nulls, one-class heads, a 70k-class head, 16 tasks."""
    M = _M()
    rng = np.random.default_rng(B + len(classes))
    outs, tgts = [], []
    for C in classes:
        z = np.round(rng.standard_normal((B, C)) * 2).astype(np.float32) / 2  # multiples of 0.5: many exact ties, bf16-exact
        y = rng.integers(0, C, size=B).astype(np.int64)
        y[rng.random(B) < 0.3] = 0
        z[np.arange(B), y] += rng.integers(0, 4, size=B).astype(np.float32)
        outs.append(z)
        tgts.append(y)
    offs = np.concatenate([[0], np.cumsum(classes)]).tolist()
    pad = 3  # a row pitch larger than sum C_k (the model's head-GEMM output may be padded)
    cat = torch.zeros(B, offs[-1] + pad, device=DEV, dtype=dtype)
    cat[:, : offs[-1]] = torch.from_numpy(np.concatenate(outs, axis=1)).to(DEV).to(dtype)
    tg = torch.from_numpy(np.stack(tgts)).to(DEV)
    ranks, _ = M.hier_metrics(cat, offs, tg, want_ranks=True)
    ref_r = np.stack([MO.target_rank(o, y) for o, y in zip(outs, tgts)])
    assert np.array_equal(ranks.cpu().numpy(), ref_r)
    counters = torch.full((4 * len(classes) + 4,), 7, dtype=torch.int64, device=DEV)  # ADDED to, not overwritten
    M.hier_metrics(cat, offs, tg, counters=counters)
    assert np.array_equal(counters.cpu().numpy() - 7, MO.phase_counters(outs, tgts))
    # one-hot targets are arg-maxed by the host mirror
    onehots = [torch.from_numpy(np.eye(C, dtype=np.float32)[y]).to(DEV) for C, y in zip(classes, tgts)]
    ol = [cat[:, offs[i]:offs[i + 1]] for i in range(len(classes))]
    assert M.compute_chain_accuracy_vectorized(ol, onehots) == pytest.approx(MO.chain_accuracy(outs, tgts), abs=1e-12)
    assert M.compute_partial_chain_accuracy_vectorized(ol, onehots) == pytest.approx(MO.partial_chain_accuracy(outs, tgts), abs=1e-12)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_topk_predictions_vs_oracle_and_golden(dtype):
    M = _M()
    keys, batches = make_case("metrics_six")
    g = load_golden("metrics_six")
    for bi, (outputs, _) in enumerate(batches):
        o = _dev(outputs, dtype)
        res = M.topk_predictions(o, k=5)
        for k in keys:
            zin = o[k].float().cpu().numpy()  # what the kernel saw (bf16-rounded in the bf16 case)
            idx, prob = MO.softmax_topk(zin, 5)
            assert np.array_equal(res[k][0].numpy(), idx)
            np.testing.assert_allclose(res[k][1].numpy(), prob, rtol=2e-5, atol=1e-7)
            if dtype == torch.float32:  # the unmodified reference call sequence (tie-free inputs)
                assert np.array_equal(res[k][0].numpy(), g[f"b{bi}.{k}.topk_idx"])
                np.testing.assert_allclose(res[k][1].numpy(), g[f"b{bi}.{k}.topk_prob"], rtol=2e-5, atol=1e-7)
    # ties and k > C: (value desc, index asc), short heads are truncated
    z = {"taxa_L10": torch.tensor([[1.0, 3.0, 3.0, 3.0, 0.0]], device=DEV), "taxa_L20": torch.tensor([[0.5, 0.5]], device=DEV)}
    r = M.topk_predictions(z, k=4)
    assert r["taxa_L10"][0].tolist() == [[1, 2, 3, 0]] and r["taxa_L20"][0].tolist() == [[0, 1]]
    assert r["taxa_L20"][1].tolist() == [[0.5, 0.5]]


def test_metrics_on_model_outputs_use_the_fused_logits():
    """The accumulator consumes the model's LogitsDict (one head GEMM output, class offsets attached) without a concat."""
    import linnaeus_b200 as L
    from linnaeus_b200.config import make_synthetic_config

    M = _M()
    cfg, nc = make_synthetic_config("sm", 64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(1, 1), conv_depths=(1, 1, 1, 1), n_tasks=3)
    model = L.build_model(cfg, num_classes=nc).to(DEV).eval()
    x = torch.randn(9, 3, 64, 64, device=DEV)
    meta = torch.randn(9, 15, device=DEV)
    with torch.no_grad():
        out = model(x, meta)
    assert getattr(out, "cat", None) is not None
    keys = list(out.keys())
    tg = {k: out[k].float().argmax(1) for k in keys}
    tg[keys[0]][:4] = (tg[keys[0]][:4] + 1) % out[keys[0]].shape[1]  # four samples wrong on the first task
    acc = M.HierMetricsAccumulator()
    acc.update(out, tg)
    m = acc.compute()
    ref = MO.phase_metrics([({k: out[k].float().cpu().numpy() for k in keys}, {k: tg[k].cpu().numpy() for k in keys})], keys)
    assert m["samples"] == 9 and m["chain_accuracy"] == pytest.approx(5 / 9)
    for f in ("chain_accuracy", "partial_chain_accuracy"):
        assert m[f] == pytest.approx(ref[f], abs=1e-12)
    assert m["acc1"] == pytest.approx(ref["acc1"]) and m["acc3"] == pytest.approx(ref["acc3"])


def test_cpu_tensors_are_rejected():
    M = _M()
    with pytest.raises(RuntimeError):
        M.accuracy(torch.randn(4, 5), torch.zeros(4, dtype=torch.int64))


def test_validate_one_pass_matches_per_batch_evaluation():
    """engine.validate_one_pass (no per-batch host sync) against the same loop done by hand with the reference-named metric functions
    and the loss called per batch; mask_meta changes the logits, so it must change the loss."""
    import linnaeus_b200 as L
    import linnaeus_b200.loss as LL
    from linnaeus_b200.config import make_synthetic_config
    from linnaeus_b200.engine import validate_one_pass

    M = _M()
    torch.manual_seed(0)
    cfg, nc = make_synthetic_config("sm", 64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(1, 1), conv_depths=(1, 1, 1, 1), n_tasks=3)
    model = L.build_model(cfg, nc).to(DEV).train()
    keys = list(nc.keys())
    crit = {k: LL.CrossEntropyLoss() for k in keys}
    tw = LL.StaticTaskWeighting(keys)
    batches = []
    for b in (5, 8, 3):
        tg = {k: torch.randint(0, nc[k], (b,), device=DEV) for k in keys}
        batches.append((torch.randn(b, 3, 64, 64, device=DEV), tg, torch.randn(b, 15, device=DEV)))
    res = validate_one_pass(cfg, model, batches, crit, tw)
    assert model.training  # restored
    model.eval()
    losses, chain, part, tot, c1 = [], 0.0, 0.0, 0, {k: 0 for k in keys}
    with torch.no_grad():
        for x, tg, aux in batches:
            out = model(x, aux)
            losses.append(float(LL.weighted_hierarchical_loss(out, tg, crit, tw, None, 0, is_validation=True, config=cfg)[0]))
            ol, tl = [out[k] for k in keys], [tg[k] for k in keys]
            chain += M.compute_chain_accuracy_vectorized(ol, tl) * x.shape[0]
            part += M.compute_partial_chain_accuracy_vectorized(ol, tl) * x.shape[0]
            tot += x.shape[0]
            for k in keys:
                c1[k] += int((out[k].float().argmax(1) == tg[k]).sum())
    assert res["batches"] == 3 and res["samples"] == tot
    assert res["loss"] == pytest.approx(sum(losses) / 3, rel=1e-5)
    assert res["chain_accuracy"] == pytest.approx(chain / tot, abs=1e-12) and res["partial_chain_accuracy"] == pytest.approx(part / tot, abs=1e-12)
    assert res["acc1"] == pytest.approx({k: 100.0 * c1[k] / tot for k in keys})
    masked = validate_one_pass(cfg, model, batches, crit, tw, mask_meta=True)
    assert masked["loss"] != res["loss"] and masked["samples"] == tot
