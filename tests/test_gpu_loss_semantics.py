"""GPU: the fused hierarchical loss against the oracle (itself pinned to the unmodified reference in
tests/test_oracle_vs_reference.py) on the cases round 1 left unpinned: class-weight power in train / validation / PHASE1
(R/loss/masking.py:469-518,696-698, hierarchical_loss.py:313-334, gradient_weighting.py:334-352), the validation divisor under
ignore_index = 0 criteria, SoftTargetCrossEntropy (basic_loss.py:188-228) and criterion-level class weights (:76-90)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
TASKS = [("taxa_L10", 100), ("taxa_L20", 40), ("taxa_L30", 12), ("taxa_L40", 4)]


def rel_err(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


def _case(seed, B=37, soft=False):
    g = torch.Generator().manual_seed(seed)
    logits = {t: torch.randn(B, c, generator=g) * 2 for t, c in TASKS}
    targets = {t: torch.randint(0, c, (B,), generator=g) for t, c in TASKS}
    targets["taxa_L10"][:5] = 0
    targets["taxa_L30"][2] = 0
    if soft:
        for t, c in TASKS:
            a = torch.nn.functional.one_hot(targets[t], c).float()
            targets[t] = 0.65 * a + 0.35 * a[torch.randperm(B, generator=g)]
    cw = {t: 0.5 + torch.rand(c, generator=g) for t, c in TASKS}
    weights = {t: 0.5 + 0.3 * i for i, (t, _) in enumerate(TASKS)}
    return logits, targets, cw, weights


def _run_ours(kind, logits, targets, weights, cw=None, crit_w=None, phase1=False, is_validation=False, cw_train=True, cw_val=False):
    import linnaeus_b200.loss as LL
    from linnaeus_b200.config import get_default_config

    cfg = get_default_config()
    cfg.TRAIN.PHASE1_MASK_NULL_LOSS = phase1
    cfg.LOSS.GRAD_WEIGHTING.CLASS.TRAIN = cw_train
    cfg.LOSS.GRAD_WEIGHTING.CLASS.VAL = cw_val
    ign = 0 if phase1 else None
    cwt = crit_w or {}
    if kind == "ce":
        crit = {t: LL.CrossEntropyLoss(weight=cwt.get(t), apply_class_weights=t in cwt, ignore_index=ign) for t, _ in TASKS}
    elif kind == "ls":
        crit = {t: LL.LabelSmoothingCrossEntropy(weight=cwt.get(t), smoothing=0.1, apply_class_weights=t in cwt, ignore_index=ign) for t, _ in TASKS}
    else:
        crit = {t: LL.SoftTargetCrossEntropy(weight=cwt.get(t), apply_class_weights=t in cwt) for t, _ in TASKS}
    cw_dicts = {t: {i: float(v[i]) for i in range(v.numel())} for t, v in cw.items()} if cw else None
    tw = LL.StaticTaskWeighting([t for t, _ in TASKS], weights, class_weights=cw_dicts)
    lg = {t: v.clone().to(DEV).requires_grad_(True) for t, v in logits.items()}
    tg = {t: v.to(DEV) for t, v in targets.items()}
    total, comps, _ = LL.weighted_hierarchical_loss(lg, tg, crit, tw, None, 0, is_validation=is_validation, config=cfg)
    total.backward()
    return total, comps, lg, crit, tg


def _check(total, comps, lg, tot_o, comp_o, lo):
    assert abs(float(total) - float(tot_o)) <= 1e-5 * abs(float(tot_o))
    for t, _ in TASKS:
        assert rel_err(lg[t].grad.cpu(), lo[t].grad) < 1e-5
        assert abs(float(comps["weighted_tasks"][t]) - comp_o["weighted_tasks"][t]) <= 1e-5 * max(1.0, abs(comp_o["weighted_tasks"][t]))
        assert int(comps["null_masking"]["num_valid_samples_per_task"][t]) == comp_o["num_valid_samples_per_task"][t]


@pytest.mark.parametrize("kind", ["ce", "ls"])
@pytest.mark.parametrize("phase1,is_validation,cw_train,cw_val", [
    (False, False, True, False), (False, False, False, False), (True, False, True, False), (False, True, True, False),
    (False, True, True, True), (True, True, True, False)])
def test_class_weights_and_validation(kind, phase1, is_validation, cw_train, cw_val):
    from oracle import mformer_oracle as O

    logits, targets, cw, weights = _case(0)
    lo = {t: v.clone().requires_grad_(True) for t, v in logits.items()}
    tot_o, comp_o = O.hierarchical_loss(lo, targets, kind=kind, task_weights=weights, phase1_mask_null=phase1, is_validation=is_validation,
                                        class_weights=cw, apply_cw=(cw_val if is_validation else cw_train))
    tot_o.backward()
    total, comps, lg, _, _ = _run_ours(kind, logits, targets, weights, cw=cw, phase1=phase1, is_validation=is_validation, cw_train=cw_train,
                                       cw_val=cw_val)
    _check(total, comps, lg, tot_o, comp_o, lo)


@pytest.mark.parametrize("with_cw", [False, True])
def test_soft_target_cross_entropy(with_cw):
    from oracle import mformer_oracle as O

    logits, targets, cw, weights = _case(1, soft=True)
    lo = {t: v.clone().requires_grad_(True) for t, v in logits.items()}
    tot_o, comp_o = O.hierarchical_loss(lo, targets, kind="soft", task_weights=weights, class_weights=cw if with_cw else None, apply_cw=True)
    tot_o.backward()
    total, comps, lg, crit, tg = _run_ours("soft", logits, targets, weights, cw=cw if with_cw else None)
    _check(total, comps, lg, tot_o, comp_o, lo)
    per = crit["taxa_L20"](lg["taxa_L20"].detach(), tg["taxa_L20"])  # criterion API (per-sample vector)
    ref = O.per_sample_losses({"taxa_L20": logits["taxa_L20"]}, {"taxa_L20": targets["taxa_L20"]}, "soft")["taxa_L20"]
    assert rel_err(per.cpu(), ref) < 1e-5


@pytest.mark.parametrize("kind", ["ce", "ls", "soft"])
def test_criterion_level_class_weights(kind):
    from oracle import mformer_oracle as O

    logits, targets, cw, weights = _case(2, soft=(kind == "soft"))
    lo = {t: v.clone().requires_grad_(True) for t, v in logits.items()}
    tot_o, comp_o = O.hierarchical_loss(lo, targets, kind=kind, task_weights=weights, criterion_weights=cw)
    tot_o.backward()
    total, comps, lg, _, _ = _run_ours(kind, logits, targets, weights, crit_w=cw)
    _check(total, comps, lg, tot_o, comp_o, lo)
