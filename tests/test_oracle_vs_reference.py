"""Pins the CPU oracle to the *real* reference (build container only).

The reference has no golden vectors for this path, so the
oracle is pinned by executing ``/root/reference`` itself: forward, hierarchical
loss (all criteria + null masking + PHASE1 + task weights), every parameter
gradient, and one clip+AdamW step.  Skipped when the reference tree is absent
(GPU box); there the committed ``tests/golden/*.npz`` carry the pin.
"""
import copy

import pytest
import torch

from oracle import mformer_oracle as O
from tests.support import refload

pytestmark = pytest.mark.skipif(not refload.reference_available(), reason="reference tree not present")

TINY = dict(variant="sm", img_size=64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(2, 1), conv_depths=(1, 1, 1, 1))


def _build(head_type="Linear", meta=True, n_tasks=6, **kw):
    refload.import_reference()
    from linnaeus.models import build_model

    cfg, nc = refload.reference_config(head_type=head_type, meta=meta, n_tasks=n_tasks, **kw)
    tree = refload.synthetic_taxonomy_tree(nc) if head_type != "Linear" else None
    model = build_model(cfg, num_classes=nc, taxonomy_tree=tree)
    a = O.arch_from_config(cfg, nc)
    return cfg, nc, model, a


def _load(model, a, seed=0):
    shapes = O.param_shapes(a)
    ref = {k: tuple(v.shape) for k, v in model.state_dict().items() if not k.split(".")[-1].startswith("hmatrix")}
    assert list(ref.keys()) == list(shapes.keys())
    assert ref == shapes
    P = O.synth_state_dict(shapes, seed)
    model.load_state_dict(P, strict=False)
    return P


def _ref_loss(cfg, logits, targets, kind="ce", weights=None, null_p=1.0, phase1=False, a=None, mats=None):
    from linnaeus.loss.basic_loss import CrossEntropyLoss, LabelSmoothingCrossEntropy
    from linnaeus.loss.gradient_weighting import GradientWeighting
    from linnaeus.loss.hierarchical_loss import weighted_hierarchical_loss
    from linnaeus.loss.taxonomy_label_smoothing import TaxonomyAwareLabelSmoothingCE

    cfg = cfg.clone()
    cfg.defrost()
    cfg.TRAIN.PHASE1_MASK_NULL_LOSS = phase1
    ign = 0 if phase1 else None
    keys = list(logits.keys())
    if kind == "ce":
        crit = {k: CrossEntropyLoss(ignore_index=ign) for k in keys}
    elif kind == "ls":
        crit = {k: LabelSmoothingCrossEntropy(smoothing=0.1, ignore_index=ign) for k in keys}
    else:
        crit = {k: TaxonomyAwareLabelSmoothingCE(mats[k], ignore_index=ign) for k in keys}
    gw = GradientWeighting(keys, cfg, "static", init_weights=weights, class_weights=None)

    class Sched:
        def get_null_mask_prob(self, step):
            return null_p

    return weighted_hierarchical_loss(logits, targets, crit, gw, Sched(), 0, config=cfg)


def test_forward_sm_224_matches_reference():
    cfg, nc, model, a = _build(variant="sm", img_size=224)
    P = _load(model, a)
    x, meta, _ = O.synth_batch(a, 2, 0)
    model.eval()
    with torch.no_grad():
        r = model(x, meta)
        o = O.forward(P, a, x, meta)
    assert list(r.keys()) == list(o.keys())
    for k in r:
        torch.testing.assert_close(o[k], r[k], rtol=1e-5, atol=2e-5)
        assert torch.equal(o[k].argmax(1), r[k].argmax(1))


@pytest.mark.parametrize("head_type", ["Linear", "HierarchicalSoftmax", "ConditionalClassifier"])
@pytest.mark.parametrize("meta", [True, False])
def test_forward_tiny_variants(head_type, meta):
    cfg, nc, model, a = _build(head_type=head_type, meta=meta, **TINY)
    P = _load(model, a, seed=3)
    x, m, _ = O.synth_batch(a, 3, 1)
    model.eval()
    with torch.no_grad():
        r = model(x, m)
        o = O.forward(P, a, x, m)
    for k in r:
        torch.testing.assert_close(o[k], r[k], rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("kind,null_p,phase1", [("ce", 1.0, False), ("ce", 1.0, True), ("ls", 1.0, False), ("taxonomy", 1.0, False), ("ce", 0.0, False)])
def test_loss_and_grads_match_reference(kind, null_p, phase1):
    cfg, nc, model, a = _build(**TINY)
    P = _load(model, a, seed=1)
    x, meta, tg = O.synth_batch(a, 8, 2)
    tg["taxa_L10"][:3] = 0  # make sure nulls exist
    weights = {t: 0.5 + 0.25 * i for i, (t, _) in enumerate(a.tasks)}
    mats = O.synthetic_taxonomy_smoothing(a.tasks) if kind == "taxonomy" else None

    model.train()
    r_logits = model(x, meta)
    r_total, r_comp, _ = _ref_loss(cfg, r_logits, tg, kind, weights, null_p, phase1, a, mats)
    r_total.backward()
    r_grads = {n: p.grad.clone() for n, p in model.named_parameters()}

    leaves = {n: t.clone().requires_grad_(True) for n, t in P.items()}
    o_logits = O.forward(leaves, a, x, meta)
    o_total, o_comp = O.hierarchical_loss(
        o_logits, tg, kind=kind, task_weights=weights, null_mask_prob=null_p, phase1_mask_null=phase1, soft_matrices=mats
    )
    o_total.backward()
    torch.testing.assert_close(o_total.detach(), r_total.detach(), rtol=1e-5, atol=1e-6)
    for t in o_comp["tasks"]:
        assert abs(o_comp["tasks"][t] - r_comp["tasks"][t]) < 1e-4
        assert abs(o_comp["weighted_tasks"][t] - r_comp["weighted_tasks"][t]) < 1e-4
    for n, g in r_grads.items():
        og = leaves[n].grad
        assert og is not None, n
        scale = g.abs().max().item() + 1e-12
        assert (og - g).abs().max().item() <= 1e-4 * scale + 1e-6, (n, (og - g).abs().max().item(), scale)


def test_adamw_clip_step_matches_reference():
    """R/train.py:282-313 + R/optimizers/build.py:67-106,687-716 for two steps.

    Gradients are compared in test_loss_and_grads_match_reference; Adam divides
    by sqrt(v), which turns rounding noise on analytically-zero gradients (e.g.
    the key bias) into O(lr) updates, so here the oracle's optimizer is fed the
    reference's own gradients and must then agree tightly."""
    cfg, nc, model, a = _build(**TINY)
    P = _load(model, a, seed=2)
    from linnaeus.optimizers.build import build_optimizer

    cfg.defrost()
    cfg.LR_SCHEDULER.BASE_LR = 3e-3
    opt = build_optimizer(cfg, model)
    state = {}
    Po = {n: t.clone() for n, t in P.items()}
    for step in (1, 2):
        x, meta, tg = O.synth_batch(a, 4, 10 + step)
        model.train()
        total, _, _ = _ref_loss(cfg, model(x, meta), tg)
        opt.zero_grad(set_to_none=True)
        total.backward()
        G = {n: p.grad.clone() for n, p in model.named_parameters()}
        ref_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
        opt.step()
        norm = O.adamw_clip_step(Po, G, state, step, 3e-3)
        assert abs(norm - float(ref_norm)) <= 1e-5 * float(ref_norm)
    for n, p in model.named_parameters():
        torch.testing.assert_close(Po[n], p.detach(), rtol=1e-5, atol=1e-6, msg=lambda m, n=n: f"{n}: {m}")


def test_train_step_runs_and_decreases_loss():
    cfg, nc, model, a = _build(**TINY)
    P = {n: t.clone() for n, t in _load(model, a, seed=4).items()}
    x, meta, tg = O.synth_batch(a, 4, 5)
    state = {}
    l0, _, _ = O.train_step(P, a, x, meta, tg, state, 1, 1e-3)
    l1, _, _ = O.train_step(P, a, x, meta, tg, state, 2, 1e-3)
    assert l1 < l0


# ----------------------------------------------------------------------------- class weights, validation, soft targets
def _ref_loss_full(cfg, logits, targets, kind, weights, class_weights=None, crit_weights=None, phase1=False, is_validation=False,
                   cw_train=True, cw_val=False):
    """The unmodified reference loss with class weights / criterion weights / validation flags set the way
    R/loss/utils.py:58-150 (prepare_loss_functions) and R/main.py wire them."""
    from linnaeus.loss.basic_loss import CrossEntropyLoss, LabelSmoothingCrossEntropy, SoftTargetCrossEntropy
    from linnaeus.loss.gradient_weighting import GradientWeighting
    from linnaeus.loss.hierarchical_loss import weighted_hierarchical_loss

    cfg = cfg.clone()
    cfg.defrost()
    cfg.TRAIN.PHASE1_MASK_NULL_LOSS = phase1
    cfg.LOSS.GRAD_WEIGHTING.CLASS.TRAIN = cw_train
    cfg.LOSS.GRAD_WEIGHTING.CLASS.VAL = cw_val
    ign = 0 if phase1 else None
    keys = list(logits.keys())
    cwt = crit_weights or {}
    if kind == "ce":
        crit = {k: CrossEntropyLoss(weight=cwt.get(k), apply_class_weights=k in cwt, ignore_index=ign) for k in keys}
    elif kind == "ls":
        crit = {k: LabelSmoothingCrossEntropy(weight=cwt.get(k), smoothing=0.1, apply_class_weights=k in cwt, ignore_index=ign) for k in keys}
    else:
        crit = {k: SoftTargetCrossEntropy(weight=cwt.get(k), apply_class_weights=k in cwt) for k in keys}
    cw_dicts = None
    if class_weights:
        cw_dicts = {k: {i: float(v[i]) for i in range(0, v.numel(), 2)} for k, v in class_weights.items()}  # sparse dict: missing -> 1.0
    gw = GradientWeighting(keys, cfg, "static", init_weights=weights, class_weights=cw_dicts)

    class Sched:
        def get_null_mask_prob(self, step):
            return 1.0

    return weighted_hierarchical_loss(logits, targets, crit, gw, Sched(), 0, is_validation=is_validation, config=cfg)


def _loss_case(seed=0, B=12, soft=False):
    g = torch.Generator().manual_seed(seed)
    tasks = [("taxa_L10", 11), ("taxa_L20", 7), ("taxa_L30", 4)]
    logits = {t: torch.randn(B, C, generator=g) for t, C in tasks}
    targets = {t: torch.randint(0, C, (B,), generator=g) for t, C in tasks}
    targets["taxa_L10"][:3] = 0
    targets["taxa_L20"][1] = 0
    if soft:  # mixup-style soft targets: lam * one-hot(a) + (1 - lam) * one-hot(b)
        for t, C in tasks:
            a = torch.nn.functional.one_hot(targets[t], C).float()
            b = a[torch.randperm(B, generator=g)]
            targets[t] = 0.7 * a + 0.3 * b
    cw = {t: 0.5 + torch.rand(C, generator=g) for t, C in tasks}
    for t in cw:  # the reference looks classes up in a sparse dict: odd indices default to 1.0
        cw[t][1::2] = 1.0
    weights = {t: 0.5 + 0.25 * i for i, (t, _) in enumerate(tasks)}
    return tasks, logits, targets, cw, weights


@pytest.mark.parametrize("kind", ["ce", "ls"])
@pytest.mark.parametrize("phase1,is_validation,cw_train,cw_val", [
    (False, False, True, False), (False, False, False, False), (True, False, True, False), (True, False, False, False),
    (False, True, True, False), (False, True, True, True), (True, True, True, False), (True, True, True, True)])
def test_class_weight_power_and_validation_match_reference(kind, phase1, is_validation, cw_train, cw_val):
    """a17: how often the class-weight lookup multiplies the per-sample loss (R/loss/masking.py:696-698,
    hierarchical_loss.py:313-334, gradient_weighting.py:334-352) in train / validation, with and without PHASE1, and the
    validation divisor when the criteria carry ignore_index = 0."""
    cfg, _, _, _ = _build(**TINY)
    tasks, logits, targets, cw, weights = _loss_case()
    r_total, r_comp, _ = _ref_loss_full(cfg, logits, targets, kind, weights, class_weights=cw, phase1=phase1, is_validation=is_validation,
                                        cw_train=cw_train, cw_val=cw_val)
    o_total, o_comp = O.hierarchical_loss(logits, targets, kind=kind, task_weights=weights, phase1_mask_null=phase1, is_validation=is_validation,
                                          class_weights=cw, apply_cw=(cw_val if is_validation else cw_train))
    torch.testing.assert_close(o_total, r_total, rtol=1e-5, atol=1e-6)
    for t in o_comp["weighted_tasks"]:
        assert abs(o_comp["weighted_tasks"][t] - r_comp["weighted_tasks"][t]) < 1e-5
    if "num_valid_samples_per_task" in r_comp["null_masking"]:
        assert {t: int(v) for t, v in r_comp["null_masking"]["num_valid_samples_per_task"].items()} == o_comp["num_valid_samples_per_task"]


@pytest.mark.parametrize("with_cw", [False, True])
def test_soft_target_cross_entropy_matches_reference(with_cw):
    """a15: SoftTargetCrossEntropy (basic_loss.py:188-228) on mixup-style [B, C] targets, class weights as sum_c t_c w_c."""
    cfg, _, _, _ = _build(**TINY)
    tasks, logits, targets, cw, weights = _loss_case(seed=1, soft=True)
    r_total, r_comp, _ = _ref_loss_full(cfg, logits, targets, "soft", weights, class_weights=cw if with_cw else None)
    o_total, o_comp = O.hierarchical_loss(logits, targets, kind="soft", task_weights=weights, class_weights=cw if with_cw else None, apply_cw=True)
    torch.testing.assert_close(o_total, r_total, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("kind", ["ce", "ls", "soft"])
def test_criterion_level_class_weights_match_reference(kind):
    """a15: criteria built with weight= and apply_class_weights=True (basic_loss.py:76-90,160-175,217-221)."""
    cfg, _, _, _ = _build(**TINY)
    tasks, logits, targets, cw, weights = _loss_case(seed=2, soft=(kind == "soft"))
    r_total, _, _ = _ref_loss_full(cfg, logits, targets, kind, weights, crit_weights=cw)
    o_total, _ = O.hierarchical_loss(logits, targets, kind=kind, task_weights=weights, criterion_weights=cw)
    torch.testing.assert_close(o_total, r_total, rtol=1e-5, atol=1e-6)
