"""CPU: linnaeus_b200.gradnorm.GradNormModule against the UNMODIFIED reference module (R/loss/gradnorm.py) - initial weights for every
strategy, and the weight trajectory of several measure_and_update calls, bit for bit.  Skipped where /root/reference is absent;
the closed-form check below runs everywhere."""
import numpy as np
import pytest
import torch

from linnaeus_b200.gradnorm import GradNormModule, backbone_parameters
from tests.support import refload

KEYS = ["taxa_L10", "taxa_L20", "taxa_L30", "taxa_L40"]


@pytest.mark.skipif(not refload.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("alpha", [0.0, 0.5, 1.5])
@pytest.mark.parametrize("strategy", ["equal", "inverse_density", "class_complexity"])
def test_gradnorm_module_matches_reference(alpha, strategy):
    refload.import_reference()
    from linnaeus.loss.gradnorm import GradNormModule as Ref

    dens = None if strategy == "equal" else {"taxa_L10": 0.9, "taxa_L20": 0.5, "taxa_L30": 0.0002, "taxa_L40": 0.25}
    ncls = {"taxa_L10": 1000, "taxa_L20": 300, "taxa_L30": 40, "taxa_L40": 7}
    a = GradNormModule(KEYS, alpha=alpha, label_densities=dens, num_classes=ncls, init_strategy=strategy)
    b = Ref(KEYS, alpha=alpha, label_densities=dens, num_classes=ncls, init_strategy=strategy)
    assert torch.equal(a.task_weights, b.task_weights)
    rng = np.random.default_rng(int(alpha * 10) + len(strategy))
    for step in range(5):
        losses = {k: torch.tensor(float(rng.random() * 3 + 0.1)) for k in KEYS}
        grads = {k: torch.from_numpy(rng.standard_normal(257).astype(np.float32) * float(rng.random() * 4)) for k in KEYS}
        if step == 3:
            grads["taxa_L30"] = torch.zeros(257)  # a task without gradient: target stays, weight collapses like the reference's
        ma = a.measure_and_update(losses, grads)
        mb = b.measure_and_update(losses, grads)
        assert torch.equal(a.task_weights, b.task_weights), step
        assert torch.equal(a.initial_losses, b.initial_losses)
        assert set(ma) == set(mb)
        for k in mb:
            assert ma[k] == pytest.approx(mb[k], rel=1e-6, abs=1e-9)
        # the norms themselves (0-dim tensors) are an equivalent input
        c = GradNormModule(KEYS, alpha=alpha, init_weights=a.task_weights.clone())
        c.initial_losses.copy_(a.initial_losses)
        c.has_initted = True
        d = GradNormModule(KEYS, alpha=alpha, init_weights=a.task_weights.clone())
        d.initial_losses.copy_(a.initial_losses)
        d.has_initted = True
        c.measure_and_update(losses, grads, return_metrics=False)
        d.measure_and_update(losses, {k: v.norm(2) for k, v in grads.items()}, return_metrics=False)
        assert torch.equal(c.task_weights, d.task_weights)
    losses = {k: torch.tensor(1.0 + i) for i, k in enumerate(KEYS)}
    assert torch.equal(a(losses), b(losses))


def test_gradnorm_equalises_norms_in_closed_form():
    """alpha = 0: w_k <- w_k * norm_k / mean(norms), renormalised to sum K."""
    m = GradNormModule(KEYS, alpha=0.0)
    norms = torch.tensor([1.0, 2.0, 3.0, 6.0])
    m.measure_and_update({k: torch.tensor(1.0) for k in KEYS}, {k: norms[i] for i, k in enumerate(sorted(KEYS))}, return_metrics=False)
    w = norms / norms.mean()
    torch.testing.assert_close(m.task_weights, w * (4 / w.sum()))
    assert float(m.task_weights.sum()) == pytest.approx(4.0)


def test_backbone_parameters_default_filter():
    net = torch.nn.Module()
    net.stem = torch.nn.Linear(3, 3)
    net.head = torch.nn.ModuleDict({"taxa_L10": torch.nn.Linear(3, 2)})
    net.meta_temporal_head_1 = torch.nn.Linear(2, 3)
    net.frozen = torch.nn.Linear(3, 3)
    for p in net.frozen.parameters():
        p.requires_grad_(False)
    names = {id(p) for p in backbone_parameters(net)}
    assert names == {id(net.stem.weight), id(net.stem.bias)}


@pytest.mark.skipif(not refload.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("kind", ["static", "gradnorm"])
def test_gradient_weighting_forward_matches_reference(kind):
    """GradientWeighting.forward: per-task mean over num_valid, class weights (hard and one-hot targets, labels without an entry),
    task weights from the static vector or the GradNorm buffer - against the unmodified reference class on CPU."""
    refload.import_reference()
    from linnaeus.loss.gradient_weighting import GradientWeighting as Ref

    from linnaeus_b200.gradnorm import GradientWeighting

    rcfg, _ = refload.reference_config("sm", 64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(1, 1), conv_depths=(1, 1, 1, 1))
    keys = KEYS[:3]
    cw = {"taxa_L10": {0: 0.25, 2: 3.0}, "taxa_L30": {1: 0.5, 3: 2.0, 4: 1.5}}
    init = {"taxa_L10": 0.7, "taxa_L20": 1.1, "taxa_L30": 1.2}
    a = GradientWeighting(keys, rcfg, kind, init_weights=init, class_weights=cw, alpha=1.0)
    b = Ref(keys, rcfg, kind, init_weights=init, class_weights=cw, alpha=1.0)
    assert torch.equal(a.task_weights, b.task_weights) and a.update_interval == b.update_interval and a.zero_aux_info == b.zero_aux_info
    if kind == "gradnorm":
        assert torch.equal(a.gradnorm.task_weights, b.gradnorm.task_weights)
        upd = ({k: torch.tensor(1.0 + i) for i, k in enumerate(keys)}, {k: torch.tensor(2.0 - 0.5 * i) for i, k in enumerate(keys)})
        a.gradnorm.measure_and_update(*upd, return_metrics=False)
        b.gradnorm.measure_and_update(*upd)
    rng = np.random.default_rng(0)
    B = 13
    losses = {k: torch.from_numpy(rng.random(B).astype(np.float32)) for k in keys}
    targets = {"taxa_L10": torch.from_numpy(rng.integers(0, 6, size=B)), "taxa_L20": torch.from_numpy(rng.integers(0, 4, size=B)),
               "taxa_L30": torch.from_numpy(np.eye(5, dtype=np.float32)[rng.integers(0, 5, size=B)])}
    for nv in (None, {"taxa_L10": 9, "taxa_L30": 0}):
        wa, da = a(losses, targets, num_valid_samples_per_task=nv)
        wb, db = b(losses, targets, num_valid_samples_per_task=nv)
        assert da == db
        for k in keys:
            torch.testing.assert_close(wa[k], wb[k], rtol=1e-6, atol=0)
    # backbone selection on a model of this package
    import linnaeus_b200 as L

    cfg, nc = L.make_synthetic_config("sm", 64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(1, 1), conv_depths=(1, 1, 1, 1))
    model = L.build_model(cfg, nc)
    a.set_model(model)
    if kind == "gradnorm":
        names = {id(p): n for n, p in model.named_parameters()}
        picked = [names[id(p)] for p in a.backbone_params]
        assert picked and not any("head" in n or "meta_" in n for n in picked)
        assert len(picked) == sum(1 for n, _ in model.named_parameters() if "head" not in n and "meta_" not in n)
    else:
        assert a.backbone_params is None and a.update_gradnorm_weights_reforward(None, {}) == {}


def test_task_weight_vector_follows_the_live_gradnorm_buffer_and_caches_static_weights():
    from linnaeus_b200.gradnorm import GradientWeighting
    from linnaeus_b200.loss import StaticTaskWeighting, _task_weight_vector

    keys = KEYS[:3]
    dev = torch.device("cpu")
    st = StaticTaskWeighting(keys, {"taxa_L10": 0.5, "taxa_L20": 2.0, "taxa_L30": 3.0})
    v = _task_weight_vector(st, keys, dev)
    assert v.tolist() == [0.5, 2.0, 3.0] and _task_weight_vector(st, keys, dev) is v  # uploaded once
    assert _task_weight_vector(st, keys[::-1], dev).tolist() == [3.0, 2.0, 0.5]
    gw = GradientWeighting(keys, None, "gradnorm", init_weights=[1.0, 2.0, 3.0])
    assert _task_weight_vector(gw, keys, dev).tolist() == [1.0, 2.0, 3.0]
    gw.gradnorm.task_weights.copy_(torch.tensor([0.25, 0.5, 2.25]))  # an update must be seen by the next loss call
    assert _task_weight_vector(gw, keys, dev).tolist() == [0.25, 0.5, 2.25]
    assert _task_weight_vector(gw, [keys[2], keys[0], keys[1]], dev).tolist() == [2.25, 0.25, 0.5]
    assert _task_weight_vector(object(), keys, dev) is None
