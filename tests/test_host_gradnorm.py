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
