"""CPU (reference tree present): the drop-in boundary of SURVEY 8(b).  After ``install_into_linnaeus()`` the REFERENCE's own
``build_model`` (R/models/build.py:52-111 -> model_factory.create_model, :179-213) returns the B200 class, and that model loads a
state_dict produced by the reference model key for key; ``install_into_linnaeus_loss`` rebinds the loss at the reference's call
sites (train.py:7 / validation.py import ``weighted_hierarchical_loss`` by name)."""
import sys
import types

import pytest
import torch

from tests.support import refload

pytestmark = pytest.mark.skipif(not refload.reference_available(), reason="reference tree not present")


def test_reference_build_model_returns_the_b200_class_and_loads_a_reference_state_dict():
    import linnaeus_b200 as L
    from linnaeus_b200.mformer_v1 import mFormerV1 as Ours
    from linnaeus_b200.registry import install_into_linnaeus

    refload.import_reference()
    from linnaeus.models import build_model as ref_build
    from linnaeus.models import model_factory as mf

    kw = dict(dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(1, 1), conv_depths=(1, 1, 1, 1))
    rcfg, nc = refload.reference_config("sm", 64, **kw)
    torch.manual_seed(0)
    ref_model = ref_build(rcfg, num_classes=nc, taxonomy_tree=None)
    original = dict(mf._model_registry)
    try:
        replaced = install_into_linnaeus()
        assert "mFormerV1" in replaced and "mFormerV0" in replaced
        ours = ref_build(rcfg, num_classes=nc, taxonomy_tree=None)  # the reference's entry point, the reference's config object
        assert isinstance(ours, Ours)
        missing, unexpected = ours.load_state_dict(ref_model.state_dict(), strict=True)
        assert not missing and not unexpected
        for (n1, p1), (n2, p2) in zip(ref_model.named_parameters(), ours.named_parameters()):
            assert n1 == n2 and torch.equal(p1, p2)
        # the attributes callers touch
        assert set(ours.head.keys()) == set(nc) and ours.extra_token_num == 4 and ours.use_meta
        assert "stages" in ours.parameter_groups_metadata and "drop_params" in ours.pretrained_ckpt_handling_metadata
        with pytest.raises(RuntimeError):  # no CPU fallback behind the reference's surface either
            ours(torch.randn(1, 3, 64, 64), torch.randn(1, 15))
    finally:
        mf._model_registry.clear()
        mf._model_registry.update(original)


def test_install_into_linnaeus_loss_rebinds_the_reference_call_sites(monkeypatch):
    """linnaeus.train / linnaeus.validation pull in h5py and friends that are absent offline, so the modules are stubbed with what
    they bind at import time (``from linnaeus.loss.hierarchical_loss import weighted_hierarchical_loss``)."""
    import linnaeus_b200.loss as LL

    refload.import_reference()
    from linnaeus.loss.hierarchical_loss import weighted_hierarchical_loss as ref_loss

    for name in ("linnaeus.train", "linnaeus.validation", "linnaeus.utils.autobatch"):
        m = types.ModuleType(name)
        m.weighted_hierarchical_loss = ref_loss
        monkeypatch.setitem(sys.modules, name, m)
    import linnaeus
    import linnaeus.utils

    monkeypatch.setattr(linnaeus, "train", sys.modules["linnaeus.train"], raising=False)
    monkeypatch.setattr(linnaeus, "validation", sys.modules["linnaeus.validation"], raising=False)
    monkeypatch.setattr(linnaeus.utils, "autobatch", sys.modules["linnaeus.utils.autobatch"], raising=False)
    LL.install_into_linnaeus_loss()
    for name in ("linnaeus.train", "linnaeus.validation", "linnaeus.utils.autobatch"):
        assert sys.modules[name].weighted_hierarchical_loss is LL.weighted_hierarchical_loss
