"""CPU: checkpoint plumbing (file names, keys, module. prefix, auto-resume) against the reference's conventions
(R/utils/checkpoint.py:21-42, 1028-1131, 1307-1332)."""
import os
import time

import pytest
import torch

from linnaeus_b200 import checkpoint as C
from tests.support import refload


def test_prefix_cleaning_matches_reference():
    sd = {"a.weight": 1, "b.bias": 2}
    pre = {"module.a.weight": 1, "module.b.bias": 2}
    assert C.clean_state_dict_keys(sd, True, False) == pre
    assert C.clean_state_dict_keys(pre, False, True) == sd
    assert C.clean_state_dict_keys(sd, False, False) == sd and C.clean_state_dict_keys(pre, True, True) == pre
    if refload.reference_available():
        refload.import_reference()
        from linnaeus.utils.checkpoint import _clean_state_dict_keys

        for d, ddp, has in ((sd, True, False), (pre, False, True), (sd, False, False), (pre, True, True), ({"module.x": 1, "y": 2}, False, True)):
            assert C.clean_state_dict_keys(d, ddp, has) == _clean_state_dict_keys(d, ddp, has)


def test_save_load_roundtrip_and_auto_resume(tmp_path):
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.LayerNorm(3))
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    net(torch.randn(2, 4)).sum().backward()
    opt.step()
    assert C.auto_resume_helper(str(tmp_path)) is None
    p1 = C.save_checkpoint(str(tmp_path), net, opt, epoch=1, config={"k": 1}, iteration=7)
    assert os.path.basename(p1) == "ckpt_epoch_1.pth" and os.path.exists(tmp_path / "latest.pth")
    ck = torch.load(p1, weights_only=False)
    assert set(ck) == {"model", "optimizer", "lr_scheduler", "epoch", "config", "iteration"}  # the reference's core keys
    time.sleep(0.02)
    p2 = C.save_checkpoint(str(tmp_path), net, opt, epoch=2, extra={"wandb_run_id": "r"})
    os.utime(p2, (time.time() + 5, time.time() + 5))
    assert C.auto_resume_helper(str(tmp_path)) == p2
    # a DDP-prefixed file loads into a bare model
    ck["model"] = {f"module.{k}": v for k, v in ck["model"].items()}
    torch.save(ck, tmp_path / "ddp.pth")
    net2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.LayerNorm(3))
    opt2 = torch.optim.AdamW(net2.parameters(), lr=5e-2)
    rest = C.load_checkpoint(str(tmp_path / "ddp.pth"), net2, opt2)
    assert rest["epoch"] == 1 and rest["iteration"] == 7 and rest["config"] == {"k": 1}
    for a, b in zip(net.state_dict().values(), net2.state_dict().values()):
        assert torch.equal(a, b)
    assert opt2.param_groups[0]["lr"] == 1e-3
    with pytest.raises(RuntimeError):
        C.load_checkpoint(p1, torch.nn.Linear(4, 3))


@pytest.mark.skipif(not refload.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("arch", ["v1", "v0"])
def test_checkpoints_interchange_with_the_reference_model(arch, tmp_path):
    """A checkpoint of the REFERENCE model (DDP-prefixed, reference dict layout) loads strictly into the B200 model and a
    checkpoint written here loads strictly into the reference model: same keys, shapes and values both ways."""
    import linnaeus_b200 as L

    refload.import_reference()
    from linnaeus.models import build_model as ref_build

    if arch == "v1":
        rcfg, nc = refload.reference_config("sm", 64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(1, 1), conv_depths=(1, 1, 1, 1))
        cfg, _ = L.make_synthetic_config("sm", 64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(1, 1), conv_depths=(1, 1, 1, 1))
    else:
        kw = dict(conv_embed=(16, 32), conv_out=(32, 64), conv_depths=(1, 2), conv_strides=((2,), (1, 1)), attn_dims=(64, 128), attn_depths=(2, 1),
                  heads=(2, 4))
        rcfg, nc = refload.reference_config_v0(img_size=64, **kw)
        cfg, _ = L.make_synthetic_config_v0("sm", 64, **kw)
    torch.manual_seed(1)
    ref = ref_build(rcfg, num_classes=nc, taxonomy_tree=None)
    ours = L.build_model(cfg, nc)
    # reference -> here (as DDP would have saved it)
    torch.save({"model": {f"module.{k}": v for k, v in ref.state_dict().items()}, "optimizer": None, "lr_scheduler": None, "epoch": 3,
                "config": None, "iteration": 11}, tmp_path / "ref.pth")
    rest = C.load_checkpoint(str(tmp_path / "ref.pth"), ours, strict=True)
    assert rest["epoch"] == 3 and rest["iteration"] == 11
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
    # here -> reference
    with torch.no_grad():
        for p in ours.parameters():
            p.add_(0.25)
    path = C.save_checkpoint(str(tmp_path), ours, None, epoch=4)
    ck = torch.load(path, weights_only=False)
    missing, unexpected = ref.load_state_dict(ck["model"], strict=True)
    assert not missing and not unexpected
    for k, v in ours.state_dict().items():
        assert torch.equal(ref.state_dict()[k], v), k


def test_roundtrip_through_the_data_parallel_wrapper(tmp_path):
    """linnaeus_b200.DataParallel exposes ``module.``-prefixed keys like torch's DDP: a checkpoint saved through the wrapper loads
    back into the wrapper AND into the bare model, and a bare checkpoint loads into the wrapper (the reference
    decides from the target model's own keys, R/utils/checkpoint.py:798-830)."""
    from linnaeus_b200.flat import FlatGroup
    from linnaeus_b200.parallel import DataParallel

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.LayerNorm(3))
    dp = DataParallel(net, [FlatGroup(list(net.parameters()), with_state=False)])
    assert all(k.startswith("module.") for k in dp.state_dict())
    p_wrapped = C.save_checkpoint(str(tmp_path / "w"), dp, None, epoch=1)
    p_bare = C.save_checkpoint(str(tmp_path / "b"), net, None, epoch=1)
    want = {k: v.clone() for k, v in net.state_dict().items()}
    for path in (p_wrapped, p_bare):
        for target_wrapped in (True, False):
            net2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.LayerNorm(3))
            tgt = DataParallel(net2, [FlatGroup(list(net2.parameters()), with_state=False)]) if target_wrapped else net2
            C.load_checkpoint(path, tgt, strict=True)
            for k, v in net2.state_dict().items():
                assert torch.equal(v, want[k]), (path, target_wrapped, k)
