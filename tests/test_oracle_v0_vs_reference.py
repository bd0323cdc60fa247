"""CPU: the mFormerV0 oracle (oracle/mformer_v0_oracle.py) equals the UNMODIFIED reference (eval mode) on the same
seeded weights and inputs, key for key and logit for logit.  Build container only (needs /root/reference)."""
import pytest
import torch

from tests.support import refload

pytestmark = pytest.mark.skipif(not refload.reference_available(), reason="reference tree not present")

TINY = dict(conv_embed=(16, 32), conv_out=(32, 64), conv_depths=(1, 2), conv_strides=((2,), (1, 1)), attn_dims=(64, 128), attn_depths=(2, 1),
            heads=(2, 4))


def _build(img, meta, **kw):
    refload.import_reference()
    from linnaeus.models import build_model

    from oracle import mformer_v0_oracle as V

    cfg, nc = refload.reference_config_v0(img_size=img, meta=meta, **kw)
    model = build_model(cfg, num_classes=nc, taxonomy_tree=None).eval()
    a = V.arch_from_config(cfg, nc)
    return model, a, V


@pytest.mark.parametrize("img,meta,kw,batch", [(64, True, TINY, 3), (96, True, TINY, 2), (224, True, {}, 1)])
def test_v0_oracle_matches_reference(img, meta, kw, batch):
    model, a, V = _build(img, meta, **kw)
    ref_sd = model.state_dict()
    shapes = V.param_shapes(a)
    assert sorted(ref_sd.keys()) == sorted(shapes.keys())
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    P = V.synth_state_dict(a, 0)
    for k in ref_sd:
        if k.endswith("relative_position_index"):
            assert torch.equal(ref_sd[k], P[k]), k
    model.load_state_dict(P)
    x, m = V.synth_batch(a, batch, 0)
    with torch.no_grad():
        ref = model(x, m)
        mine = V.forward(P, a, x, m)
    assert list(ref.keys()) == list(mine.keys())
    for k in ref:
        err = float((ref[k] - mine[k]).abs().max() / ref[k].abs().max())
        assert err < 1e-5, (k, err)
