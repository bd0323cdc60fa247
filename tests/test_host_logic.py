"""CPU: host-side logic of the drop-in boundary -- config schema, registry, parameter
names/shapes, decay split, heads wiring, flat buffers, bucket planning."""
import pytest
import torch

import linnaeus_b200 as L
from linnaeus_b200.config import CfgNode, get_default_config, make_synthetic_config
from oracle import mformer_oracle as O


def test_cfgnode_yacs_surface(tmp_path):
    c = get_default_config()
    assert c.MODEL.get("NOPE", 3) == 3
    c2 = c.clone()
    c2.MODEL.IMG_SIZE = 384
    assert c.MODEL.IMG_SIZE == 224
    c.freeze()
    with pytest.raises(AttributeError):
        c.MODEL.IMG_SIZE = 1
    c.defrost()
    c.merge_from_list(["MODEL.IMG_SIZE", "320", "TRAIN.CLIP_GRAD", 1.5])
    assert c.MODEL.IMG_SIZE == 320 and c.TRAIN.CLIP_GRAD == 1.5
    y = tmp_path / "x.yaml"
    y.write_text("MODEL:\n  TYPE: mFormerV1\n  ROPE_STAGES:\n    DEPTHS: [5, 2]\nDATA:\n  META:\n    COMPONENTS:\n      TEMPORAL: {ENABLED: true, DIM: 2, IDX: 0}\n")
    c.merge_from_file(str(y))
    assert c.MODEL.ROPE_STAGES.DEPTHS == [5, 2] and c.DATA.META.COMPONENTS.TEMPORAL.DIM == 2
    with pytest.raises(KeyError):
        c.merge_from_other_cfg(CfgNode({"NOT_A_SECTION": {"A": 1}}))
    assert isinstance(CfgNode.load_cfg("A: {B: 1}").A, CfgNode)


@pytest.mark.parametrize("variant,head_type", [("sm", "Linear"), ("md", "Linear")])
def test_state_dict_keys_and_shapes_match_reference_layout(variant, head_type):
    cfg, nc = make_synthetic_config(variant, head_type=head_type)
    m = L.build_model(cfg, nc)
    shapes = O.param_shapes(O.arch_from_config(cfg, nc))  # pinned to the real reference in test_oracle_vs_reference
    sd = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert list(sd.keys()) == list(shapes.keys())
    assert sd == shapes
    assert m.extra_token_num == 4 and m.use_meta and m.meta_dims == [2, 3, 10]
    assert list(m.head.keys()) == list(nc.keys())


def test_param_counts_match_survey():
    for variant, expect in (("sm", 31_897_483), ("md", 40_770_000)):
        cfg, nc = make_synthetic_config(variant)
        n = sum(p.numel() for p in L.build_model(cfg, nc).parameters())
        assert abs(n - expect) / expect < 2e-3, (variant, n)


def test_registry_and_errors():
    assert "mFormerV1" in L.list_models()
    cfg, nc = make_synthetic_config("sm")
    cfg.MODEL.TYPE = "nope"
    with pytest.raises(ValueError):
        L.build_model(cfg, nc)
    cfg, nc = make_synthetic_config("sm")
    cfg.MODEL.ROPE_STAGES.DEPTHS = [5]
    with pytest.raises(ValueError):
        L.build_model(cfg, nc)
    cfg, nc = make_synthetic_config("sm")
    del cfg.MODEL["CONVNEXT_STAGES"]
    with pytest.raises(ValueError):
        L.build_model(cfg, nc)


def test_cpu_forward_raises_instead_of_falling_back():
    cfg, nc = make_synthetic_config("sm", 64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(1, 1), conv_depths=(1, 1, 1, 1))
    m = L.build_model(cfg, nc)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 3, 64, 64), torch.randn(1, 15))


def test_decay_split_matches_reference_rule():
    from linnaeus_b200.optim import split_decay

    cfg, nc = make_synthetic_config("sm")
    m = L.build_model(cfg, nc)
    decay, no_decay = split_decay(m.named_parameters())
    assert (len(decay), len(no_decay)) == (86, 175)  # SURVEY 8(a20) [probe]
    names = {n for n, _ in decay}
    assert "cls_token_1" in names and "stages.2.0.attn.freqs" in names and "aggregate.weight" in names
    assert all(not n.endswith(".bias") for n in names)


class _Tree:
    def build_hierarchy_matrices(self):
        return {"taxa_L20_taxa_L10": torch.ones(400, 1000)}


@pytest.mark.parametrize("head_type,sub", [("HierarchicalSoftmax", "task_classifiers"), ("ConditionalClassifier", "level_classifiers")])
def test_hierarchical_heads_share_level_classifiers(head_type, sub):
    cfg, nc = make_synthetic_config("sm", n_tasks=2, head_type=head_type)
    m = L.build_model(cfg, nc, taxonomy_tree=_Tree())
    h0, h1 = m.head["taxa_L10"], m.head["taxa_L20"]
    assert getattr(h0, sub) is getattr(h1, sub)
    w, b = h1.classifier_params()
    assert w.shape == (400, 768)
    assert "head.taxa_L10.hmatrix_taxa_L20_taxa_L10" in m.state_dict()
    assert f"head.taxa_L20.{sub}.taxa_L10.weight" in m.state_dict()
    h0.set_gradnorm_mode(True)
    assert h0.is_gradnorm_mode()
    # shared parameters appear once in parameters()
    assert len({id(p) for p in m.parameters()}) == len(list(m.parameters()))


def test_flat_group_and_bucket_plan():
    from linnaeus_b200.flat import FlatGroup, is_flat
    from linnaeus_b200.parallel import plan_buckets

    ps = [torch.nn.Parameter(torch.randn(s)) for s in ((5, 3), (7,), (2, 2, 2), (1000,))]
    vals = [p.detach().clone() for p in ps]
    g = FlatGroup(ps, with_state=False)
    for i, (p, v) in enumerate(zip(ps, vals)):
        assert torch.equal(p.detach(), v) and is_flat(p, g, i)
        assert g.offsets[i] % 4 == 0
    ps[1].grad.add_(1.0)
    assert float(g.g.sum()) == 7.0
    buckets, owner = plan_buckets([g], bucket_bytes=64)
    assert sum(b.n_params for b in buckets) == 4
    assert buckets[0].lo == 0 and buckets[-1].hi == g.numel
    for a, b in zip(buckets[:-1], buckets[1:]):
        assert a.hi == b.lo
    # the first bucket (last to become final in backward: nothing left to overlap its all-reduce with) is cut small
    qs = [torch.nn.Parameter(torch.randn(64)) for _ in range(32)]
    g2 = FlatGroup(qs, with_state=False)
    b2, _ = plan_buckets([g2], bucket_bytes=64 * 4 * 8)
    sizes = [b.hi - b.lo for b in b2]
    assert sizes[0] == 64 * 2 and all(s == 64 * 8 for s in sizes[1:-1]) and sum(sizes) == g2.numel
    b3, _ = plan_buckets([g2], bucket_bytes=64 * 4 * 8, first_bucket_bytes=64 * 4 * 8)
    assert b3[0].hi - b3[0].lo == 64 * 8
