"""Test support: import the read-only reference tree (build container only).

`/root/reference` does not exist on the GPU box; every user of this module must
skip when :func:`reference_available` is False.  Nothing under ``linnaeus_b200``
imports this file.
"""
from __future__ import annotations

import logging
import os
import sys

REF_ROOT = os.environ.get("LINNAEUS_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")
_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "linnaeus", "models"))


def import_reference():
    """Put the shims, this repo and the reference on sys.path and import it."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    for p in (REF_ROOT, _REPO, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    logging.getLogger("linnaeus").setLevel(logging.ERROR)
    import linnaeus  # noqa: F401
    logging.getLogger("linnaeus").setLevel(logging.ERROR)
    return linnaeus


def reference_config(variant: str = "sm", img_size: int = 224, n_tasks: int = 6, meta: bool = True,
                     head_type: str = "Linear", dims=None, rope_depths=None, heads=None, conv_depths=None):
    """The reference's own default config + arch YAML + the synthetic heads/meta
    of SURVEY.md 8(d), with the deterministic oracle settings of 8(c)."""
    import_reference()
    from linnaeus.config import get_default_config
    from yacs.config import CfgNode as CN
    import yaml

    cfg = get_default_config()
    cfg.defrost()
    with open(os.path.join(REF_ROOT, f"configs/model/archs/mFormerV1/mFormerV1_{variant}.yaml")) as f:
        arch = yaml.safe_load(f)["MODEL"]
    for k, v in arch.items():
        cfg.MODEL[k] = CN(v, new_allowed=True) if isinstance(v, dict) else v
    if dims is not None:
        cfg.MODEL.CONVNEXT_STAGES.DIMS = list(dims)
        cfg.MODEL.ROPE_STAGES.DIMS = [dims[2], dims[3]]
    if conv_depths is not None:
        cfg.MODEL.CONVNEXT_STAGES.DEPTHS = list(conv_depths)
    if rope_depths is not None:
        cfg.MODEL.ROPE_STAGES.DEPTHS = list(rope_depths)
    if heads is not None:
        cfg.MODEL.ROPE_STAGES.NUM_HEADS = list(heads)
    cfg.MODEL.IMG_SIZE = img_size
    cfg.MODEL.USE_FLASH_ATTN = False
    cfg.MODEL.DROP_PATH_RATE = 0.0
    cfg.MODEL.DROP_RATE = 0.0
    cfg.MODEL.ATTN_DROP_RATE = 0.0
    cfg.MODEL.PRETRAINED = None
    from linnaeus_b200.config import SYNTH_TASKS, SYNTH_CLASSES, SYNTH_META
    tasks = SYNTH_TASKS[:n_tasks]
    cfg.DATA.TASK_KEYS_H5 = list(tasks)
    cfg.MODEL.CLASSIFICATION.HEADS = CN(new_allowed=True)
    for t in tasks:
        cfg.MODEL.CLASSIFICATION.HEADS[t] = CN({"TYPE": head_type}, new_allowed=True)
    cfg.DATA.META.ACTIVE = bool(meta)
    cfg.DATA.META.COMPONENTS = CN(new_allowed=True)
    if meta:
        for name, dim, idx in SYNTH_META:
            cfg.DATA.META.COMPONENTS[name] = CN({"ENABLED": True, "DIM": dim, "IDX": idx}, new_allowed=True)
    cfg.TRAIN.AMP_OPT_LEVEL = "O0"
    cfg.TRAIN.GRADIENT_CHECKPOINTING.ENABLED_NORMAL_STEPS = False
    cfg.LOSS.GRAD_WEIGHTING.TASK.TYPE = "static"
    cfg.LOSS.GRAD_WEIGHTING.TASK.GRADNORM_ENABLED = False
    cfg.LOSS.GRAD_WEIGHTING.CLASS.METHOD = "none"
    num_classes = dict(zip(tasks, SYNTH_CLASSES[:n_tasks]))
    return cfg, num_classes


def synthetic_taxonomy_tree(num_classes: dict):
    """TaxonomyTree over the synthetic hierarchy (child i -> parent
    0 if i == 0 else 1 + (i-1) mod (C_parent-1)); needed for hierarchical heads."""
    import_reference()
    from linnaeus.utils.taxonomy.taxonomy_tree import TaxonomyTree

    tasks = list(num_classes.keys())
    hmap = {}
    for lo, hi in zip(tasks[:-1], tasks[1:]):
        cp = num_classes[hi]
        hmap[lo] = {i: (0 if i == 0 else 1 + (i - 1) % (cp - 1)) for i in range(num_classes[lo])}
    return TaxonomyTree(hmap, tasks, num_classes)


def reference_config_v0(img_size: int = 224, n_tasks: int = 6, meta: bool = True, **arch_kw):
    """Reference default config + the mFormerV0 arch section produced by linnaeus_b200.config.make_synthetic_config_v0
    (same values as configs/model/archs/mFormerV0/mFormerV0_sm.yaml), with the deterministic eval settings."""
    import_reference()
    from linnaeus.config import get_default_config
    from yacs.config import CfgNode as CN
    from linnaeus_b200.config import SYNTH_CLASSES, SYNTH_META, SYNTH_TASKS, make_synthetic_config_v0

    mine, _ = make_synthetic_config_v0("sm", img_size, n_tasks, meta, **arch_kw)
    cfg = get_default_config()
    cfg.defrost()
    cfg.MODEL.TYPE = "mFormerV0"
    cfg.MODEL.IMG_SIZE = img_size
    cfg.MODEL.DROP_PATH_RATE = 0.0
    cfg.MODEL.DROP_RATE = 0.0
    cfg.MODEL.ATTN_DROP_RATE = 0.0
    cfg.MODEL.PRETRAINED = None
    cfg.MODEL.ONLY_LAST_CLS = False
    cfg.MODEL.CONV_STAGES = CN({k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in mine.MODEL.CONV_STAGES.items()}, new_allowed=True)
    cfg.MODEL.ATTENTION_STAGES = CN({k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in mine.MODEL.ATTENTION_STAGES.items()},
                                    new_allowed=True)
    tasks = SYNTH_TASKS[:n_tasks]
    cfg.DATA.TASK_KEYS_H5 = list(tasks)
    cfg.MODEL.CLASSIFICATION.HEADS = CN(new_allowed=True)
    for t in tasks:
        cfg.MODEL.CLASSIFICATION.HEADS[t] = CN({"TYPE": "Linear"}, new_allowed=True)
    cfg.DATA.META.ACTIVE = bool(meta)
    cfg.DATA.META.COMPONENTS = CN(new_allowed=True)
    if meta:
        for name, dim, idx in SYNTH_META:
            cfg.DATA.META.COMPONENTS[name] = CN({"ENABLED": True, "DIM": dim, "IDX": idx}, new_allowed=True)
    cfg.TRAIN.GRADIENT_CHECKPOINTING.ENABLED_NORMAL_STEPS = False
    return cfg, dict(zip(tasks, SYNTH_CLASSES[:n_tasks]))
