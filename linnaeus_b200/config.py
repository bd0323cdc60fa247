"""Minimal yacs-compatible ``CfgNode`` and the hot-path default config tree.

`yacs` is not installable in the build image, and the reference only uses a
small part of its API (attribute access, ``get``, ``clone``, ``defrost`` /
``freeze``, ``merge_from_file`` / ``merge_from_other_cfg`` / ``merge_from_list``,
``load_cfg``, ``new_allowed``).  This class covers that surface so that

* the YAML schema of the reference's ``configs/**`` loads unchanged, and
* a real ``yacs.config.CfgNode`` passed in by ``linnaeus/main.py`` works too,
  because the model only ever uses ``.get`` / attribute access / ``hasattr``.

Reference behaviour followed: ``linnaeus/config.py:69-982`` (default tree; only
the keys the hot path reads are reproduced in :func:`get_default_config`),
``linnaeus/utils/config_utils.py:107-168`` (``MODEL.BASE`` arch YAML merge).
"""

from __future__ import annotations

import copy
from ast import literal_eval
from typing import Any

import yaml

__all__ = ["CfgNode", "CN", "get_default_config", "load_arch_yaml", "make_synthetic_config"]


class CfgNode(dict):
    IMMUTABLE = "__immutable__"
    NEW_ALLOWED = "__new_allowed__"

    def __init__(self, init_dict: dict | None = None, key_list=None, new_allowed: bool = False):
        init_dict = {} if init_dict is None else init_dict
        super().__init__()
        self.__dict__[CfgNode.IMMUTABLE] = False
        self.__dict__[CfgNode.NEW_ALLOWED] = new_allowed
        for k, v in init_dict.items():
            if isinstance(v, dict) and not isinstance(v, CfgNode):
                v = CfgNode(v, new_allowed=new_allowed)
            dict.__setitem__(self, k, v)

    # attribute access -------------------------------------------------
    def __getattr__(self, name: str) -> Any:
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name: str, value: Any) -> None:
        if self.is_frozen():
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        if name in self.__dict__:
            raise AttributeError(f"Invalid attempt to modify internal CfgNode state: {name}")
        if isinstance(value, dict) and not isinstance(value, CfgNode):
            value = CfgNode(value, new_allowed=True)
        self[name] = value

    def __str__(self) -> str:
        return yaml.safe_dump(self.to_dict(), default_flow_style=None)

    def __repr__(self) -> str:
        return f"CfgNode({dict.__repr__(self)})"

    # yacs API ---------------------------------------------------------
    def to_dict(self) -> dict:
        out = {}
        for k, v in self.items():
            out[k] = v.to_dict() if isinstance(v, CfgNode) else (list(v) if isinstance(v, tuple) else v)
        return out

    def dump(self, **kwargs) -> str:
        return yaml.safe_dump(self.to_dict(), **kwargs)

    def is_frozen(self) -> bool:
        return self.__dict__[CfgNode.IMMUTABLE]

    def is_new_allowed(self) -> bool:
        return self.__dict__[CfgNode.NEW_ALLOWED]

    def set_new_allowed(self, is_new_allowed: bool) -> None:
        self.__dict__[CfgNode.NEW_ALLOWED] = is_new_allowed
        for v in self.values():
            if isinstance(v, CfgNode):
                v.set_new_allowed(is_new_allowed)

    def _immutable(self, flag: bool) -> None:
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._immutable(flag)

    def freeze(self) -> None:
        self._immutable(True)

    def defrost(self) -> None:
        self._immutable(False)

    def clone(self) -> "CfgNode":
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        new = CfgNode(new_allowed=self.is_new_allowed())
        for k, v in self.items():
            dict.__setitem__(new, k, copy.deepcopy(v, memo))
        new.__dict__[CfgNode.IMMUTABLE] = self.is_frozen()
        return new

    @classmethod
    def load_cfg(cls, cfg_file_obj_or_str) -> "CfgNode":
        if hasattr(cfg_file_obj_or_str, "read"):
            cfg_file_obj_or_str = cfg_file_obj_or_str.read()
        data = yaml.safe_load(cfg_file_obj_or_str) or {}
        return cls(data, new_allowed=True)

    def merge_from_file(self, cfg_filename: str) -> None:
        with open(cfg_filename) as f:
            other = self.load_cfg(f)
        self.merge_from_other_cfg(other)

    def merge_from_other_cfg(self, other: "CfgNode") -> None:
        _merge_a_into_b(other, self, self, [])

    def merge_from_list(self, cfg_list: list) -> None:
        if len(cfg_list) % 2 != 0:
            raise ValueError(f"Override list has odd length: {cfg_list}")
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            d = self
            parts = full_key.split(".")
            for sub in parts[:-1]:
                if sub not in d:
                    if d.is_new_allowed():
                        dict.__setitem__(d, sub, CfgNode(new_allowed=True))
                    else:
                        raise KeyError(f"Non-existent key: {full_key}")
                d = d[sub]
            last = parts[-1]
            if last not in d and not d.is_new_allowed():
                raise KeyError(f"Non-existent key: {full_key}")
            dict.__setitem__(d, last, _decode_value(v))


CN = CfgNode


def _decode_value(v):
    if isinstance(v, dict) and not isinstance(v, CfgNode):
        return CfgNode(v, new_allowed=True)
    if not isinstance(v, str):
        return v
    try:
        return literal_eval(v)
    except (ValueError, SyntaxError):
        return v


def _merge_a_into_b(a: CfgNode, b: CfgNode, root: CfgNode, key_list: list) -> None:
    for k, v_ in a.items():
        full_key = ".".join(key_list + [k])
        v = copy.deepcopy(v_)
        v = _decode_value(v)
        if k in b:
            if isinstance(v, CfgNode) and isinstance(b[k], CfgNode):
                _merge_a_into_b(v, b[k], root, key_list + [k])
            else:
                # yacs coerces tuple<->list; everything else is replaced as given
                if isinstance(b[k], tuple) and isinstance(v, list):
                    v = tuple(v)
                elif isinstance(b[k], list) and isinstance(v, tuple):
                    v = list(v)
                dict.__setitem__(b, k, v)
        elif b.is_new_allowed():
            dict.__setitem__(b, k, v)
        else:
            raise KeyError(f"Non-existent config key: {full_key}")


# ---------------------------------------------------------------------------
# Default tree: only the keys the hot path reads, with the
# reference's default values (linnaeus/config.py).
# ---------------------------------------------------------------------------
def get_default_config() -> CfgNode:
    c = CfgNode()
    c.DATA = CfgNode()
    c.DATA.IMG_SIZE = 224
    c.DATA.TASK_KEYS_H5 = ["taxa_L10", "taxa_L20", "taxa_L30", "taxa_L40"]
    c.DATA.META = CfgNode(new_allowed=True)
    c.DATA.META.ACTIVE = True
    c.DATA.META.COMPONENTS = CfgNode(new_allowed=True)

    c.MODEL = CfgNode(new_allowed=True)
    c.MODEL.TYPE = "mFormerV1"
    c.MODEL.NAME = "mFormerV1_sm"
    c.MODEL.BASE = []
    c.MODEL.IMG_SIZE = 224
    c.MODEL.IN_CHANS = 3
    c.MODEL.DROP_RATE = 0.0
    c.MODEL.DROP_PATH_RATE = 0.1
    c.MODEL.ATTN_DROP_RATE = 0.0
    c.MODEL.LABEL_SMOOTHING = 0.1
    c.MODEL.ONLY_LAST_CLS = False
    c.MODEL.EXTRA_TOKEN_NUM = 1
    c.MODEL.USE_FLASH_ATTN = False
    c.MODEL.META_DIMS = []
    c.MODEL.PRETRAINED = None
    c.MODEL.PRETRAINED_SOURCE = None
    c.MODEL.FIND_UNUSED_PARAMETERS = False
    c.MODEL.CLASSIFICATION = CfgNode()
    c.MODEL.CLASSIFICATION.HEADS = CfgNode(new_allowed=True)

    c.TRAIN = CfgNode(new_allowed=True)
    c.TRAIN.AMP_OPT_LEVEL = "O1"
    c.TRAIN.CLIP_GRAD = 5.0
    c.TRAIN.ACCUMULATION_STEPS = 1
    c.TRAIN.PHASE1_MASK_NULL_LOSS = False
    c.TRAIN.GRADIENT_CHECKPOINTING = CfgNode()
    c.TRAIN.GRADIENT_CHECKPOINTING.ENABLED_NORMAL_STEPS = True
    c.TRAIN.GRADIENT_CHECKPOINTING.ENABLED_GRADNORM_STEPS = True

    c.LOSS = CfgNode(new_allowed=True)
    c.LOSS.TASK_SPECIFIC = CfgNode(new_allowed=True)
    c.LOSS.GRAD_WEIGHTING = CfgNode(new_allowed=True)
    c.LOSS.GRAD_WEIGHTING.TASK = CfgNode(new_allowed=True)
    c.LOSS.GRAD_WEIGHTING.TASK.TYPE = "static"
    c.LOSS.GRAD_WEIGHTING.TASK.INIT_WEIGHTS = None
    c.LOSS.GRAD_WEIGHTING.CLASS = CfgNode(new_allowed=True)
    c.LOSS.GRAD_WEIGHTING.CLASS.METHOD = "none"
    c.LOSS.GRAD_WEIGHTING.CLASS.TRAIN = True
    c.LOSS.GRAD_WEIGHTING.CLASS.VAL = False

    c.OPTIMIZER = CfgNode(new_allowed=True)
    c.OPTIMIZER.NAME = "adamw"
    c.OPTIMIZER.EPS = 1e-8
    c.OPTIMIZER.BETAS = (0.9, 0.999)
    c.OPTIMIZER.WEIGHT_DECAY = 0.05
    c.LR_SCHEDULER = CfgNode(new_allowed=True)
    c.LR_SCHEDULER.BASE_LR = 1e-4
    c.LR_SCHEDULER.REFERENCE_BS = 512

    c.DEBUG = CfgNode(new_allowed=True)
    c.DEBUG.LOSS = CfgNode(new_allowed=True)
    c.DEBUG.LOSS.NULL_MASKING = False
    return c


# ---------------------------------------------------------------------------
# Architecture presets.  Values are those of the reference's arch YAMLs
# (configs/model/archs/mFormerV1/mFormerV1_{sm,md,lg,xl}.yaml), restated here
# so benchmarks and tests do not need the reference tree at run time.
# ---------------------------------------------------------------------------
_ARCH_V1 = {
    #        convnext depths      dims                     rope depths  heads      drop_path
    "sm": ([3, 3, 9, 3], [96, 192, 384, 768], [5, 2], [6, 12], 0.2),
    "md": ([3, 3, 27, 3], [96, 192, 384, 768], [10, 2], [6, 12], 0.3),
    "lg": ([3, 3, 27, 3], [192, 384, 768, 1536], [10, 2], [12, 24], 0.4),
    "xl": ([3, 3, 27, 3], [256, 512, 1024, 2048], [22, 2], [16, 32], 0.5),
}

SYNTH_TASKS = ["taxa_L10", "taxa_L20", "taxa_L30", "taxa_L40", "taxa_L50", "taxa_L60"]
SYNTH_CLASSES = [1000, 400, 120, 40, 12, 4]
SYNTH_META = (("TEMPORAL", 2, 0), ("SPATIAL", 3, 1), ("ELEVATION", 10, 2))


def load_arch_yaml(cfg: CfgNode, path: str) -> CfgNode:
    """Merge an arch YAML's ``MODEL`` section into ``cfg`` (wholesale, like the
    reference's ``load_model_base_config``, config_utils.py:107-168)."""
    with open(path) as f:
        data = yaml.safe_load(f)
    model = data.get("MODEL", data)
    was_frozen = cfg.is_frozen()
    cfg.defrost()
    for k, v in model.items():
        cfg.MODEL[k] = CfgNode(v, new_allowed=True) if isinstance(v, dict) else v
    if was_frozen:
        cfg.freeze()
    return cfg


def make_synthetic_config(
    variant: str = "sm",
    img_size: int = 224,
    n_tasks: int = 6,
    meta: bool = True,
    head_type: str = "Linear",
    drop_path: float = 0.0,
    dims=None,
    heads=None,
    rope_depths=None,
    conv_depths=None,
) -> tuple[CfgNode, dict[str, int]]:
    """The synthetic benchmark / parity configuration of SURVEY.md section 8(d):
    mFormerV1_<variant>, ``n_tasks`` ranks with classes (1000,400,120,40,12,4),
    three metadata components (2+3+10 dims).  Returns ``(cfg, num_classes)``."""
    depths, dims_, rdepths, heads_, _dp = _ARCH_V1[variant]
    dims = list(dims) if dims is not None else dims_
    heads = list(heads) if heads is not None else heads_
    rdepths = list(rope_depths) if rope_depths is not None else rdepths
    depths = list(conv_depths) if conv_depths is not None else depths
    c = get_default_config()
    c.MODEL.TYPE = "mFormerV1"
    c.MODEL.NAME = f"mFormerV1_{variant}"
    c.MODEL.IMG_SIZE = img_size
    c.DATA.IMG_SIZE = img_size
    c.MODEL.DROP_PATH_RATE = drop_path
    c.MODEL.CONVNEXT_STAGES = CfgNode(
        {"DEPTHS": list(depths), "DIMS": list(dims), "LAYER_SCALE_INIT_VALUE": 1e-6}, new_allowed=True
    )
    c.MODEL.ROPE_STAGES = CfgNode(
        {
            "DEPTHS": list(rdepths),
            "DIMS": [dims[2], dims[3]],
            "NUM_HEADS": list(heads),
            "MLP_RATIO": [4.0, 4.0],
            "ROPE_THETA": 10000.0,
            "ROPE_MIXED": True,
        },
        new_allowed=True,
    )
    tasks = SYNTH_TASKS[:n_tasks]
    c.DATA.TASK_KEYS_H5 = list(tasks)
    num_classes = dict(zip(tasks, SYNTH_CLASSES[:n_tasks]))
    for t in tasks:
        c.MODEL.CLASSIFICATION.HEADS[t] = CfgNode({"TYPE": head_type}, new_allowed=True)
    c.DATA.META.ACTIVE = bool(meta)
    if meta:
        for name, dim, idx in SYNTH_META:
            c.DATA.META.COMPONENTS[name] = CfgNode({"ENABLED": True, "DIM": dim, "IDX": idx}, new_allowed=True)
        c.MODEL.EXTRA_TOKEN_NUM = 1 + len(SYNTH_META)
    c.TRAIN.GRADIENT_CHECKPOINTING.ENABLED_NORMAL_STEPS = False
    return c, num_classes


class SyntheticTaxonomy:
    """The synthetic hierarchy of SURVEY.md 8(d) (child i -> parent 0 if i == 0 else 1 + (i - 1) mod (C_parent - 1)) with the one
    method the hierarchical head types call on a TaxonomyTree: ``build_hierarchy_matrices`` (R/utils/taxonomy/taxonomy_tree.py:384-404:
    ``{"<parent>_<child>": [C_parent, C_child] 0/1 matrix}``, registered by the heads as ``hmatrix_*`` buffers)."""

    def __init__(self, num_classes: dict[str, int]):
        self.task_keys = list(num_classes.keys())
        self.num_classes = dict(num_classes)

    def parent_of(self, child_task: str, i: int) -> int:
        cp = self.num_classes[self.task_keys[self.task_keys.index(child_task) + 1]]
        return 0 if i == 0 else 1 + (i - 1) % (cp - 1)

    def build_hierarchy_matrices(self) -> dict:
        import torch

        out = {}
        for child, parent in zip(self.task_keys[:-1], self.task_keys[1:]):
            m = torch.zeros((self.num_classes[parent], self.num_classes[child]), dtype=torch.float32)
            for i in range(self.num_classes[child]):
                m[self.parent_of(child, i), i] = 1.0
            out[f"{parent}_{child}"] = m
        return out


# ---------------------------------------------------------------------------
# mFormerV0 (config 5): arch table mirrored from configs/model/archs/mFormerV0/*.yaml
# ---------------------------------------------------------------------------
_ARCH_V0 = {
    #       stem  conv embed   conv out     conv depths  conv strides              attn dims     attn depths  heads
    "sm": (64, [64, 96], [96, 192], [2, 3], [[2, 1], [1, 1, 1]], [384, 768], [5, 2], [8, 8]),
}


def make_synthetic_config_v0(variant: str = "sm", img_size: int = 224, n_tasks: int = 6, meta: bool = True, conv_embed=None, conv_out=None,
                             conv_depths=None, conv_strides=None, attn_dims=None, attn_depths=None, heads=None) -> tuple[CfgNode, dict[str, int]]:
    """Synthetic mFormerV0 configuration: RelativeAttention variant, same synthetic heads /
    metadata components as the V1 benchmark.  Returns ``(cfg, num_classes)``."""
    stem, ce, co, cd, cstr, ad, adep, hd = _ARCH_V0[variant]
    ce = list(conv_embed) if conv_embed is not None else ce
    co = list(conv_out) if conv_out is not None else co
    cd = list(conv_depths) if conv_depths is not None else cd
    cstr = [list(s) for s in conv_strides] if conv_strides is not None else cstr
    ad = list(attn_dims) if attn_dims is not None else ad
    adep = list(attn_depths) if attn_depths is not None else adep
    hd = list(heads) if heads is not None else hd
    c = get_default_config()
    c.MODEL.TYPE = "mFormerV0"
    c.MODEL.NAME = f"mFormerV0_{variant}"
    c.MODEL.IMG_SIZE = img_size
    c.DATA.IMG_SIZE = img_size
    c.MODEL.DROP_PATH_RATE = 0.0
    c.MODEL.DROP_RATE = 0.0
    c.MODEL.ATTN_DROP_RATE = 0.0
    c.MODEL.CONV_STAGES = CfgNode({"STEM_OUT": ce[0], "EMBED_DIMS": ce, "OUT_CHANNELS": co, "DEPTHS": cd, "STRIDE_SEQS": cstr}, new_allowed=True)
    c.MODEL.ATTENTION_STAGES = CfgNode(
        {"EMBED_DIMS": ad, "DEPTHS": adep, "NUM_HEADS": hd, "MLP_RATIO": [4.0, 4.0],
         "ATTENTION_TYPE": ["RelativeAttention", "RelativeAttention"],
         "STRIDE_SEQS": [[2] + [1] * (adep[0] - 1), [2] + [1] * (adep[1] - 1)]}, new_allowed=True)
    tasks = SYNTH_TASKS[:n_tasks]
    c.DATA.TASK_KEYS_H5 = list(tasks)
    num_classes = dict(zip(tasks, SYNTH_CLASSES[:n_tasks]))
    for t in tasks:
        c.MODEL.CLASSIFICATION.HEADS[t] = CfgNode({"TYPE": "Linear"}, new_allowed=True)
    c.DATA.META.ACTIVE = bool(meta)
    if meta:
        for name, dim, idx in SYNTH_META:
            c.DATA.META.COMPONENTS[name] = CfgNode({"ENABLED": True, "DIM": dim, "IDX": idx}, new_allowed=True)
        c.MODEL.EXTRA_TOKEN_NUM = 1 + len(SYNTH_META)
    c.TRAIN.GRADIENT_CHECKPOINTING.ENABLED_NORMAL_STEPS = False
    return c, num_classes
