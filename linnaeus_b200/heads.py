"""Classification heads (parameter holders + the reference's wiring).

``configure_classification_heads`` follows ``linnaeus/models/heads/utils.py:162-364``:
Linear heads own ``fc``; hierarchical head types (HierarchicalSoftmax /
ConditionalClassifier) share ONE ModuleDict of per-level ``nn.Linear`` registered under
every head (``task_classifiers`` / ``level_classifiers``) plus the ``hmatrix_*`` buffers.
As shipped, the reference's hierarchical refinement never fires (the buffer key it looks
up is the reverse of the one registered: SURVEY F4; taxonomy_tree.py:384-404 vs
hierarchical_softmax_head.py:164-190), so each head's output is its own level's linear
logits -- which is what ``classifier_params`` exposes to the concatenated head GEMM.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .registry import create_head, register_head


@register_head("Linear")
class LinearHead(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = True):
        super().__init__()
        self.fc = nn.Linear(in_features, out_features, bias=bias)

    def classifier_params(self):
        return self.fc.weight, self.fc.bias


class _HierarchicalBase(nn.Module):
    _SUB = "level_classifiers"

    def __init__(self, in_features, task_key, task_keys, taxonomy_tree, num_classes, use_bias=True,
                 level_classifiers_override=None, **_ignored):
        super().__init__()
        if task_key not in task_keys:
            raise ValueError(f"Primary task key '{task_key}' not found in task_keys list.")
        if task_key not in num_classes:
            raise ValueError(f"num_classes missing for primary task key '{task_key}'")
        self.in_features = in_features
        self.primary_task_key = task_key
        self.task_keys = list(task_keys)
        self.num_classes = num_classes
        self.taxonomy_tree = taxonomy_tree
        self._gradnorm_mode = False
        if level_classifiers_override is not None:
            shared = level_classifiers_override
        else:
            shared = nn.ModuleDict({tk: nn.Linear(in_features, num_classes[tk], bias=use_bias) for tk in self.task_keys})
        setattr(self, self._SUB, shared)
        if taxonomy_tree is not None and hasattr(taxonomy_tree, "build_hierarchy_matrices"):
            for pair_key, matrix in taxonomy_tree.build_hierarchy_matrices().items():
                self.register_buffer(f"hmatrix_{pair_key}", matrix)

    # BaseHierarchicalHead API (base_hierarchical_head.py:4-18), used by GradNorm
    def set_gradnorm_mode(self, mode: bool) -> None:
        self._gradnorm_mode = bool(mode)

    def is_gradnorm_mode(self) -> bool:
        return self._gradnorm_mode

    def classifier_params(self):
        lin = getattr(self, self._SUB)[self.primary_task_key]
        return lin.weight, lin.bias


@register_head("HierarchicalSoftmax")
class HierarchicalSoftmaxHead(_HierarchicalBase):
    _SUB = "task_classifiers"


@register_head("ConditionalClassifier")
class ConditionalClassifierHead(_HierarchicalBase):
    _SUB = "level_classifiers"


def configure_classification_heads(heads_config, in_features, num_classes_dict=None, task_keys=None, taxonomy_tree=None,
                                   use_bias: bool = True) -> nn.ModuleDict:
    heads = nn.ModuleDict()
    if not isinstance(heads_config, dict):
        return heads
    hier = ("HierarchicalSoftmax", "ConditionalClassifier")
    any_hier = any(isinstance(c, dict) and c.get("TYPE", "") in hier for c in heads_config.values())
    shared = None
    if any_hier and task_keys and num_classes_dict:
        shared = nn.ModuleDict()
        for tk in task_keys:
            if num_classes_dict.get(tk) is None:
                raise ValueError(f"num_classes missing for task '{tk}'")
            shared[tk] = nn.Linear(in_features, num_classes_dict[tk], bias=use_bias)
    for task, cfg in heads_config.items():
        if not isinstance(cfg, dict):
            continue
        n_cls = num_classes_dict.get(task) if num_classes_dict else None
        if n_cls is None:
            n_cls = cfg.get("OUT_FEATURES")
            if n_cls is None:
                continue
        typ = cfg.get("TYPE", "Linear")
        bias = cfg.get("USE_BIAS", cfg.get("use_bias", use_bias))
        if typ == "Linear":
            heads[task] = create_head("Linear", in_features=in_features, out_features=n_cls, bias=bias)
        elif typ in hier:
            if not all([task_keys, taxonomy_tree is not None, num_classes_dict, shared is not None]):
                raise ValueError(f"Hierarchical context missing for hierarchical head '{task}'.")
            heads[task] = create_head(typ, in_features=in_features, task_key=task, task_keys=task_keys, taxonomy_tree=taxonomy_tree,
                                      num_classes=num_classes_dict, use_bias=bias, level_classifiers_override=shared)
        else:
            raise ValueError(f"Head type '{typ}' not found in registry.")
    return heads
