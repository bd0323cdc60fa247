"""Model / head registries and ``build_model`` -- the drop-in boundary.

Mirrors ``linnaeus/models/model_factory.py:94-213`` and ``linnaeus/models/build.py:52-111``:
``register_model(name)`` overwrites an existing entry (with a warning), and
``create_model(config, **kw)`` instantiates ``_model_registry[config.MODEL.TYPE]``.
:func:`install_into_linnaeus` re-registers the B200 models into the *reference's*
registry so ``linnaeus/main.py`` and ``LinnaeusInferenceHandler`` build them unchanged.
"""
from __future__ import annotations

import logging

import torch.nn as nn

logger = logging.getLogger("linnaeus_b200")

_model_registry: dict[str, type[nn.Module]] = {}
_head_registry: dict[str, type[nn.Module]] = {}


def register_model(name: str):
    def deco(cls):
        if name in _model_registry:
            logger.warning("Model '%s' already registered. Overwriting.", name)
        _model_registry[name] = cls
        return cls

    return deco


def register_head(name: str):
    def deco(cls):
        if name in _head_registry:
            logger.warning("Head '%s' already registered. Overwriting.", name)
        _head_registry[name] = cls
        return cls

    return deco


def create_head(name: str, **kwargs) -> nn.Module:
    if name not in _head_registry:
        raise ValueError(f"Head type '{name}' not found in registry. Available: {list(_head_registry)}")
    return _head_registry[name](**kwargs)


def list_models() -> list[str]:
    return sorted(_model_registry)


def create_model(config, **kwargs) -> nn.Module:
    name = config.MODEL.TYPE
    if name not in _model_registry:
        raise ValueError(f"Model type '{name}' not found in registry. Available models: {list_models()}")
    return _model_registry[name](config, **kwargs)


def build_model(config, num_classes: dict[str, int] | None = None, taxonomy_tree=None) -> nn.Module:
    """``linnaeus.models.build_model`` (build.py:52-111).  ``MODEL.PRETRAINED`` is loaded
    as a plain state_dict (strict=False); the reference's stitched-checkpoint remapping
    lives in its own ``utils/checkpoint.py`` and keeps working on this model because the
    parameter names are identical."""
    model = create_model(config=config, num_classes=num_classes, taxonomy_tree=taxonomy_tree)
    pretrained = config.MODEL.get("PRETRAINED", None) if hasattr(config.MODEL, "get") else None
    if pretrained:
        import torch

        sd = torch.load(pretrained, map_location="cpu")
        sd = sd.get("model", sd)
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
        model.load_state_dict(sd, strict=False)
    return model


def install_into_linnaeus() -> list[str]:
    """Re-register the B200 implementations under the reference's own registry
    (``linnaeus.models.model_factory``); returns the names replaced."""
    from linnaeus.models import model_factory as mf  # the reference package must be importable

    from . import mformer_v0, mformer_v1  # noqa: F401  (populate _model_registry)

    names = []
    for name, cls in _model_registry.items():
        mf.register_model(name)(cls)
        names.append(name)
    return names
