"""linnaeus_b200 -- B200 (sm_100a) native mFormer training / inference hot path.

Public surface (mirrors ``linnaeus.models`` / ``linnaeus.loss`` for this path):
``build_model(cfg, num_classes, taxonomy_tree)``, ``register_model``, ``create_model``,
``weighted_hierarchical_loss``, ``FlatAdamW``, ``DataParallel``, ``install_into_linnaeus``.

Either side of that path, imported on demand: ``linnaeus_b200.aug`` (selective mixup / CutMix feeding the
model), ``linnaeus_b200.metrics`` (validation metrics, inference top-k), ``linnaeus_b200.postprocess`` (hierarchical consistency
of the predictions), ``linnaeus_b200.gradnorm`` (GradNorm task weighting),
``linnaeus_b200.checkpoint`` (checkpoint interchange with the reference).
"""
from .config import CfgNode, get_default_config, make_synthetic_config, make_synthetic_config_v0  # noqa: F401
from .registry import build_model, create_model, install_into_linnaeus, list_models, register_head, register_model  # noqa: F401
from . import mformer_v1  # noqa: F401  (registers "mFormerV1")
from .mformer_v1 import mFormerV1  # noqa: F401
from . import mformer_v0  # noqa: F401  (registers "mFormerV0": inference path)
from .mformer_v0 import mFormerV0  # noqa: F401

__version__ = "0.1.0"
