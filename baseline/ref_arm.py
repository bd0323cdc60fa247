"""Reference arm of bench.py: the UNMODIFIED reference (pip-installed into baseline/_ref with
`python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref /root/reference`) driven through its own
public API on the host CPU - `linnaeus.models.build_model`, `linnaeus.loss.hierarchical_loss.weighted_hierarchical_loss`,
`linnaeus.optimizers.build.build_optimizer`, the clip + step sequence of `linnaeus/train.py:282-313`.  Two third-party packages the
reference imports (yacs, termcolor) cannot be installed offline; baseline/shims provides stand-ins.  None of this repo's kernels,
models or engine is on this path."""
from __future__ import annotations

import logging
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "linnaeus", "models"))


def _import():
    for p in (REF, os.path.join(HERE, "shims")):
        if p not in sys.path:
            sys.path.insert(0, p)
    logging.getLogger("linnaeus").setLevel(logging.ERROR)
    import linnaeus  # noqa: F401

    logging.getLogger("linnaeus").setLevel(logging.ERROR)


def _config_v1(variant: str, img: int):
    """The reference's default config + the arch section of configs/model/archs/mFormerV1/mFormerV1_<variant>.yaml (values mirrored
    in linnaeus_b200.config._ARCH_V1; the YAML files are not part of the installed package) + the synthetic heads / metadata."""
    from linnaeus.config import get_default_config
    from yacs.config import CfgNode as CN

    from linnaeus_b200.config import _ARCH_V1, SYNTH_CLASSES, SYNTH_META, SYNTH_TASKS

    depths, dims, rdepths, heads, _dp = _ARCH_V1[variant]
    cfg = get_default_config()
    cfg.defrost()
    cfg.MODEL.TYPE = "mFormerV1"
    cfg.MODEL.CONVNEXT_STAGES = CN({"DEPTHS": list(depths), "DIMS": list(dims), "LAYER_SCALE_INIT_VALUE": 1e-6}, new_allowed=True)
    cfg.MODEL.ROPE_STAGES = CN({"DEPTHS": list(rdepths), "DIMS": [dims[2], dims[3]], "NUM_HEADS": list(heads), "MLP_RATIO": [4.0, 4.0],
                                "ROPE_THETA": 10000.0, "ROPE_MIXED": True}, new_allowed=True)
    cfg.MODEL.IMG_SIZE = img
    cfg.MODEL.USE_FLASH_ATTN = False
    cfg.MODEL.DROP_PATH_RATE = 0.0
    cfg.MODEL.DROP_RATE = 0.0
    cfg.MODEL.ATTN_DROP_RATE = 0.0
    cfg.MODEL.PRETRAINED = None
    cfg.MODEL.ONLY_LAST_CLS = False
    cfg.DATA.TASK_KEYS_H5 = list(SYNTH_TASKS)
    cfg.MODEL.CLASSIFICATION.HEADS = CN(new_allowed=True)
    for t in SYNTH_TASKS:
        cfg.MODEL.CLASSIFICATION.HEADS[t] = CN({"TYPE": "Linear"}, new_allowed=True)
    cfg.DATA.META.ACTIVE = True
    cfg.DATA.META.COMPONENTS = CN(new_allowed=True)
    for name, dim, idx in SYNTH_META:
        cfg.DATA.META.COMPONENTS[name] = CN({"ENABLED": True, "DIM": dim, "IDX": idx}, new_allowed=True)
    cfg.MODEL.EXTRA_TOKEN_NUM = 1 + len(SYNTH_META)
    cfg.TRAIN.AMP_OPT_LEVEL = "O0"
    cfg.TRAIN.GRADIENT_CHECKPOINTING.ENABLED_NORMAL_STEPS = False
    cfg.LOSS.GRAD_WEIGHTING.TASK.TYPE = "static"
    cfg.LOSS.GRAD_WEIGHTING.TASK.GRADNORM_ENABLED = False
    cfg.LOSS.GRAD_WEIGHTING.CLASS.METHOD = "none"
    return cfg, dict(zip(SYNTH_TASKS, SYNTH_CLASSES))


def _config_v0(variant: str, img: int):
    from linnaeus.config import get_default_config
    from yacs.config import CfgNode as CN

    from linnaeus_b200.config import SYNTH_CLASSES, SYNTH_META, SYNTH_TASKS, make_synthetic_config_v0

    mine, _ = make_synthetic_config_v0(variant, img)
    cfg = get_default_config()
    cfg.defrost()
    cfg.MODEL.TYPE = "mFormerV0"
    cfg.MODEL.IMG_SIZE = img
    cfg.MODEL.DROP_PATH_RATE = 0.0
    cfg.MODEL.DROP_RATE = 0.0
    cfg.MODEL.ATTN_DROP_RATE = 0.0
    cfg.MODEL.PRETRAINED = None
    cfg.MODEL.ONLY_LAST_CLS = False
    cfg.MODEL.CONV_STAGES = CN({k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in mine.MODEL.CONV_STAGES.items()}, new_allowed=True)
    cfg.MODEL.ATTENTION_STAGES = CN({k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in mine.MODEL.ATTENTION_STAGES.items()},
                                    new_allowed=True)
    cfg.DATA.TASK_KEYS_H5 = list(SYNTH_TASKS)
    cfg.MODEL.CLASSIFICATION.HEADS = CN(new_allowed=True)
    for t in SYNTH_TASKS:
        cfg.MODEL.CLASSIFICATION.HEADS[t] = CN({"TYPE": "Linear"}, new_allowed=True)
    cfg.DATA.META.ACTIVE = True
    cfg.DATA.META.COMPONENTS = CN(new_allowed=True)
    for name, dim, idx in SYNTH_META:
        cfg.DATA.META.COMPONENTS[name] = CN({"ENABLED": True, "DIM": dim, "IDX": idx}, new_allowed=True)
    cfg.TRAIN.GRADIENT_CHECKPOINTING.ENABLED_NORMAL_STEPS = False
    return cfg, dict(zip(SYNTH_TASKS, SYNTH_CLASSES))


def _batch(nc: dict, batch: int, img: int, seed: int = 42):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 3, img, img, generator=g)
    meta = torch.randn(batch, 15, generator=g)
    tg = {k: torch.randint(0, c, (batch,), generator=g) for k, c in nc.items()}
    return x, meta, tg


def time_train(variant: str, img: int, batch: int, steps: int, warmup: int):
    """-> (img/s, median seconds per step, threads).  One step = forward + 6-rank hierarchical CE loss + backward + clip 5.0 + AdamW."""
    _import()
    from linnaeus.loss.basic_loss import CrossEntropyLoss
    from linnaeus.loss.gradient_weighting import GradientWeighting
    from linnaeus.loss.hierarchical_loss import weighted_hierarchical_loss
    from linnaeus.models import build_model
    from linnaeus.optimizers.build import build_optimizer

    torch.set_num_threads(os.cpu_count() or 1)
    cfg, nc = _config_v1(variant, img)
    cfg.LR_SCHEDULER.BASE_LR = 1e-4 * batch / 512.0
    cfg.TRAIN.CLIP_GRAD = 5.0
    torch.manual_seed(0)
    model = build_model(cfg, num_classes=nc, taxonomy_tree=None).train()
    opt = build_optimizer(cfg, model)
    keys = list(nc.keys())
    crit = {k: CrossEntropyLoss() for k in keys}
    gw = GradientWeighting(keys, cfg, "static")

    class Sched:
        def get_null_mask_prob(self, step):
            return 1.0

    x, meta, tg = _batch(nc, batch, img)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = model(x, meta)
        total, _, _ = weighted_hierarchical_loss(out, tg, crit, gw, Sched(), i, config=cfg)
        opt.zero_grad(set_to_none=True)
        total.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return batch / med, med, torch.get_num_threads()


def time_infer(arch: str, variant: str, img: int, batch: int, steps: int, warmup: int):
    _import()
    from linnaeus.models import build_model

    torch.set_num_threads(os.cpu_count() or 1)
    cfg, nc = (_config_v0 if arch == "v0" else _config_v1)(variant, img)
    torch.manual_seed(0)
    model = build_model(cfg, num_classes=nc, taxonomy_tree=None).eval()
    x, meta, _ = _batch(nc, batch, img)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            model(x, meta)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return batch / med, med, torch.get_num_threads()
