"""Checkpoint interchange with the reference (second half; section 5 "Checkpoint / resume").

The reference writes one ``torch.save`` dict per checkpoint - ``model``, ``optimizer``, ``lr_scheduler``, ``epoch``, ``config``,
``iteration`` (+ optional ``training_progress``, ``amp``, ``wandb_run_id``, ``metrics_tracker``) - as ``ckpt_epoch_{E}.pth`` plus
``latest.pth`` (R/utils/checkpoint.py:956-1131), adds / strips the DDP ``module.`` prefix on load (:21-42, :738-953) and resumes
from the newest ``*.pth`` by mtime (:1307-1332).  Because this package keeps the reference's parameter names and shapes, and
``FlatAdamW.state_dict()`` has ``torch.optim.AdamW``'s layout, the two sides read each other's files.  Host-side plumbing only:
no kernels here, the tensors move with ``torch.save`` / ``torch.load``."""
from __future__ import annotations

import os
from typing import Any

import torch


def clean_state_dict_keys(sd: dict[str, Any], model_is_ddp: bool, ckpt_has_module_prefix: bool) -> dict[str, Any]:
    """Add or strip the ``module.`` prefix so the keys match the target model (R/utils/checkpoint.py:21-42)."""
    if model_is_ddp and not ckpt_has_module_prefix:
        return {f"module.{k}": v for k, v in sd.items()}
    if not model_is_ddp and ckpt_has_module_prefix:
        return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    return dict(sd)


def save_checkpoint(directory: str, model, optimizer, epoch: int, config=None, lr_scheduler=None, iteration: int = 0,
                    extra: dict | None = None) -> str:
    """Write ``ckpt_epoch_{epoch}.pth`` and ``latest.pth`` with the reference's keys; returns the epoch file's path."""
    os.makedirs(directory, exist_ok=True)
    state = {
        "model": model.state_dict(),
        "optimizer": optimizer.state_dict() if optimizer is not None else None,
        "lr_scheduler": lr_scheduler.state_dict() if lr_scheduler is not None else None,
        "epoch": epoch,
        "config": config,
        "iteration": iteration,
    }
    if extra:
        state.update(extra)
    path = os.path.join(directory, f"ckpt_epoch_{epoch}.pth")
    torch.save(state, path)
    torch.save(state, os.path.join(directory, "latest.pth"))
    return path


def load_checkpoint(path: str, model, optimizer=None, lr_scheduler=None, strict: bool = True, map_location="cpu") -> dict:
    """Load a checkpoint written by either side: model weights (prefix-cleaned), optimizer and scheduler state when given.
    Returns the remaining entries (``epoch``, ``iteration``, ``config``, ...)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = ckpt["model"] if "model" in ckpt else ckpt
    # like the reference (utils/checkpoint.py:798-830) the target model's own keys decide: torch's DDP and this package's
    # DataParallel wrapper both expose ``module.``-prefixed names
    is_ddp = any(k.startswith("module.") for k in model.state_dict())
    has_prefix = any(k.startswith("module.") for k in sd)
    model.load_state_dict(clean_state_dict_keys(sd, is_ddp, has_prefix), strict=strict)
    if optimizer is not None and ckpt.get("optimizer") is not None:
        optimizer.load_state_dict(ckpt["optimizer"])
    if lr_scheduler is not None and ckpt.get("lr_scheduler") is not None:
        lr_scheduler.load_state_dict(ckpt["lr_scheduler"])
    return {k: v for k, v in ckpt.items() if k not in ("model", "optimizer", "lr_scheduler")}


def auto_resume_helper(output_dir: str, config=None) -> str | None:
    """The newest ``*.pth`` in ``output_dir`` by modification time, or None (R/utils/checkpoint.py:1307-1332)."""
    cks = [c for c in os.listdir(output_dir) if c.endswith(".pth")]
    if not cks:
        return None
    return os.path.join(output_dir, max(cks, key=lambda c: os.path.getmtime(os.path.join(output_dir, c))))
