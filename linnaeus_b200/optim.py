"""Flat-buffer AdamW with fused global-norm clipping (the optimizer slice of the hot path).

Follows ``torch.optim.AdamW`` as configured by R/optimizers/build.py:67-106,687-716
(two groups: decay / no-decay = 1-D tensors and ``.bias``) and the clipping of
R/train.py:282-313 (``clip_grad_norm_(CLIP_GRAD)``; the reference's two extra norm
passes are logging only).  Parameters and gradients are re-homed into two contiguous
float32 buffers so that one ``lnx_sumsq`` + one ``lnx_adamw`` launch per group replaces
the multi-tensor foreach kernels; the clip coefficient never leaves the device.

The flat gradient buffers are also what ``linnaeus_b200.DataParallel`` all-reduces.
"""
from __future__ import annotations


import torch

from ._lib import call
from .flat import FlatGroup, weights_changed


def split_decay(named_params) -> tuple[list, list]:
    """set_weight_decay (optimizers/build.py:687-716)."""
    decay, no_decay = [], []
    for name, p in named_params:
        if not p.requires_grad:
            continue
        (no_decay if (p.ndim == 1 or name.endswith(".bias")) else decay).append((name, p))
    return decay, no_decay


class FlatAdamW(torch.optim.Optimizer):
    """``FlatAdamW(model.named_parameters(), lr=..., weight_decay=..., clip_grad=5.0)``.

    ``param_groups`` keeps the reference's layout (group 0 decay, group 1 no decay) so LR
    schedulers that write ``group['lr']`` work unchanged.  ``zero_grad`` always zeroes
    in place (``set_to_none`` would detach ``p.grad`` from the flat buffer)."""

    def __init__(self, named_params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05, clip_grad=0.0, grad_scale=1.0):
        named = list(named_params)
        if named and isinstance(named[0], torch.nn.Parameter):
            named = [(f"p{i}", p) for i, p in enumerate(named)]
        seen, uniq = set(), []
        for n, p in named:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append((n, p))
        decay, no_decay = split_decay(uniq)
        groups = [
            {"params": [p for _, p in decay], "weight_decay": weight_decay},
            {"params": [p for _, p in no_decay], "weight_decay": 0.0},
        ]
        super().__init__(groups, dict(lr=lr, betas=tuple(betas[:2]), eps=eps, weight_decay=weight_decay))
        dev = uniq[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdamW needs CUDA parameters (no CPU fallback exists)")
        self.flat = [FlatGroup(g["params"]) if g["params"] else None for g in self.param_groups]
        self.clip_grad = float(clip_grad)
        self.grad_scale = float(grad_scale)  # e.g. 1/world_size folded into the update
        self._step = 0
        self._scratch = torch.zeros(3, dtype=torch.float32, device=dev)  # sumsq | norm | coef
        self._norm_ws = torch.zeros(1024, dtype=torch.float32, device=dev)  # LNX_SUMSQ_WORKSPACE: deterministic global norm
        self._step_dev = torch.zeros(1, dtype=torch.float32, device=dev)  # device-side step count (CUDA-graph safe)
        self._lr_dev: torch.Tensor | None = None

    @property
    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm measured by the last ``step`` (0-dim device tensor)."""
        return self._scratch[1]

    def flat_grads(self) -> list[torch.Tensor]:
        return [f.g for f in self.flat if f is not None]

    def set_device_lr(self, lr: float | None) -> None:
        """Keep the learning rate in device memory (a captured CUDA graph then follows
        scheduler updates made with ``set_device_lr`` between replays)."""
        if lr is None:
            self._lr_dev = None
        elif self._lr_dev is None:
            self._lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=self._scratch.device)
        else:
            self._lr_dev.fill_(float(lr))

    def zero_grad(self, set_to_none: bool = False) -> None:  # noqa: ARG002
        for f in self.flat:
            if f is not None:
                f.g.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("closures are not supported")
        self._step += 1
        self._step_dev.add_(1.0)
        t = self._step
        sc = self._scratch
        sc[0:1].zero_()
        for f in self.flat:
            if f is not None:
                call("lnx_sumsq", f.g.data_ptr(), f.numel, sc.data_ptr(), self._norm_ws.data_ptr())
        call("lnx_clip_coef", sc.data_ptr(), self.grad_scale, self.clip_grad, sc[1:].data_ptr(), sc[2:].data_ptr())
        for g, f in zip(self.param_groups, self.flat):
            if f is None:
                continue
            b1, b2 = g["betas"]
            call("lnx_adamw", f.p.data_ptr(), f.g.data_ptr(), f.m.data_ptr(), f.v.data_ptr(), f.numel, float(g["lr"]), float(b1), float(b2),
                 float(g["eps"]), float(g["weight_decay"]), 1.0 - b1 ** t, 1.0 - b2 ** t, self.grad_scale, sc[2:].data_ptr(),
                 None if self._lr_dev is None else self._lr_dev.data_ptr(), self._step_dev.data_ptr())
        weights_changed()  # the kernel wrote the parameters behind torch's version counters
        return None

    # checkpoint interchange (R/utils/checkpoint.py:956-1200 saves ``optimizer.state_dict()`` of a torch.optim.AdamW, :738-953
    # loads it back): the state dict has torch.optim.AdamW's shape - per parameter index ``step`` / ``exp_avg`` / ``exp_avg_sq``
    # in param_groups order - so a reference checkpoint resumes here and one written here resumes in the reference.
    def step_count(self) -> int:
        """Optimizer steps taken so far.  The device counter is the source of truth: CUDA-graph replays advance it without
        running ``step()`` on the host (one device->host read; call it at checkpoint time, not per step)."""
        self._step = int(round(float(self._step_dev.item())))
        return self._step

    def state_dict(self):
        sd = super().state_dict()
        state = {}
        self.step_count()
        if self._step > 0:
            idx = 0
            for g, f in zip(self.param_groups, self.flat):
                for i, p in enumerate(g["params"]):
                    lo, hi = f.span(i)
                    n = p.numel()
                    state[idx] = {"step": torch.tensor(float(self._step)), "exp_avg": f.m[lo:lo + n].view_as(p).clone(),
                                  "exp_avg_sq": f.v[lo:lo + n].view_as(p).clone()}
                    idx += 1
        sd["state"] = state
        return sd

    def load_state_dict(self, sd):
        sd = dict(sd)
        legacy_m, legacy_v = sd.pop("flat_m", None), sd.pop("flat_v", None)  # round-1 development format
        legacy_step = sd.pop("flat_step", None)
        state = sd.get("state", {})
        super().load_state_dict({"state": {}, "param_groups": sd["param_groups"]})  # hyper-parameters; validates the group layout
        if legacy_m is not None:
            self._step = int(legacy_step or 0)
            for f, m, v in zip(self.flat, legacy_m, legacy_v):
                if f is not None:
                    f.m.copy_(m)
                    f.v.copy_(v)
        else:
            steps = set()
            idx = 0
            for g, f in zip(self.param_groups, self.flat):
                for i, p in enumerate(g["params"]):
                    st = state.get(idx, state.get(str(idx)))
                    lo, _ = f.span(i)
                    n = p.numel()
                    if st is None:  # torch leaves parameters that never received a gradient without state
                        f.m[lo:lo + n].zero_()
                        f.v[lo:lo + n].zero_()
                    else:
                        if tuple(st["exp_avg"].shape) != tuple(p.shape):
                            raise ValueError(f"optimizer state {idx}: shape {tuple(st['exp_avg'].shape)} != parameter shape {tuple(p.shape)}")
                        f.m[lo:lo + n].copy_(st["exp_avg"].reshape(-1))
                        f.v[lo:lo + n].copy_(st["exp_avg_sq"].reshape(-1))
                        steps.add(int(float(st["step"])))
                    idx += 1
            if len(steps) > 1:
                raise ValueError(f"FlatAdamW keeps one step count for all parameters; the state dict has {sorted(steps)}")
            self._step = steps.pop() if steps else 0
        self._step_dev.fill_(float(self._step))
        if self._lr_dev is not None:  # a captured graph reads the learning rate from device memory
            self.set_device_lr(self.param_groups[0]["lr"])


def build_optimizer(config, model) -> FlatAdamW:
    """AdamW slice of ``linnaeus.optimizers.build_optimizer`` (build.py:67-106)."""
    if config.OPTIMIZER.NAME.lower() != "adamw":
        raise NotImplementedError("linnaeus_b200 implements the AdamW path only (north_star scope)")
    return FlatAdamW(
        model.named_parameters(),
        lr=config.LR_SCHEDULER.BASE_LR,
        betas=tuple(config.OPTIMIZER.BETAS[:2]),
        eps=config.OPTIMIZER.EPS,
        weight_decay=config.OPTIMIZER.WEIGHT_DECAY,
        clip_grad=float(config.TRAIN.get("CLIP_GRAD", 0.0)),
    )
