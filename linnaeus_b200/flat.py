"""Contiguous float32 parameter / gradient buffers (device agnostic; plumbing only).

Parameters are re-homed as views of one flat tensor per group, gradients likewise, so
the optimizer is two flat launches and data-parallel reduction is a few large
all-reduces over contiguous slices (reverse-execution-order buckets)."""
from __future__ import annotations

import torch

# Bumped by everything that rewrites parameter memory WITHOUT going through a torch op on the parameter tensor itself (the fused
# AdamW kernel, the data-parallel broadcast of the flat buffer): such writes do not advance ``param._version``, and the cached
# bf16 weights of the inference path (mformer_v1._Bf16Shadow) key on this counter together with the parameter versions.
WEIGHTS_EPOCH = [0]


def weights_changed() -> None:
    WEIGHTS_EPOCH[0] += 1


class FlatGroup:
    def __init__(self, params: list[torch.nn.Parameter], with_state: bool = True):
        self.params = list(params)
        dev = self.params[0].device
        self.offsets, off = [], 0
        for p in self.params:
            # every tensor stays 16-byte aligned - and so does its bf16 shadow for matrices (8-element granularity): the tcgen05 GEMMs
            # take 16-byte aligned operands only (a [1,2,1] aggregate weight would otherwise push every later matrix onto the slow path)
            a = 8 if p.ndim >= 2 else 4
            off = (off + a - 1) // a * a
            self.offsets.append(off)
            off += p.numel()
        off = (off + 7) // 8 * 8
        self.numel = off
        self.p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.g = torch.zeros(off, dtype=torch.float32, device=dev)
        self.m = torch.zeros(off, dtype=torch.float32, device=dev) if with_state else None
        self.v = torch.zeros(off, dtype=torch.float32, device=dev) if with_state else None
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                n = p.numel()
                self.p[o:o + n].copy_(p.detach().reshape(-1))
                if p.grad is not None:
                    self.g[o:o + n].copy_(p.grad.reshape(-1))
                p.data = self.p[o:o + n].view(p.shape)
                p.grad = self.g[o:o + n].view(p.shape)
        weights_changed()

    def span(self, i: int) -> tuple[int, int]:
        return self.offsets[i], self.offsets[i] + self.params[i].numel()


def is_flat(p: torch.nn.Parameter, group: FlatGroup, i: int) -> bool:
    o = group.offsets[i]
    return p.data_ptr() == group.p.data_ptr() + 4 * o and p.grad is not None and p.grad.data_ptr() == group.g.data_ptr() + 4 * o
