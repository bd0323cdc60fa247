"""Measurement of the 8(f) N3 row: selective-mixup apply step at the bench shapes (B = 256, 3x224x224 fp32 images, 6 one-hot
targets, 15 metadata columns in 3 chunks) against the reference's call sequence (torch ops + the per-(sample, chunk) Python loop
of R/aug/gpu/selective_mixup.py:394-560, restated) on the same GPU tensors."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linnaeus_b200.aug as A

dev = "cuda"
B, classes, chunks = 256, (1000, 400, 120, 40, 12, 4), [(0, 2), (2, 5), (5, 15)]
torch.manual_seed(0)
images = torch.randn(B, 3, 224, 224, device=dev)
targets = {f"taxa_L{10 * (i + 1)}": torch.nn.functional.one_hot(torch.randint(0, C, (B,), device=dev), C).float() for i, C in enumerate(classes)}
aux = torch.randn(B, 15, device=dev)
aux[torch.rand(B, 15, device=dev) < 0.05] = 0
masks = aux != 0
perm = torch.randperm(B, device=dev)
lam = torch.tensor(0.3, device=dev)
pick = torch.rand(B, device=dev)


def ours():
    return A.mixup_apply(images, targets, aux.clone(), masks.clone(), perm, lam, pick, chunks)


def reference_sequence():
    mi = lam * images + (1 - lam) * images[perm]
    mt = {k: lam * v + (1 - lam) * v[perm] for k, v in targets.items()}
    a1, m1 = aux.clone(), masks.clone()
    for lo, hi in chunks:
        part = (a1[:, lo:hi] == 0).any(dim=1)
        a1[part, lo:hi] = 0.0
        m1[part, lo:hi] = False
    a2, m2 = a1[perm], m1[perm]
    oa, om = torch.empty_like(a1), torch.empty_like(m1)
    for i in range(B):
        rnd = pick[i].item()
        for lo, hi in chunks:
            z1, z2 = bool(torch.all(a1[i, lo:hi] == 0)), bool(torch.all(a2[i, lo:hi] == 0))
            if not z1 and not z2:
                src = (a1, m1) if rnd < 0.5 else (a2, m2)
            elif not z1:
                src = (a1, m1)
            elif not z2:
                src = (a2, m2)
            else:
                oa[i, lo:hi], om[i, lo:hi] = 0.0, False
                continue
            oa[i, lo:hi], om[i, lo:hi] = src[0][i, lo:hi], src[1][i, lo:hi]
    return mi, mt, oa, om


out = torch.empty_like(images)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
from linnaeus_b200._lib import call
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
ms = []
for _ in range(8):
    flush.zero_()
    e0.record()
    call("lnx_mix_pairs", images.data_ptr(), perm.data_ptr(), lam.data_ptr(), out.data_ptr(), B, images.numel() // B)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms = sorted(ms)[len(ms) // 2]
nbytes = 3 * images.numel() * 4
print(f"lnx_mix_pairs images [256,3,224,224] f32: {ms:.3f} ms, algorithmic {nbytes / 1e6:.0f} MB -> {nbytes / ms / 1e6:.0f} GB/s (L2 flushed between launches)")


def timeit(fn, n):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print(f"mixup_apply (images + 6 targets + metadata), 9 launches, no sync: {timeit(ours, 20):.3f} ms per batch")
print(f"reference call sequence on the same GPU tensors:                {timeit(reference_sequence, 2):.1f} ms per batch ({B * (1 + 2 * len(chunks))} host syncs)")
