"""Replays the RNG calls of the reference's GPUSelectiveMixup.__call__ (R/aug/gpu/selective_mixup.py:140-330, 326-369,
:474) on the CPU generator so a test knows the draws a seeded reference run used: rand(1) [probability gate], one randperm per
group with more than one member in unique() order, the Beta(alpha, alpha) sample, rand(B) [metadata picks]."""
import torch


def replay_draws(group_ids: torch.Tensor, alpha: float, seed: int):
    """-> (gate float, perm int64 [B], lam float32 0-dim, pick float32 [B]) for torch.manual_seed(seed) on the CPU."""
    torch.manual_seed(seed)
    gate = torch.rand(1).item()
    B = group_ids.shape[0]
    perm = torch.arange(B)
    for g in group_ids.unique():
        if g.item() == -1:
            continue
        idx = (group_ids == g).nonzero(as_tuple=True)[0]
        if idx.numel() > 1:
            perm[idx] = idx[torch.randperm(idx.numel())]
    lam = torch.distributions.beta.Beta(alpha, alpha).sample()
    pick = torch.rand(B)
    return gate, perm, lam, pick


def replay_cutmix_draws(group_ids: torch.Tensor, alpha: float, seed: int, size, minmax=None):
    """The same for GPUSelectiveCutMix.__call__ (R/aug/gpu/selective_cutmix.py:176-213, :513): rand(1), the per-group randperms,
    the Beta sample (+ MINMAX rescale), two Python random.randint calls inside rand_bbox (seed with random.seed(seed)), rand(B).
    -> (perm, lam float, (cx, cy), pick)"""
    import random

    torch.manual_seed(seed)
    random.seed(seed)
    torch.rand(1)
    B = group_ids.shape[0]
    perm = torch.arange(B)
    for g in group_ids.unique():
        if g.item() == -1:
            continue
        idx = (group_ids == g).nonzero(as_tuple=True)[0]
        if idx.numel() > 1:
            perm[idx] = idx[torch.randperm(idx.numel())]
    lam = torch.distributions.beta.Beta(alpha, alpha).sample()
    if minmax is not None:
        lam = minmax[0] + (minmax[1] - minmax[0]) * lam
    cx, cy = random.randint(0, size[2]), random.randint(0, size[3])
    pick = torch.rand(B)
    return perm, lam.item(), (cx, cy), pick
