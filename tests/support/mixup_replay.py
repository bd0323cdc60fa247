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
