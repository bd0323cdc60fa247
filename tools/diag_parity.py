"""Prints the measured parity margins of the CUDA model against the reference goldens and the CPU oracle (what the tolerances in
tests/test_gpu_model.py are set from).  python tools/diag_parity.py [case ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests import test_gpu_model as T

cases = sys.argv[1:] or ["tiny_ce", "tiny_hsm", "tiny_cond", "sm224_ce", "sm224_b32", "xl384_shallow"]
for name in cases:
    L, O, cfg, nc, kind, z, a, P, x, meta, tg, model = T._setup(name)
    leaves = {n: t.clone().requires_grad_(True) for n, t in P.items()}
    lo = O.forward(leaves, a, x, meta)
    to, _ = O.hierarchical_loss(lo, tg, kind=kind, soft_matrices=O.synthetic_taxonomy_smoothing(a.tasks) if kind == "taxonomy" else None)
    to.backward()
    for dtype in (torch.float32, torch.bfloat16):
        model.zero_grad(set_to_none=True)
        model.set_compute_dtype(dtype).train()
        out, total = T._loss(L, O, model, a, kind, x, meta, tg, cfg)
        total.backward()
        le = abs(float(total.detach()) - float(z["loss"])) / abs(float(z["loss"]))
        lg = max(float((out[t].detach().float().cpu() - torch.from_numpy(z[f"logits/{t}"])).abs().max() / torch.from_numpy(z[f"logits/{t}"]).abs().max())
                 for t, _ in a.tasks)
        gmax = max(float(z[f"gnorm/{n}"]) for n, _ in model.named_parameters())
        errs = []
        for n, p in model.named_parameters():
            if n in T.ZERO_GRAD:
                continue
            ref = leaves[T.oracle_leaf(n)].grad if hasattr(T, "oracle_leaf") else leaves[n].grad
            g = p.grad.detach().float().cpu()
            l2 = float((g - ref).norm() / (ref.norm() + 1e-4 * gmax))
            gn = float(z[f"gnorm/{n}"])
            ne = abs(float(g.norm()) - gn) / (gn + 1e-3 * gmax)
            mx = float((g - ref).abs().max() / (ref.abs().max() + 1e-6 * gmax))
            errs.append((l2, ne, mx, n))
        errs.sort(reverse=True)
        med = sorted(e[0] for e in errs)[len(errs) // 2]
        print(f"{name:14s} {str(dtype):15s} loss {le:.2e} logits {lg:.2e} grad relL2 median {med:.2e} max-elem worst {max(e[2] for e in errs):.2e} "
              f"norm-err worst {max(e[1] for e in errs):.2e}")
        for l2, ne, mx, n in errs[:4]:
            print(f"      relL2 {l2:.2e} norm {ne:.2e} maxelem {mx:.2e}  {n}")
        rest = [e for e in errs if not e[3].startswith("meta_")]
        print("      -- without the metadata heads (ReLU-flip family):")
        for l2, ne, mx, n in rest[:6]:
            print(f"      relL2 {l2:.2e} norm {ne:.2e} maxelem {mx:.2e}  {n}")
