"""CPU: the oracle must reproduce the committed reference outputs
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import mformer_oracle as O
from tests.support.golden import CASES, load_case


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_reference_golden(name):
    cfg, nc, kind, z = load_case(name)
    a = O.arch_from_config(cfg, nc)
    P = O.synth_state_dict(O.param_shapes(a), int(z["wseed"]))
    x, meta, tg = O.synth_batch(a, int(z["batch"]), int(z["dseed"]))
    leaves = {n: t.clone().requires_grad_(True) for n, t in P.items()}
    logits = O.forward(leaves, a, x, meta)
    mats = O.synthetic_taxonomy_smoothing(a.tasks) if kind == "taxonomy" else None
    total, _ = O.hierarchical_loss(logits, tg, kind=kind, soft_matrices=mats)
    total.backward()
    assert abs(float(total) - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    for t, _ in a.tasks:
        ref = torch.from_numpy(z[f"logits/{t}"])
        torch.testing.assert_close(logits[t].detach(), ref, rtol=1e-5, atol=3e-5)
        assert torch.equal(logits[t].argmax(1), ref.argmax(1))
    # hierarchical head types share ONE ModuleDict of per-level classifiers: the reference lists it once (under the first head),
    # the oracle's functional forward reads level t through head.<t>.<sub>.<t>
    first = a.tasks[0][0]

    def leaf_of(n):
        parts = n.split(".")
        if parts[0] == "head" and len(parts) == 5:
            return ".".join(["head", parts[3], parts[2], parts[3], parts[4]])
        return n

    names = [n for n in leaves if f"gnorm/{n}" in z.files and (not n.startswith("head.") or len(n.split(".")) != 5 or n.split(".")[1] == first)]
    assert len(names) == sum(1 for k in z.files if k.startswith("gnorm/"))
    gmax = max(float(z[f"gnorm/{n}"]) for n in names)
    for n in names:
        p = leaves[leaf_of(n)]
        gn = float(z[f"gnorm/{n}"])
        g = p.grad.flatten()
        assert abs(float(g.norm()) - gn) <= 1e-4 * gn + 1e-6 * gmax, n
        head = torch.from_numpy(z[f"ghead/{n}"])
        tol = 1e-4 * float(head.abs().max()) + 1e-6
        assert float((g[: head.numel()] - head).abs().max()) <= tol, n
