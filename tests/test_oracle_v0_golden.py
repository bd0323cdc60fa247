"""CPU: the mFormerV0 oracle reproduces the committed reference logits (tests/golden/v0_*.npz)."""
import pytest
import torch

from oracle import mformer_v0_oracle as V
from tests.support.golden_v0 import CASES, load_case


@pytest.mark.parametrize("name", list(CASES))
def test_v0_oracle_reproduces_reference_golden(name):
    cfg, nc, batch, wseed, dseed, z = load_case(name)
    a = V.arch_from_config(cfg, nc)
    P = V.synth_state_dict(a, wseed)
    x, m = V.synth_batch(a, batch, dseed)
    with torch.no_grad():
        out = V.forward(P, a, x, m)
    for t, _ in a.tasks:
        ref = torch.from_numpy(z[f"logits.{t}"])
        torch.testing.assert_close(out[t], ref, rtol=1e-5, atol=3e-5)
        assert torch.equal(out[t].argmax(1), ref.argmax(1))
