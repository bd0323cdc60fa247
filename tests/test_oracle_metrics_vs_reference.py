"""Build container only: the metrics oracle against the UNMODIFIED reference functions on fresh random inputs (skipped where
/root/reference is absent; the committed goldens cover that case)."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO
from tests.support import refload

pytestmark = pytest.mark.skipif(not refload.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_metrics_oracle_vs_reference(seed):
    refload.import_reference()
    from linnaeus.utils.metrics.basic import accuracy
    from linnaeus.utils.metrics.chain_accuracy import compute_chain_accuracy_vectorized, compute_partial_chain_accuracy_vectorized

    rng = np.random.default_rng(100 + seed)
    classes = [(50, 20, 6, 3), (9, 4), (300, 100, 30, 10, 4, 2), (2, 2, 2)][seed]
    B = [40, 7, 65, 128][seed]
    outs, tgts, onehots = [], [], []
    for C in classes:
        z = rng.standard_normal((B, C)).astype(np.float32)
        y = rng.integers(0, C, size=B)
        y[rng.random(B) < 0.4] = 0
        z[np.arange(B), y] += 3.0 * rng.random(B).astype(np.float32)
        outs.append(z)
        tgts.append(y.astype(np.int64))
        onehots.append(np.eye(C, dtype=np.float32)[y])
    to, tt, th = [torch.from_numpy(z) for z in outs], [torch.from_numpy(y) for y in tgts], [torch.from_numpy(h) for h in onehots]
    for z, y, a, b in zip(outs, tgts, to, tt):
        ks = tuple(k for k in (1, 2, 3, 5) if k <= z.shape[1])
        np.testing.assert_allclose(MO.accuracy(z, y, ks), accuracy(a, b, topk=ks), atol=1e-4)  # the reference divides in float32
        np.testing.assert_allclose(MO.accuracy(z, y, ks, ignore_index=0), accuracy(a, b, topk=ks, ignore_index=0), atol=1e-4)
    assert MO.chain_accuracy(outs, tgts) == pytest.approx(compute_chain_accuracy_vectorized(to, tt), abs=1e-12)
    assert MO.chain_accuracy(outs, onehots) == pytest.approx(compute_chain_accuracy_vectorized(to, th), abs=1e-12)
    assert MO.chain_accuracy(outs, tgts, ignore_index=0) == compute_chain_accuracy_vectorized(to, tt, ignore_index=0)
    assert MO.partial_chain_accuracy(outs, tgts) == pytest.approx(compute_partial_chain_accuracy_vectorized(to, tt), abs=1e-12)
    assert MO.partial_chain_accuracy(outs, onehots) == pytest.approx(compute_partial_chain_accuracy_vectorized(to, th), abs=1e-12)
