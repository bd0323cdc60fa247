"""CPU: the metrics oracle reproduces the committed reference outputs (tests/golden/metrics_*.npz, produced by the unmodified
reference functions through tests/golden/make_golden_metrics.py)."""
import numpy as np
import pytest

from oracle import metrics_oracle as MO
from tests.support.golden_metrics import CASES, load_golden, make_case


@pytest.mark.parametrize("name", list(CASES))
def test_metrics_oracle_reproduces_reference_golden(name):
    keys, batches = make_case(name)
    g = load_golden(name)
    for bi, (outputs, targets) in enumerate(batches):
        ol, tl = [outputs[k] for k in keys], [targets[k] for k in keys]
        for k in keys:
            C = outputs[k].shape[1]
            ks = tuple(x for x in (1, 3, 5) if x <= C)
            np.testing.assert_allclose(MO.accuracy(outputs[k], targets[k], ks), g[f"b{bi}.{k}.acc"], rtol=0, atol=1e-4)  # the reference divides in float32
            np.testing.assert_allclose(MO.accuracy(outputs[k], targets[k], ks, ignore_index=0), g[f"b{bi}.{k}.acc_ignore0"], rtol=0, atol=1e-4)
            idx, prob = MO.softmax_topk(outputs[k], 5)
            assert np.array_equal(idx, g[f"b{bi}.{k}.topk_idx"])
            np.testing.assert_allclose(prob, g[f"b{bi}.{k}.topk_prob"], rtol=1e-5, atol=1e-7)
        assert MO.chain_accuracy(ol, tl) == pytest.approx(float(g[f"b{bi}.chain"]), abs=1e-12)
        assert MO.chain_accuracy(ol, tl, ignore_index=0) == 0.0
        assert MO.partial_chain_accuracy(ol, tl) == pytest.approx(float(g[f"b{bi}.partial"]), abs=1e-12)
    m = MO.phase_metrics(batches, keys)
    np.testing.assert_allclose([m["acc1"][k] for k in keys], g["acc1"], rtol=0, atol=1e-9)
    np.testing.assert_allclose([m["acc3"][k] for k in keys], g["acc3"], rtol=0, atol=1e-9)
    assert m["chain_accuracy"] == pytest.approx(float(g["chain_accuracy"]), abs=1e-12)
    assert m["partial_chain_accuracy"] == pytest.approx(float(g["partial_chain_accuracy"]), abs=1e-12)
    for f in ("null_acc1", "non_null_acc1"):
        assert m[f] == pytest.approx({k: float(v) for k, v in zip(keys, g[f]) if not np.isnan(v)}, abs=1e-9)


def test_rank_tie_rule_is_first_index():
    z = np.array([[1.0, 3.0, 3.0, 3.0, 0.0]], dtype=np.float32)
    assert [int(MO.target_rank(z, np.array([y]))[0]) for y in range(5)] == [3, 0, 1, 2, 4]
    # one-hot / soft targets are arg-maxed like the reference does
    onehot = np.eye(5, dtype=np.float32)[[1]]
    assert MO.chain_accuracy([z], [onehot]) == 1.0
    assert MO.accuracy(z, np.array([3]), topk=(1, 2, 3)) == [0.0, 0.0, 100.0]
    # counters of an empty-non-null batch: partial chain falls back to 1.0
    assert MO.partial_chain_accuracy([z], [np.array([0])]) == 1.0
