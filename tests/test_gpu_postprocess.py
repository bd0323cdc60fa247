"""GPU: lnx_hier_consistency (R/inference/postprocessing.py:14-171 for a whole batch) against the frozen outputs of the unmodified
reference function and against the oracle on fresh random cases (bit-exact: integer / copy work), chained behind lnx_hier_topk, and
through the reference-named single-result wrapper."""
import os
from dataclasses import dataclass, field
from typing import Any

import numpy as np
import pytest
import torch

from tests.support.consistency_cases import LEVELS, make_case, taxon_id

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"


def _run_kernel(idx, prob, parent, offs, null_idx):
    import linnaeus_b200.postprocess as PP

    d_idx = torch.tensor(idx, dtype=torch.int32, device=DEV)
    d_prob = torch.tensor(prob, dtype=torch.float32, device=DEV)
    ch = PP.enforce_consistency_batch(d_idx, d_prob, torch.tensor(parent, dtype=torch.int32, device=DEV), [int(o) for o in offs], [int(n) for n in null_idx])
    return d_idx.cpu().numpy(), d_prob.cpu().numpy(), ch.cpu().numpy()


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_kernel_matches_reference_golden(name):
    z = np.load(os.path.join(GOLD, f"consistency_{name}.npz"))
    oi, op, ch = _run_kernel(z["idx"], z["prob"], z["parent"], z["class_off"], z["null_idx"])
    assert np.array_equal(oi, z["out_idx"])
    assert np.array_equal(op, z["out_prob"])
    assert np.array_equal(ch, z["changed"])


@pytest.mark.parametrize("seed,K,B,kk,null_links,nulls_known", [(21, 6, 1000, 5, True, True), (22, 2, 1, 1, False, True), (23, 7, 257, 2, True, False),
                                                                 (24, 1, 64, 3, True, True)])
def test_kernel_matches_oracle(seed, K, B, kk, null_links, nulls_known):
    from oracle.postprocess_oracle import enforce_consistency

    case = make_case(seed, K=K, B=B, kk=kk, null_links=null_links)
    null_idx = case["null_idx"] if nulls_known else [0 if k % 2 == 0 else -1 for k in range(K)]  # ranks without a null class keep their prediction
    parent = np.concatenate([np.asarray(r, np.int32) for r in case["parent"]])
    offs = np.cumsum([0] + [len(r) for r in case["parent"]])
    oi, op, ch = _run_kernel(case["idx"], case["prob"], parent, offs, null_idx)
    for b in range(B):
        preds = [[(int(c), float(p)) for c, p in zip(case["idx"][k, b], case["prob"][k, b])] for k in range(K)]
        got, changed = enforce_consistency(preds, case["parent"], null_idx)
        for k in range(K):
            assert bool(ch[k, b]) == changed[k]
            if changed[k]:
                assert oi[k, b].tolist() == [null_idx[k]] + [-1] * (kk - 1) and op[k, b].tolist() == [1.0] + [0.0] * (kk - 1)
            else:
                assert oi[k, b].tolist() == [c for c, _ in got[k]] and op[k, b].tolist() == [np.float32(p) for _, p in got[k]]


def test_topk_then_consistency_on_logits():
    """logits -> lnx_hier_topk -> lnx_hier_consistency in two launches and one read-back; top-1 of unchanged rows = argmax."""
    import linnaeus_b200.postprocess as PP
    from oracle.postprocess_oracle import enforce_consistency

    case = make_case(31, K=5, B=128, kk=3)
    keys = case["task_keys"]
    g = torch.Generator(device=DEV).manual_seed(3)
    outputs = {t: torch.randn(128, case["num_classes"][t], device=DEV, generator=g) for t in keys}
    parent = torch.tensor(np.concatenate([np.asarray(r, np.int32) for r in case["parent"]]), device=DEV)
    offs = tuple(int(o) for o in np.cumsum([0] + [len(r) for r in case["parent"]]))
    res = PP.topk_consistent_predictions(outputs, parent, offs, case["null_idx"], k=3)
    for b in range(128):
        preds = []
        for t in keys:
            p = torch.softmax(outputs[t][b].float(), 0)
            v, i = torch.topk(p, 3)
            preds.append([(int(c), float(x)) for c, x in zip(i.tolist(), v.tolist())])
        got, changed = enforce_consistency(preds, case["parent"], case["null_idx"])
        for k, t in enumerate(keys):
            idx, prob, ch = res[t]
            assert bool(ch[b]) == changed[k]
            assert int(idx[b, 0]) == got[k][0][0]
            if not changed[k]:
                assert abs(float(prob[b, 0]) - got[k][0][1]) < 1e-5


@dataclass
class _Task:
    rank_level: Any
    temperature: float
    predictions: list = field(default_factory=list)


@dataclass
class _Result:
    taxonomy_context: Any
    tasks: list
    subtree_roots: Any = None


class _Rank:
    def __init__(self, v):
        self.value, self.name = v, f"L{v}"

    def __hash__(self):
        return hash(self.value)

    def __eq__(self, o):
        return isinstance(o, _Rank) and o.value == self.value


class _Tree:
    def __init__(self, case):
        self.task_keys, self.num_classes = case["task_keys"], case["num_classes"]
        self._p = {(t, c): (case["task_keys"][k + 1], p) for k, t in enumerate(case["task_keys"]) for c, p in enumerate(case["parent"][k]) if p >= 0}

    def get_parent(self, node):
        return self._p.get(node)


def test_reference_named_wrapper_single_result():
    """enforce_hierarchical_consistency(result, taxonomy_data, class_maps) with duck-typed objects: same outputs as the oracle."""
    import types

    import linnaeus_b200.postprocess as PP
    from oracle.postprocess_oracle import enforce_consistency

    case = make_case(41, K=6, B=24, kk=3)
    K = 6
    ranks = [_Rank(LEVELS[k]) for k in range(K)]
    i2t = {r: {c: taxon_id(r.value, c) for c in range(case["num_classes"][case["task_keys"][k]])} for k, r in enumerate(ranks)}
    maps = types.SimpleNamespace(idx_to_taxon_id=i2t, taxon_id_to_idx={r: {t: c for c, t in m.items()} for r, m in i2t.items()},
                                 null_taxon_ids={r: i2t[r][0] for r in ranks})
    tdata = types.SimpleNamespace(taxonomy_tree=_Tree(case))
    for b in range(24):
        tasks = [_Task(ranks[k], 1.0, [(i2t[ranks[k]][int(c)], float(p)) for c, p in zip(case["idx"][k, b], case["prob"][k, b])]) for k in range(K)]
        out = PP.enforce_hierarchical_consistency(_Result(None, tasks), tdata, maps)
        assert [t.rank_level.value for t in out.tasks] == sorted((r.value for r in ranks), reverse=True)
        preds = [[(int(c), float(p)) for c, p in zip(case["idx"][k, b], case["prob"][k, b])] for k in range(K)]
        got, _ = enforce_consistency(preds, case["parent"], case["null_idx"])
        by = {t.rank_level.value: t.predictions for t in out.tasks}
        for k in range(K):
            assert by[LEVELS[k]] == [(i2t[ranks[k]][c], p) for c, p in got[k]]
