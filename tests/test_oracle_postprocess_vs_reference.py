"""CPU: the consistency oracle against the UNMODIFIED reference function R/inference/postprocessing.py::enforce_hierarchical_consistency
(build container only: needs /root/reference; `typus` is shimmed under tests/support/ref_shims), sample by sample on seeded random
taxonomies and predictions, with and without tree links for the null classes."""
import importlib
import os
import sys
import types

import pytest

from tests.support.consistency_cases import LEVELS, make_case, taxon_id
from tests.support.refload import REF_ROOT, import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")


def _reference_modules():
    import_reference()
    if "linnaeus.inference" not in sys.modules:  # the package __init__ pulls the whole handler (PIL, full typus): load the two files only
        pkg = types.ModuleType("linnaeus.inference")
        pkg.__path__ = [os.path.join(REF_ROOT, "linnaeus", "inference")]
        sys.modules["linnaeus.inference"] = pkg
    pp = importlib.import_module("linnaeus.inference.postprocessing")
    art = importlib.import_module("linnaeus.inference.artifacts")
    from linnaeus.utils.taxonomy.taxonomy_tree import TaxonomyTree
    from typus.constants import RankLevel
    from typus.models.classification import HierarchicalClassificationResult, TaskPrediction

    return pp, art, TaxonomyTree, RankLevel, HierarchicalClassificationResult, TaskPrediction


def reference_objects(case):
    pp, art, TaxonomyTree, RankLevel, Result, TaskPrediction = _reference_modules()
    K = len(case["task_keys"])
    tree = TaxonomyTree(case["hierarchy_map"], case["task_keys"], case["num_classes"])
    tdata = art.TaxonomyData(taxonomy_tree=tree, source="synthetic", linnaeus_task_keys=case["task_keys"])
    rls = [RankLevel(LEVELS[k]) for k in range(K)]
    i2t = {rl: {c: taxon_id(LEVELS[k], c) for c in range(case["num_classes"][case["task_keys"][k]])} for k, rl in enumerate(rls)}
    t2i = {rl: {t: c for c, t in m.items()} for rl, m in i2t.items()}
    maps = art.ClassIndexMapData(idx_to_taxon_id=i2t, taxon_id_to_idx=t2i, null_taxon_ids={rl: i2t[rl][0] for rl in rls},
                                 num_classes_per_rank={rl: len(i2t[rl]) for rl in rls})
    return pp, tdata, maps, rls, Result, TaskPrediction


def reference_run(case, b, objs):
    pp, tdata, maps, rls, Result, TaskPrediction = objs
    K = len(rls)
    tasks = [TaskPrediction(rank_level=rls[k], temperature=1.0,
                            predictions=[(maps.idx_to_taxon_id[rls[k]][int(c)], float(p)) for c, p in zip(case["idx"][k, b], case["prob"][k, b])])
             for k in range(K)]
    out = pp.enforce_hierarchical_consistency(Result(taxonomy_context=None, tasks=tasks), tdata, maps)
    by_rank = {t.rank_level: t.predictions for t in out.tasks}
    return [[(maps.taxon_id_to_idx[rls[k]][tid], p) for tid, p in by_rank[rls[k]]] for k in range(K)]


@pytest.mark.parametrize("seed,K,null_links", [(0, 6, True), (1, 6, False), (2, 3, True), (3, 4, False), (4, 7, True), (5, 2, True)])
def test_oracle_matches_reference_function(seed, K, null_links):
    from oracle.postprocess_oracle import enforce_consistency

    case = make_case(seed, K=K, B=48, null_links=null_links)
    objs = reference_objects(case)
    n_changed = 0
    for b in range(case["idx"].shape[1]):
        preds = [[(int(c), float(p)) for c, p in zip(case["idx"][k, b], case["prob"][k, b])] for k in range(K)]
        got, changed = enforce_consistency(preds, case["parent"], case["null_idx"])
        ref = reference_run(case, b, objs)
        assert got == ref, (seed, b, got, ref)
        n_changed += sum(changed)
    assert n_changed > 0  # the cases exercise the nullification branches
