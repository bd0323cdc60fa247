"""CPU: the selective-mixup oracle reproduces the committed outputs of the unmodified reference class (tests/golden/aug_*.npz)
given the draws that run used."""
import numpy as np
import pytest

from oracle import aug_oracle as AO
from tests.support.golden_aug import CASES, covered_columns, load_golden, make_case


@pytest.mark.parametrize("name", list(CASES))
def test_aug_oracle_reproduces_reference_golden(name):
    c, g = make_case(name), load_golden(name)
    eff = AO.exclude_null_group_ids(c["targets"], c["group_ids"])
    assert np.array_equal(eff, g["eff_gids"])
    assert AO.is_ingroup_permutation(g["perm"], eff)
    aux, masks = c["aux"].copy(), c["masks"].copy()
    mi, mt, ma, mm = AO.mixup_apply(c["images"], c["targets"], aux, masks, g["perm"], g["lam"], g["pick"], c["chunks"])
    assert np.array_equal(mi, g["mixed_images"])  # bit-exact: same fp32 expression
    for k in c["targets"]:
        assert np.array_equal(mt[k], g["mixed_targets." + k])
    cov = covered_columns(c["chunks"], c["aux"].shape[1])  # outside the chunks the reference returns torch.empty_like garbage
    assert np.array_equal(ma[:, cov], g["mixed_aux"][:, cov]) and np.array_equal(mm[:, cov], g["mixed_masks"][:, cov])
    assert np.array_equal(aux, g["aux_after"]) and np.array_equal(masks, g["masks_after"])  # the in-place side effect


def test_aug_oracle_chunk_rules():
    aux = np.array([[1, 2, 0, 0, 0], [0, 0, 3, 4, 5], [6, 7, 8, 0, 9], [1, 1, 1, 1, 1]], dtype=np.float32)
    mask = aux != 0
    perm = np.array([1, 0, 3, 2])
    _, _, oa, om = AO.mixup_apply(np.zeros((4, 1), np.float32), {}, aux, mask, perm, 0.5, np.array([0.9, 0.1, 0.9, 0.1], np.float32), [(0, 2), (2, 5)])
    assert aux[2].tolist() == [6, 7, 0, 0, 0]  # the partially zero chunk was wiped in place
    assert oa.tolist() == [[1, 2, 3, 4, 5], [1, 2, 3, 4, 5], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1]]
    assert om.all()
    assert not AO.is_ingroup_permutation(np.array([1, 0, 2]), np.array([0, 1, 1]))
    assert AO.is_ingroup_permutation(np.array([0, 2, 1]), np.array([-1, 1, 1]))


@pytest.mark.parametrize("name", ["aug_three_chunks", "aug_single_chunk"])
def test_cutmix_oracle_reproduces_reference_golden(name):
    c, g = make_case(name), load_golden(name.replace("aug_", "cutmix_"))
    aux, masks = c["aux"].copy(), c["masks"].copy()
    mi, mt, ma, mm = AO.cutmix_apply(c["images"], c["targets"], aux, masks, g["eff_gids"], g["perm"], tuple(g["box"]), g["pick"], c["chunks"])
    assert np.array_equal(mi, g["mixed_images"])
    for k in c["targets"]:
        assert np.array_equal(mt[k], g["mixed_targets." + k])
    assert np.array_equal(ma, g["mixed_aux"]) and np.array_equal(mm, g["mixed_masks"])
    assert np.array_equal(aux, g["aux_after"]) and np.array_equal(masks, g["masks_after"])
