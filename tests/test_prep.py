"""Host-side prep mirrors the reference: the naming tests of tests/testthat/test-resnmtf.R:189-332, the
check_* error strings (R/utils.r:220-454), restriction matrices, shared-name maps."""
import warnings

import numpy as np
import pytest

from resnmtf_b200 import prep
from resnmtf_b200.prep import NamedMatrix


def two(n1=10, n2=10, seed=0):
    rng = np.random.default_rng(seed)
    return [np.abs(rng.standard_normal((n1, n1))), np.abs(rng.standard_normal((n2, n2)))]


def test_row_restriction_with_mismatched_unnamed_rows_errors():
    with pytest.raises(ValueError, match="Row restriction matrices implies shared rows between views"):
        prep.give_names(two(10, 20), 2, phi=np.ones((2, 2)))


def test_row_names_copied_when_row_restriction_given():
    out = prep.give_names(two(), 2, phi=np.ones((2, 2)))
    assert out["row_names"][0] == out["row_names"][1]


def test_col_restriction_with_mismatched_unnamed_cols_errors():
    with pytest.raises(ValueError, match="Column restriction matrices implies shared columns between"):
        prep.give_names(two(10, 20), 2, psi=np.ones((2, 2)))


def test_col_names_copied_when_col_restriction_given():
    out = prep.give_names(two(), 2, psi=np.ones((2, 2)))
    assert out["col_names"][0] == out["col_names"][1]


def _named(rows0, rows1, cols0, cols1):
    a, b = two()
    return [NamedMatrix(a, rows0, cols0), NamedMatrix(b, rows1, cols1)]


R10 = [f"row_{i}" for i in range(1, 11)]
R514 = [f"row_{i}" for i in range(5, 15)]
C10 = [f"col_{i}" for i in range(1, 11)]
C514 = [f"col_{i}" for i in range(5, 15)]


def test_one_row_missing_a_name():
    with pytest.raises(ValueError, match="Some rows missing names. Check row names."):
        prep.give_names(_named(R10[:9] + [None], R514, C10, C514), 2)


def test_one_view_rows_not_named():
    with pytest.raises(ValueError, match="At least one view is missing row names. Please name missing rows."):
        prep.give_names(_named(R10, None, C10, C514), 2)


def test_one_column_missing_a_name():
    with pytest.raises(ValueError, match="Some columns missing names. Check columns names."):
        prep.give_names(_named(R10, R514, C10[:9] + [None], C514), 2)


def test_one_view_columns_not_named():
    with pytest.raises(ValueError,
                       match="At least one view is missing column names. Please name missing columns."):
        prep.give_names(_named(R10, R514, C10, None), 2)


def test_no_row_names_present_names_added():
    res = prep.give_names(_named(None, None, C10, C514), 2)
    assert res["row_names"][0] == [f"row_{i}" for i in range(1, 11)]
    assert res["row_names"][1] == [f"row_{i}" for i in range(11, 21)]


def test_no_col_names_present_names_added():
    res = prep.give_names(_named(R10, R514, None, None), 2)
    assert res["col_names"][0] == [f"col_{i}" for i in range(1, 11)]
    assert res["col_names"][1] == [f"col_{i}" for i in range(11, 21)]


def test_all_named_no_change():
    res = prep.give_names(_named(R10, R514, C10, C514), 2)
    assert res["row_names"] == [R10, R514] and res["col_names"] == [C10, C514]


def test_negative_matrix_warns_and_is_shifted_per_column():
    x = np.array([[1.0, -2.0], [3.0, 4.0], [-1.0, 0.5]])
    with pytest.warns(UserWarning, match="Matrix is not non-negative. Has been made non-negative."):
        out = prep.make_non_neg_inner(x)
    assert np.array_equal(out, x + np.array([1.0, 2.0])[None, :])
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert np.array_equal(prep.make_non_neg_inner(np.abs(x)), np.abs(x))


def test_matrix_normalisation_unit_column_sums():
    x = np.abs(np.random.default_rng(0).standard_normal((7, 4))) + 0.1
    np.testing.assert_allclose(prep.matrix_normalisation(x).sum(0), np.ones(4), rtol=1e-15)


def test_init_rest_mats():
    assert np.array_equal(prep.init_rest_mats(None, 2), np.zeros((2, 2)))
    m = np.zeros((2, 2))
    m[0, 1] = 200.0
    m[1, 1] = 9.0
    assert np.array_equal(prep.init_rest_mats(m, 2), np.array([[0.0, 200.0], [200.0, 0.0]]))


def _check(**over):
    args = dict(data=two(), init_f=None, init_s=None, init_g=None, k_vec=None, phi=np.zeros((2, 2)),
                xi=np.zeros((2, 2)), psi=np.zeros((2, 2)), n_iters=None, k_min=3, k_max=8,
                distance="euclidean", num_repeats=5, no_clusts=False, sample_rate=0.9, n_stability=5,
                stability=True, stab_thres=0.4, remove_unstable=True, spurious=True)
    args.update(over)
    return prep.check_inputs(**args)


@pytest.mark.parametrize("over,msg", [
    (dict(k_min=8, k_max=8), "k_max must be greater than k_min."),
    (dict(n_iters=2.5), "n_iters  must be a positive integer."),
    (dict(num_repeats="a"), "num_repeats  must be a numeric."),
    (dict(stability="yes"), "stability must be a boolean."),
    (dict(stab_thres=1.5), "stab_thres must be between 0 and 1."),
    (dict(sample_rate=0.0), "sample_rate must be greater than 0 and less than or equal to 1."),
    (dict(distance="chebyshev"), "distance must be one of 'euclidean', 'manhattan' or 'cosine'."),
    (dict(phi=-np.ones((2, 2))), "phi  must be a non-negative matrix."),
    (dict(psi=np.zeros((3, 3))), "psi  must be of the same dimensions as data."),
    (dict(k_vec=[3]), "k_vec must be a vector of the same length as the number of views."),
    (dict(k_vec=[11, 3]), "k_vec must be a vector of integers less than or equal to the"),
    (dict(k_max=11), "k_max must be less than or equal to the minimum rank of the views."),
    (dict(init_f=[np.ones((10, 3))] * 2), "init_f must be a list of matrices or NULL."),
])
def test_check_inputs_error_strings(over, msg):
    with pytest.raises(ValueError, match=msg.replace("(", r"\(").replace(")", r"\)")):
        _check(**over)


def test_check_inputs_returns_prepped_views():
    out = _check()
    for m in out:
        np.testing.assert_allclose(m.x.sum(0), np.ones(10), rtol=1e-14)
        assert m.x.flags.f_contiguous and (m.x >= 0).all()


def test_reorder_data_and_maps_partial_overlap():
    rn = [["a", "b", "c", "d"], ["c", "x", "a"], ["q"]]
    cn = [["u", "v"], ["v", "u"], ["u"]]
    data = [NamedMatrix(np.zeros((len(r), len(c))), r, c) for r, c in zip(rn, cn)]
    ro = prep.reorder_data(data, 3, rn, cn)
    assert ro["row_indices"][0] == {1: ["a", "c"], 2: None}
    assert ro["row_indices"][1] == {0: ["c", "a"], 2: None}
    assert ro["col_indices"][2] == {0: ["u"], 1: ["u"]}
    maps = prep.shared_maps(ro["row_indices"], rn)
    iv, iw = maps[(0, 1)]
    assert list(iv) == [0, 2] and list(iw) == [2, 0]
    assert maps[(0, 2)][0].size == 0  # NA pair
    assert prep.shared_maps(None, rn) == {}  # R's NULL: nothing set


def test_reorder_matches_power_set_definition():
    """reorder_data's shortcut (names(v) & names(w)) equals the reference's power-set construction."""
    from itertools import combinations

    rng = np.random.default_rng(3)
    pool = [f"n{i}" for i in range(12)]
    names = [list(rng.permutation(pool)[: rng.integers(3, 10)]) for _ in range(4)]
    V = 4
    data = [NamedMatrix(np.zeros((len(r), 1)), r, ["c"]) for r in names]
    ro = prep.reorder_data(data, V, names, [["c"]] * V)["row_indices"]
    # reference construction: for every non-empty subset A, rows in all of A and in none of the others
    subsets, lists = [], []
    for size in range(1, V + 1):
        for a in combinations(range(V), size):
            inter = set(names[a[0]]).intersection(*[set(names[i]) for i in a[1:]])
            others = set().union(*[set(names[i]) for i in range(V) if i not in a]) if len(a) < V else set()
            rows = inter - others
            if rows:
                subsets.append(a)
                lists.append(rows)
    for v in range(V):
        for w in range(V):
            if v == w:
                continue
            common = set().union(*[l for a, l in zip(subsets, lists) if v in a and w in a] or [set()])
            assert (set(ro[v][w]) if ro[v][w] else set()) == common
