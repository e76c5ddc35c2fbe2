"""Host post-processing mirrors (R/obtain_bicl.r, R/stability_analysis.r): the reference's own unit tests
(test-resnmtf.R:1-35) plus checks of the restated helpers."""
import numpy as np

from resnmtf_b200 import bicluster as B
from resnmtf_b200 import stability as S


def test_test_cond_identifies_zero_rows_and_columns():
    """test-resnmtf.R:1-25."""
    assert not S.test_cond([np.arange(1, 10).reshape(3, 3)], attempt=2)
    assert S.test_cond([np.array([[1, 2, 3], [0, 0, 0], [7, 8, 9]])], attempt=2)
    assert S.test_cond([np.array([[1, 0, 3], [4, 0, 6], [7, 0, 9]])], attempt=2)
    assert S.test_cond([np.zeros((3, 3))], attempt=2)
    rng = np.random.default_rng(0)
    assert not S.test_cond([rng.uniform(1, 10, (10, 10))], attempt=2)
    assert S.test_cond([rng.uniform(1, 10, (10, 10))], attempt=1)


def test_view_shuffling_keeps_dims_and_multiset():
    """test-resnmtf.R:27-35."""
    rng = np.random.default_rng(1)
    data = [np.abs(rng.standard_normal((10, 10))) for _ in range(2)]
    out = [B.shuffle_view(x, rng) for x in data]
    assert len(out) == 2 and out[0].shape == (10, 10)
    assert np.allclose(np.sort(out[0].ravel()), np.sort(data[0].ravel()))


def test_jaccard_closed_form_equals_cartesian_product_sets():
    """|A x B n C x D| = |A n C| |B n D| replaces cart_prod/jaccard_func (R/utils.r:117-145)."""
    rng = np.random.default_rng(2)
    rc, cc = (rng.random((12, 3)) < 0.4).astype(float), (rng.random((9, 3)) < 0.4).astype(float)
    tr, tc = (rng.random((12, 2)) < 0.4).astype(float), (rng.random((9, 2)) < 0.4).astype(float)
    jac = S.jaccard_main(rc, cc, tr, tc, 3, 2)
    for i in range(3):
        a = {(r, c) for r in np.flatnonzero(rc[:, i]) for c in np.flatnonzero(cc[:, i])}
        for j in range(2):
            b = {(r, c) for r in np.flatnonzero(tr[:, j]) for c in np.flatnonzero(tc[:, j])}
            u = len(a | b)
            assert np.isclose(jac[i, j], 0.0 if u == 0 else len(a & b) / u)


def test_relevance_edge_cases():
    z = np.zeros((5, 2))
    o = np.ones((5, 2))
    assert S.relevance_results(z, z, o, o) == 0.0
    assert S.relevance_results(o, o, z, z) == 0.0
    assert S.relevance_results(z, z, z, z) == 1.0
    assert np.allclose(S.relevance_results(o, o, o, o), [1.0, 1.0])


def test_jsd_properties():
    p = np.array([0.2, 0.3, 0.5])
    assert abs(B.jsd(p, p)) < 1e-15
    assert np.isclose(B.jsd(np.array([1.0, 0.0]), np.array([0.0, 1.0])), 1.0)  # log2 units: max is 1
    rng = np.random.default_rng(3)
    a, b = rng.random(200), rng.random(200) + 2.0
    assert B.jsd_calc(a, a) < 1e-12 and B.jsd_calc(a, b) > 0.5


def test_density_integrates_to_one():
    x = np.random.default_rng(4).standard_normal(500)
    gx, gy = B.r_density(x)
    assert abs(np.trapezoid(gy, gx) - 1.0) < 2e-3
    assert abs(gx[np.argmax(gy)]) < 0.5


def test_bisilhouette_prefers_the_true_biclustering():
    rng = np.random.default_rng(5)
    n = 30
    rc = np.zeros((3 * n, 3))
    for i in range(3):
        rc[i * n:(i + 1) * n, i] = 1
    x = rc @ (10 * np.eye(3)) @ rc.T + 0.1 * np.abs(rng.standard_normal((3 * n, 3 * n)))
    good = B.bisilhouette(x, rc, rc)["bisil"]
    perm = rng.permutation(3 * n)
    bad = B.bisilhouette(x, rc[perm], rc)["bisil"]
    assert good > 0.8 and good > bad + 0.5
    assert B.bisilhouette(x, np.zeros_like(rc), np.zeros_like(rc))["bisil"] == 0.0


def test_binarise_thresholds_and_relations():
    f = np.array([[0.6, 0.1], [0.3, 0.5], [0.1, 0.4]])  # n = 3: threshold 1/3
    g = np.array([[0.5, 0.5], [0.5, 0.5]])              # p = 2: threshold 1/2 (strict)
    rows, cols = B.binarise([f], [g])
    assert np.array_equal(rows[0], [[1, 0], [0, 1], [0, 1]])
    assert np.array_equal(cols[0], np.zeros((2, 2)))
