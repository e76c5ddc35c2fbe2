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


def test_bw_nrd0_of_all_columns_at_once_equals_the_per_column_function():
    from resnmtf_b200 import bicluster as B

    rng = np.random.default_rng(3)
    cols = np.abs(rng.standard_normal((5000, 7))) ** 3
    cols[:, 4] = 0.25            # constant column: sd = IQR = 0 -> |x[1]|
    cols[:, 5] = 0.0             # all zero -> 1
    cols[2500:, 6] = 0.0         # IQR = 0 but sd > 0 -> sd
    want = np.array([B.bw_nrd0(cols[:, c]) for c in range(7)])
    assert np.array_equal(B.bw_nrd0_columns(cols), want)


def test_draw_subsample_drops_empty_rows_and_columns_and_reuses_view_one_sample():
    """The sampling loop of stability_repeat (R/stability_analysis.r:222-253) through the sums-only interface: rows and
    columns that are all zero inside the sub-sample are dropped, views with view 1's dimensions reuse its sample, a
    view of other dimensions draws its own, and the final sub-samples have no all-zero row or column."""
    from resnmtf_b200.prep import NamedMatrix

    rng = np.random.default_rng(5)
    a = rng.random((60, 40)) + 0.1
    a[7, :] = 0.0                      # an all-zero row: dropped whenever it is drawn
    a[:, 11] = 0.0                     # an all-zero column
    b = a.copy()
    c = rng.random((50, 30)) + 0.1     # other dimensions: its own sample
    data = [NamedMatrix(x, [f"r{i}" for i in range(x.shape[0])], [f"c{j}" for j in range(x.shape[1])])
            for x in (a, b, c)]
    shapes = [m.shape for m in data]
    out = S.draw_subsample(S.host_sums(data), shapes, shapes[0], 3, 0.9, np.random.default_rng(1))
    assert out is not None
    rows, cols = out
    assert np.array_equal(rows[0], rows[1]) and np.array_equal(cols[0], cols[1])
    assert 7 not in rows[0] and 11 not in cols[0]
    assert len(rows[2]) == 45 and len(cols[2]) == 27
    for v in range(3):
        sub = data[v].x[np.ix_(rows[v], cols[v])]
        assert not (sub.sum(0) == 0).any() and not (sub.sum(1) == 0).any()


def test_resident_matrix_has_shape_and_names_but_no_host_values():
    from resnmtf_b200.fitpool import ResidentMatrix
    from resnmtf_b200.prep import as_named

    m = ResidentMatrix((5, 3), ["a", "b", "c", "d", "e"], ["x", "y", "z"])
    assert m.shape == (5, 3) and m.x is None and m.rownames[4] == "e"
    cp = as_named(m)
    assert isinstance(cp, ResidentMatrix) and cp.shape == (5, 3) and cp.colnames == ["x", "y", "z"]


def test_split_bisilhouette_combines_pieces_in_bicluster_order():
    """native_route._SplitBisil deals the biclusters of one fit over the GPUs of the pool and sums the per-bicluster
    values in bicluster order: with a host stand-in for the per-GPU entry point the result equals the whole score."""
    import threading

    import numpy as np

    from resnmtf_b200 import bicluster as B
    from resnmtf_b200.native_route import _SplitBisil

    rng = np.random.default_rng(3)
    n, p, k = 120, 40, 5
    x = np.abs(rng.standard_normal((n, p)))
    rc = (rng.random((n, k)) < 0.3).astype(float)
    cc = (rng.random((p, k)) < 0.4).astype(float)
    rc[:, 1] = 0.0
    calls = []

    class Handle:
        def __init__(self, g):
            self.g = g

        def bisil_part(self, rcm, ccm, want, method):
            live = [j for j in range(k) if rcm[:, j].any() and ccm[:, j].any()]
            full = B.bisilhouette(x, rcm, ccm, method=method)["vals"]
            vals = np.zeros(k)
            for j, v in zip(live, full):
                if want[j]:
                    vals[j] = v
            calls.append((self.g, [j for j in range(k) if want[j]]))
            return vals, len(live)

    class Runner:
        pool = [0, 1, 2]
        gpu_locks = [threading.Lock() for _ in range(3)]
        bisil_lock = threading.Lock()
        bisil_load = [0.0, 0.0, 0.0]  # shared by the fits of one batch (a second fit would start from the first one's loads)

        def handle(self, view, g):
            return Handle(g)

    whole = B.bisilhouette(x, rc, cc)
    got = _SplitBisil(Runner(), 0, first_gpu=1).bisil(rc, cc)
    assert got["vals"] == whole["vals"] and abs(got["bisil"] - whole["bisil"]) <= 1e-15
    assert sorted(j for _, js in calls for j in js) == [0, 2, 3, 4]  # every live bicluster exactly once
    assert len({g for g, _ in calls}) == 3                            # over all three "GPUs"
    first = list(Runner.bisil_load)
    assert min(first) > 0.0
    _SplitBisil(Runner(), 0, first_gpu=1).bisil(rc, cc)                # a second fit adds to the same table
    assert all(b > a for a, b in zip(first, Runner.bisil_load))
