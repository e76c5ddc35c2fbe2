"""Data-handle operations and the SVD initialisation behind the C ABI (rn_native.cu): prep, download, sums, shuffles,
sub-samples, device-to-device copies, top singular triplets -- against NumPy / LAPACK on the same inputs."""
import numpy as np
import pytest

from resnmtf_b200 import _lib as L
from resnmtf_b200 import synth
from resnmtf_b200.device import DeviceData

pytestmark = pytest.mark.gpu


def raw_view(n, p, seed, negative=True):
    rng = np.random.default_rng(seed)
    x, _, _ = synth.planted_view(n, p, 3, rng, 0.3, 0.3, 5.0, 1.0)
    if negative:
        x[:, ::3] -= 0.7  # some columns dip below zero: the shift of make_non_neg_inner is per column
    return np.asfortranarray(x)


@pytest.mark.parametrize("n,p", [(100, 50), (257, 131), (1000, 300), (64, 32)])
def test_prep_on_the_device_equals_the_reference_prep(ctx, n, p):
    x = raw_view(n, p, 1)
    d, neg = DeviceData.prepped(ctx, x)
    assert neg is True
    got = d.download()
    np.testing.assert_allclose(got, synth.prep(x), rtol=1e-14, atol=0)
    cs, rs = d.sums()
    np.testing.assert_allclose(cs, np.ones(p), rtol=0, atol=1e-13)
    np.testing.assert_allclose(rs, got.sum(axis=1), rtol=1e-13, atol=0)
    d.close()
    d2, neg2 = DeviceData.prepped(ctx, synth.prep(x))
    assert neg2 is False
    d2.close()


def test_upload_download_round_trip_is_exact(ctx):
    x = raw_view(333, 77, 2)
    d = DeviceData(ctx, x)
    assert np.array_equal(d.download(), x)
    d.close()


def test_subsample_is_the_index_gather(ctx):
    x = synth.prep(raw_view(500, 200, 3))
    rng = np.random.default_rng(4)
    rows = rng.permutation(500)[:450]
    cols = rng.permutation(200)[:180]
    d = DeviceData(ctx, x)
    sub = d.subsample(rows, cols)
    assert sub.shape == (450, 180)
    assert np.array_equal(sub.download(), x[np.ix_(rows, cols)])
    cs, rs = sub.sums()
    np.testing.assert_allclose(cs, x[np.ix_(rows, cols)].sum(0), rtol=1e-13)
    np.testing.assert_allclose(rs, x[np.ix_(rows, cols)].sum(1), rtol=1e-13)
    sub.close()
    d.close()


@pytest.mark.parametrize("n,p", [(180, 180), (1000, 300), (129, 65)])
def test_shuffle_permutes_all_entries_reproducibly(ctx, n, p):
    x = synth.prep(raw_view(n, p, 5, negative=False))
    d = DeviceData(ctx, x)
    a = d.shuffle(seed=11, renormalise=False)
    b = d.shuffle(seed=11, renormalise=False)
    c = d.shuffle(seed=12, renormalise=False)
    xa, xb, xc = a.download(), b.download(), c.download()
    assert np.array_equal(xa, xb)            # a function of the seed only
    assert not np.array_equal(xa, xc)
    assert np.array_equal(np.sort(xa.ravel()), np.sort(x.ravel()))  # a permutation of ALL entries
    assert (xa != x).mean() > 0.9            # ... that moves almost every entry
    assert (xa.sum(0) != 0).all() and (xa.sum(1) != 0).all()
    # the planted structure is gone: entries land anywhere (position correlation ~ 0)
    assert abs(np.corrcoef(xa.ravel(), x.ravel())[0, 1]) < 0.05
    r = d.shuffle(seed=11, renormalise=True)
    xr = r.download()
    np.testing.assert_allclose(xr, xa / xa.sum(0)[None, :], rtol=1e-14)
    for h in (a, b, c, r, d):
        h.close()


def test_shuffle_rejects_draws_with_an_all_zero_line(ctx):
    """A sparse view: most shuffles leave some row or column all zero; the result never does (R/obtain_bicl.r:13-18)."""
    rng = np.random.default_rng(6)
    x = np.zeros((12, 10), order="F")
    idx = rng.permutation(120)[:70]
    x.ravel(order="F")[idx] = rng.random(70) + 0.1
    d = DeviceData(ctx, x)
    tries = []
    for seed in range(20):
        s = d.shuffle(seed=seed, renormalise=False)
        xs = s.download()
        assert (xs.sum(0) != 0).all() and (xs.sum(1) != 0).all()
        assert np.array_equal(np.sort(xs.ravel()), np.sort(x.ravel()))
        tries.append(s.attempts)
        s.close()
    assert max(tries) > 1
    d.close()


def _check_triplets(x, u, dv, v, k, tol):
    uu, ss, vt = np.linalg.svd(x, full_matrices=False)
    np.testing.assert_allclose(dv, ss[:k], rtol=1e-11, atol=1e-13 * ss[0])
    for c in range(k):
        gap = min(abs(ss[c] - ss[j]) for j in range(len(ss)) if j != c) / ss[0]
        if gap < 1e-6:
            continue  # vectors of (nearly) repeated singular values are not unique
        np.testing.assert_allclose(u[:, c], np.abs(uu[:, c]), rtol=0, atol=tol / max(gap, 1e-3) * 1e-3)
        np.testing.assert_allclose(v[:, c], np.abs(vt[c]), rtol=0, atol=tol / max(gap, 1e-3) * 1e-3)


@pytest.mark.parametrize("n,p,k", [(40, 30, 3), (300, 120, 5), (1000, 300, 8), (200, 500, 6), (2000, 900, 16)])
def test_top_singular_triplets_match_lapack(ctx, n, p, k):
    """|U_k|, d_k, |V_k| from the device (Gram matrix of the smaller side -- both orientations --, subspace iteration or
    the dense route for tiny sides) against LAPACK's full svd of the view, planted spectrum."""
    x = synth.prep(raw_view(n, p, 7, negative=False))
    d = DeviceData(ctx, x)
    u, dv, v = d.svd_topk(k)
    _check_triplets(x, u, dv, v, k, 1e-10)
    u2, dv2, v2 = d.svd_topk(min(k, 3))  # cached: a narrower request is a slice of the same numbers
    assert np.array_equal(u2, u[:, :min(k, 3)]) and np.array_equal(dv2, dv[:min(k, 3)])
    d.close()


def test_top_singular_triplets_of_shuffled_data_match_lapack(ctx):
    """The nearly degenerate spectrum of a shuffled view (one dominant triplet, the rest in the noise bulk)."""
    x = synth.prep(raw_view(1500, 600, 8, negative=False))
    d = DeviceData(ctx, x)
    s = d.shuffle(seed=3)
    xs = s.download()
    u, dv, v = s.svd_topk(6)
    _check_triplets(xs, u, dv, v, 6, 1e-9)
    s.close()
    d.close()


def test_device_to_device_copy_carries_the_view_and_its_triplets(ctx):
    if L.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from resnmtf_b200.device import Context

    x = synth.prep(raw_view(700, 260, 9, negative=False))
    d = DeviceData(ctx, x)
    u, dv, v = d.svd_topk(4)
    with Context(1) as other:
        c = d.copy_to(other)
        assert np.array_equal(c.download(), x)
        u2, dv2, v2 = c.svd_topk(4)
        assert np.array_equal(u2, u) and np.array_equal(dv2, dv) and np.array_equal(v2, v)
        c.close()
    d.close()


@pytest.mark.parametrize("method", ["euclidean", "manhattan", "cosine"])
@pytest.mark.parametrize("n,p,k", [(400, 150, 5), (1000, 333, 8), (130, 70, 3)])
def test_bisilhouette_on_the_resident_view_matches_the_host_restatement(ctx, method, n, p, k):
    """resnmtf_data_bisil (distance blocks of bisilhouette on the device) against the host restatement of
    the same definition (scipy's direct-difference distances): overlapping clusters, an empty cluster, a single-row
    cluster, all three distances, ragged sizes (tiles of 64 rows / 16 columns are padded)."""
    from resnmtf_b200 import bicluster as B

    rng = np.random.default_rng(23)
    x = synth.prep(synth.planted_view(n, p, 3, rng, row_prob=0.3, col_prob=0.3)[0])
    rc = (rng.random((n, k)) < 0.25).astype(float)
    cc = (rng.random((p, k)) < 0.3).astype(float)
    if k >= 5:
        rc[:, 3] = 0.0            # empty row cluster
        rc[:, 4] = 0.0
        rc[17, 4] = 1.0           # single-row cluster
    d = DeviceData(ctx, x)
    host = B.bisilhouette(x, rc, cc, method=method)
    dev = d.bisil(rc, cc, method=method)
    assert len(dev["vals"]) == len(host["vals"])
    np.testing.assert_allclose(dev["vals"], host["vals"], rtol=1e-10, atol=1e-12)
    assert abs(dev["bisil"] - host["bisil"]) <= 1e-12
    again = d.bisil(rc, cc, method=method)
    assert again == dev           # fixed summation order: bit-identical
    d.close()


def test_bisilhouette_single_cluster_scores_against_the_rows_outside(ctx):
    """One non-empty bicluster only: the comparison group is the complement of its rows; all rows in it: score 0."""
    from resnmtf_b200 import bicluster as B

    rng = np.random.default_rng(5)
    x = synth.prep(synth.planted_view(300, 90, 3, rng, row_prob=0.3, col_prob=0.3)[0])
    rc = np.zeros((300, 3))
    cc = np.zeros((90, 3))
    rc[:120, 1] = 1.0
    cc[10:50, 1] = 1.0
    d = DeviceData(ctx, x)
    host = B.bisilhouette(x, rc, cc)
    dev = d.bisil(rc, cc)
    np.testing.assert_allclose(dev["vals"], host["vals"], rtol=1e-10, atol=1e-12)
    rc[:, 1] = 1.0
    assert d.bisil(rc, cc)["bisil"] == 0.0 == B.bisilhouette(x, rc, cc)["bisil"]
    assert d.bisil(np.zeros((300, 3)), cc)["bisil"] == 0.0
    d.close()


def test_bisilhouette_of_a_subset_of_the_biclusters_combines_to_the_whole(ctx):
    """resnmtf_data_bisil_part: the per-bicluster values are independent, so the biclusters of one fit can be scored in
    pieces (on different GPUs holding copies of the view) and summed in bicluster order -- bit-identical to the one
    call; this is what the native route does with the sweep fits of a multi-GPU call (native_route._SplitBisil)."""
    rng = np.random.default_rng(29)
    n, p, k = 500, 130, 6
    x = synth.prep(synth.planted_view(n, p, 3, rng, row_prob=0.3, col_prob=0.3)[0])
    rc = np.asfortranarray((rng.random((n, k)) < 0.25).astype(float))
    cc = np.asfortranarray((rng.random((p, k)) < 0.3).astype(float))
    rc[:, 2] = 0.0  # an empty bicluster does not count towards the mean
    d = DeviceData(ctx, x)
    whole = d.bisil(rc, cc)
    live = [j for j in range(k) if rc[:, j].any() and cc[:, j].any()]
    vals = np.zeros(k)
    for piece in ([0, 5], [1], [3, 4], [2]):
        want = np.zeros(k, dtype=np.int32)
        want[piece] = 1
        part, n_live = d.bisil_part(rc, cc, want)
        assert n_live == len(live)
        assert all(part[j] == 0.0 for j in range(k) if j not in piece)
        vals += part
    total = 0.0
    for j in live:
        total += float(vals[j])
    assert [float(vals[j]) for j in live] == whole["vals"]
    assert total / len(live) == whole["bisil"]
    d.close()
