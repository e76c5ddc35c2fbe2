"""Pins the CPU oracle against everything the reference's own tests hold for the hot path
(tests/testthat/test-resnmtf.R:63-184).  The reference has no seeds and no golden numbers, so these are
the property-level checks its test-suite makes, re-expressed on the oracle: cluster sizes {60,60,60} on the
3-block data, colSums(F) = colSums(G) = 1, reconstruction column sums, and phi/psi pulling shared rows /
columns together more than unshared ones."""
import numpy as np
import pytest

from oracle import resnmtf_oracle as O
from resnmtf_b200 import synth


def block_problem(seed):
    views, rc = synth.block_views(2, seed=seed)
    data = [O.matrix_normalisation(O.make_non_neg(x)[0]) for x in views]
    return data, rc


def fit(data, k, phi=None, psi=None, xi=None, rn=None, cn=None, seed=0):
    n_v = len(data)
    z = np.zeros((n_v, n_v))
    dn_r, dn_c = O.default_names(data)
    rn = dn_r if rn is None else rn
    cn = dn_c if cn is None else cn
    rng = np.random.default_rng(seed)
    noise = [np.abs(np.sqrt(0.05) * rng.standard_normal((k, k))) for _ in range(n_v)]
    res = O.res_nmtf_loop(data, O.shared_names(rn), O.shared_names(cn), rn, cn, None, None, None, [k] * n_v,
                          z if phi is None else phi, z if xi is None else xi, z if psi is None else psi,
                          noise=noise, max_iters=2000)
    rows, cols, _ = O.binarise(res["output_f"], res["output_g"], res["output_s"])
    return res, rows, cols


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fixed_k_cluster_sizes_and_normalisation(seed):
    """test-resnmtf.R:98-118 (no stability, no spurious removal)."""
    data, _ = block_problem(seed)
    res, rows, cols = fit(data, 3, seed=seed)
    assert len(res["output_f"]) == 2 and res["output_f"][0].shape == (180, 3)
    for v in range(2):
        assert sorted(rows[v].sum(0)) == [60, 60, 60]
        assert sorted(cols[v].sum(0)) == [60, 60, 60]
    np.testing.assert_allclose(res["output_f"][0].sum(0), np.ones(3), rtol=0, atol=1e-12)
    np.testing.assert_allclose(res["output_g"][0].sum(0), np.ones(3), rtol=0, atol=1e-12)
    recon = res["output_f"][0] @ res["output_s"][0] @ res["output_g"][0].T
    assert np.mean(recon.sum(0) - 1.0) < 1e-3


def test_phi_psi_partial_overlap_pulls_shared_names_together():
    """test-resnmtf.R:140-184: rows/cols 1..120 shared, 121..180 / 181..240 not; phi = psi = 1000 e_12."""
    data, _ = block_problem(4)
    rn = [[f"row_{i}" for i in range(1, 181)],
          [f"row_{i}" for i in range(1, 121)] + [f"row_{i}" for i in range(181, 241)]]
    cn = [[f"col_{i}" for i in range(1, 181)],
          [f"col_{i}" for i in range(1, 121)] + [f"col_{i}" for i in range(181, 241)]]
    rest = np.zeros((2, 2))
    rest[0, 1] = 1000.0
    phi = O.init_rest_mats(rest, 2)
    psi = O.init_rest_mats(rest, 2)
    res, rows, cols = fit(data, 3, phi=phi, psi=psi, rn=rn, cn=cn, seed=5)
    f, g = res["output_f"], res["output_g"]
    assert np.mean(np.abs(f[0][120:180] - f[1][120:180])) > np.mean(np.abs(f[0][:120] - f[1][:120]))
    assert np.mean(np.abs(g[0][120:180] - g[1][120:180])) > np.mean(np.abs(g[0][:120] - g[1][:120]))
    for v in range(2):
        assert sorted(rows[v].sum(0)) == [60, 60, 60]
        assert sorted(cols[v].sum(0)) == [60, 60, 60]


def test_init_rest_mats_semantics():
    """R/update_steps.r:12-24: NULL -> zeros; diagonal dropped; a symmetric input is doubled."""
    assert np.array_equal(O.init_rest_mats(None, 3), np.zeros((3, 3)))
    m = np.array([[5.0, 2.0], [0.0, 7.0]])
    assert np.array_equal(O.init_rest_mats(m, 2), np.array([[0.0, 2.0], [2.0, 0.0]]))
    s = np.array([[0.0, 3.0], [3.0, 0.0]])
    assert np.array_equal(O.init_rest_mats(s, 2), 2 * s)


def test_star_prod_relevant_weighting_and_na():
    """R/utils.r:63-78: (1/n_v) sum_w phi_w n_w M_w, NA pairs skipped, unshared rows keep the own factor."""
    rng = np.random.default_rng(0)
    f0, f1, f2 = rng.random((4, 2)), rng.random((6, 2)), rng.random((5, 2))
    names = [["a", "b", "c", "d"], ["c", "x", "a", "y", "z", "w"], ["p", "q", "r", "s", "t"]]
    idx = O.shared_names(names)[0]
    assert idx[2] is None and idx[1] == ["a", "c"]
    out = O.star_prod_relevant(np.array([0.0, 2.0, 3.0]), [f0, f1, f2], f0, idx, names[0], names)
    masked = f0.copy()
    masked[0] = f1[2]
    masked[2] = f1[0]
    np.testing.assert_allclose(out, 2.0 * masked * 6 / 4, rtol=1e-15)


def test_nan_ratio_becomes_one_only_on_uncoupled_branch():
    """R/update_steps.r:152-163: 0/0 -> ratio 1 without phi; NaN survives with phi (quirk Q6)."""
    x = np.zeros((3, 2))
    f = [np.zeros((3, 1)), np.zeros((3, 1))]
    s, g, lam = np.ones((1, 1)), np.zeros((2, 1)), np.zeros(1)
    names = [["a", "b", "c"]] * 2
    idx = O.shared_names(names)
    out = O.update_f(x, f, s, g, lam, np.zeros((2, 2)), 0, idx[0], names)
    assert np.array_equal(out, np.zeros((3, 1)))
    phi = np.array([[0.0, 1.0], [1.0, 0.0]])
    out = O.update_f(x, f, s, g, lam, phi, 0, idx[0], names)
    assert np.isnan(out).all()


def test_convergence_rule_first_comparison_is_against_zero():
    """R/main.r:53-81: err_temp starts at 0; the loop stops at the first sweep whose mean error moved by
    <= 1e-6, and Error is the mean of the last 10 entries of All_Error (R/main.r:126-127)."""
    data, _ = block_problem(7)
    res, _, _ = fit(data, 3, seed=7)
    errs = np.concatenate([[0.0], res["All_Error"]])
    diffs = np.abs(np.diff(errs))
    assert (diffs[:-1] > 1e-6).all() and diffs[-1] <= 1e-6
    assert np.isclose(res["Error"], np.mean(res["All_Error"][-10:]), rtol=1e-15)
