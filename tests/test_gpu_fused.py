"""Parity of the one-pass fused kernels (RESNMTF_IMPL_FUSED: one read of X per update-iteration, 8-row groups resident
in the shared memory of a CTA cluster) against the CPU oracle, sweep by sweep, through the C ABI.  Every test runs three
times: with the library's own choice between the two kernel generations, with rn_fused_step forced (rn_fused.cuh, 1008
columns per CTA) and with rn_fused2_step forced (rn_fused2.cuh, <= 672 columns per CTA, no column padding).
Run on the B200 box:  python -m pytest tests -m gpu -x -q"""
import numpy as np
import pytest

from helpers import RTOL, Problem, compare_trace, rel_err
from resnmtf_b200 import _lib as L
from resnmtf_b200 import synth
from test_gpu_parity import single_view_problem, two_view_problem

pytestmark = pytest.mark.gpu

KIND2_MAX_COLS = 8 * 42 * 16  # widest view rn_fused2_step takes (8-CTA cluster, 42 blocks of 16 columns per CTA)


@pytest.fixture(autouse=True, params=["auto", "kind1", "kind2"])
def kernel_generation(request, monkeypatch):
    if request.param == "auto":
        monkeypatch.delenv("RESNMTF_FUSED_KIND", raising=False)
    else:
        monkeypatch.setenv("RESNMTF_FUSED_KIND", request.param[-1])
    return request.param


@pytest.fixture
def any_width(monkeypatch):
    """Lets narrow views (p << 1008) take the fused path too: the kernel pads the columns to the cluster width."""
    monkeypatch.setenv("RESNMTF_FUSED_MAX_PAD", "100000000")


def assert_fused(fit):
    assert fit.counters()["impl"] == L.IMPL_FUSED, "the view fell back to the two-pass kernels"


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8])
def test_fused_every_k(ctx, k, any_width):
    prob = single_view_problem(300, 200, k, seed=100 + k)
    compare_trace(prob, ctx, n_iters=6, err_mode=L.ERR_DIRECT, impl=L.IMPL_FUSED)


@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (7, 5), (8, 8), (9, 9), (63, 7), (64, 8), (65, 9), (129, 65),
                                   (257, 93), (1000, 37)])
def test_fused_ragged_shapes(ctx, shape, any_width):
    n, p = shape
    k = min(3, n, p)
    prob = single_view_problem(n, p, k, seed=300 + n + p, n_planted=2)
    compare_trace(prob, ctx, n_iters=4, err_mode=L.ERR_DIRECT, impl=L.IMPL_FUSED)


@pytest.mark.parametrize("shape,k", [((100, 1000), 3), ((700, 1008), 8), ((257, 1900), 5), ((90, 2016), 4),
                                     ((333, 4000), 5), ((1200, 3300), 8), ((40, 4032), 2)])
def test_fused_cluster_widths(ctx, shape, k):
    """p up to 1008 runs on single CTAs, up to 2016 on CTA pairs, ... up to 4032 on 4-CTA clusters (default padding rule)."""
    n, p = shape
    prob = single_view_problem(n, p, k, seed=500 + n + p, n_planted=4)
    fit = prob.device_fit(ctx, err_mode=L.ERR_ALGEBRAIC, impl=L.IMPL_FUSED)
    try:
        fit.run(1)
        assert_fused(fit)
    finally:
        fit.close()
    compare_trace(prob, ctx, n_iters=4, err_mode=L.ERR_DIRECT, impl=L.IMPL_FUSED)


def test_fused_wide_view_falls_back(ctx):
    """p > 8064 does not fit an 8-CTA cluster: the view runs the two-pass TMA kernels and says so."""
    prob = single_view_problem(64, 8100, 3, seed=9)
    fit = prob.device_fit(ctx, err_mode=L.ERR_ALGEBRAIC, impl=L.IMPL_FUSED)
    try:
        fit.run(2)
        assert fit.counters()["impl"] == L.IMPL_TMA
    finally:
        fit.close()
    compare_trace(prob, ctx, n_iters=2, err_mode=L.ERR_DIRECT, impl=L.IMPL_FUSED)


@pytest.mark.parametrize("err_mode", [L.ERR_AUTO, L.ERR_ALGEBRAIC, L.ERR_DIRECT])
def test_fused_error_modes(ctx, err_mode, any_width):
    prob = single_view_problem(500, 260, 4, seed=42)
    compare_trace(prob, ctx, n_iters=8, err_mode=err_mode, impl=L.IMPL_FUSED)


@pytest.mark.parametrize("clusters", [1, 2, 3, 7, 1000])
def test_fused_forced_cluster_counts(ctx, clusters, monkeypatch):
    """Any number of clusters (row groups split unevenly, clusters with a single group) gives the same factors."""
    monkeypatch.setenv("RESNMTF_FU_CLUSTERS", str(clusters))
    prob = single_view_problem(1500, 1800, 5, seed=7)
    compare_trace(prob, ctx, n_iters=5, err_mode=L.ERR_DIRECT, impl=L.IMPL_FUSED)


@pytest.mark.parametrize("cfg", [
    dict(phi=200.0), dict(psi=200.0), dict(xi=50.0), dict(phi=200.0, psi=100.0, xi=50.0),
    dict(phi=1000.0, psi=1000.0, partial=True), dict(phi=5.0, partial=True),
])
def test_fused_two_views_coupled(ctx, cfg, any_width):
    prob = two_view_problem(seed=11, **cfg)
    compare_trace(prob, ctx, n_iters=6, err_mode=L.ERR_DIRECT, impl=L.IMPL_FUSED)


def test_fused_mixed_with_two_pass_views(ctx):
    """A fit whose views do not all qualify: view 1 (p = 1000) runs fused, view 2 (p = 300) the TMA kernels;
    the phi coupling between them crosses the two kernel families."""
    rng = np.random.default_rng(21)
    shapes = [(400, 1000), (400, 300)]
    k = 4
    data = [synth.prep(synth.planted_view(n, p, 3, rng, 0.3, 0.3)[0]) for n, p in shapes]
    inits = [synth.random_factors(n, p, k, rng) for n, p in shapes]
    rn = [[f"r{i}" for i in range(400)]] * 2
    cn = [[f"c{i}" for i in range(1000)], [f"d{i}" for i in range(300)]]
    phi = np.zeros((2, 2))
    phi[0, 1] = phi[1, 0] = 150.0
    prob = Problem(data, [k, k], [i[0] for i in inits], [i[1] for i in inits], [i[2] for i in inits], phi=phi,
                   row_names=rn, col_names=cn)
    fit = prob.device_fit(ctx, err_mode=L.ERR_ALGEBRAIC, impl=L.IMPL_FUSED)
    try:
        fit.run(1)
        assert_fused(fit)
    finally:
        fit.close()
    compare_trace(prob, ctx, n_iters=5, err_mode=L.ERR_DIRECT, impl=L.IMPL_FUSED)


def test_fused_bitwise_repeatable_and_matches_two_pass(ctx):
    """Fixed summation orders: two runs are bit-identical; the two-pass kernels agree within the parity bar."""
    prob = single_view_problem(3000, 2000, 6, seed=77)
    outs = []
    for impl in (L.IMPL_FUSED, L.IMPL_FUSED, L.IMPL_TMA):
        fit = prob.device_fit(ctx, err_mode=L.ERR_ALGEBRAIC, impl=impl)
        try:
            fit.run(25)
            outs.append(fit.get_factors(0) + (np.asarray(fit.errors()),))
        finally:
            fit.close()
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    for a, b in zip(outs[0], outs[2]):
        assert rel_err(a, b) <= RTOL


def test_fused_convergence_rule(ctx, any_width):
    """The device-side stop rule (R/main.r:55-81) fires on the oracle's sweep."""
    prob = single_view_problem(400, 300, 3, seed=5)
    ref = prob.oracle(max_iters=400)
    fit = prob.device_fit(ctx, err_mode=L.ERR_AUTO, impl=L.IMPL_FUSED)
    try:
        done = fit.run(None, 1.0e-6, max_iters=400)
        assert_fused(fit)
        assert done == len(ref["All_Error"])
        f, s, g, _, _ = fit.get_factors(0)
        assert rel_err(f, ref["raw_f"][0]) <= RTOL
        assert rel_err(g, ref["raw_g"][0]) <= RTOL
        assert rel_err(s, ref["raw_s"][0]) <= RTOL
        assert rel_err(fit.errors(), ref["All_Error"]) <= 1e-7
    finally:
        fit.close()


@pytest.mark.parametrize("p,k", [(2100, 3), (5000, 5), (5900, 4), (7000, 8), (8064, 6)])
def test_fused_odd_cluster_sizes(ctx, p, k, monkeypatch, kernel_generation):
    """Clusters of 3, 5, 6, 7 and 8 CTAs (p up to 8 x 1008 columns)."""
    if kernel_generation == "kind2" and p > KIND2_MAX_COLS:
        pytest.skip("wider than rn_fused2_step's 8 x 672 columns")
    monkeypatch.setenv("RESNMTF_FUSED_MAX_PAD", "100000000")
    monkeypatch.setenv("RESNMTF_FUSED_MIN_SM_PCT", "1")
    prob = single_view_problem(150, p, k, seed=900 + p, n_planted=4)
    fit = prob.device_fit(ctx, err_mode=L.ERR_ALGEBRAIC, impl=L.IMPL_FUSED)
    try:
        fit.run(1)
        assert_fused(fit)
    finally:
        fit.close()
    compare_trace(prob, ctx, n_iters=3, err_mode=L.ERR_DIRECT, impl=L.IMPL_FUSED)


def test_fused_many_phi_partners(ctx, any_width):
    """Five views, every pair phi-coupled with partial, permuted row overlaps (4 partners per view: both lanes-per-row
    passes of the map prefetch), psi and xi on top; many row groups so the prefetch pipeline runs in steady state."""
    rng = np.random.default_rng(31)
    V, n, k = 5, 700, 3
    ps = [90, 120, 70, 100, 80]
    data = [synth.prep(synth.planted_view(n, p, 3, rng, 0.3, 0.3)[0]) for p in ps]
    inits = [synth.random_factors(n, p, k, rng) for p in ps]
    rn = []
    for v in range(V):  # view v names 70 % of a common pool, in its own order
        pool = rng.permutation(900)[:n]
        rn.append([f"r{j}" for j in pool])
    cn = [[f"c{v}_{j}" for j in range(p)] for v, p in enumerate(ps)]
    cn[1] = [f"c0_{j}" for j in range(90)] + [f"x{j}" for j in range(30)]  # views 0 and 1 share 90 columns
    phi = np.zeros((V, V))
    phi[np.triu_indices(V, 1)] = rng.uniform(5.0, 300.0, size=V * (V - 1) // 2)
    psi = np.zeros((V, V))
    psi[0, 1] = 40.0
    xi = np.zeros((V, V))
    xi[np.triu_indices(V, 1)] = 3.0
    from oracle import resnmtf_oracle as O

    prob = Problem(data, [k] * V, [i[0] for i in inits], [i[1] for i in inits], [i[2] for i in inits],
                   phi=O.init_rest_mats(phi, V), psi=O.init_rest_mats(psi, V), xi=O.init_rest_mats(xi, V),
                   row_names=rn, col_names=cn)
    fit = prob.device_fit(ctx, err_mode=L.ERR_ALGEBRAIC, impl=L.IMPL_FUSED)
    try:
        fit.run(1)
        assert_fused(fit)
    finally:
        fit.close()
    compare_trace(prob, ctx, n_iters=4, err_mode=L.ERR_DIRECT, impl=L.IMPL_FUSED)
