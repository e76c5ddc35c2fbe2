"""The persistent single-CTA loop for tiny fits (rn_small_sweeps, RESNMTF_IMPL_SMALL; BASELINE configs[0]): all views
and all sweeps of a batch in ONE launch, against the oracle sweep by sweep -- every k, ragged shapes up to the size
limits, all three error modes, phi / psi / xi with partial overlaps, mixed k, the stop rule and the NaN failure."""
import numpy as np
import pytest

from golden_util import load
from helpers import RTOL, Problem, compare_trace, rel_err
from oracle import resnmtf_oracle as O
from resnmtf_b200 import _lib as L
from resnmtf_b200 import synth

pytestmark = pytest.mark.gpu


def single(n, p, k, seed):
    rng = np.random.default_rng(seed)
    x = synth.prep(synth.planted_view(n, p, min(3, n, p), rng, 0.4, 0.4)[0] + 0.01)
    f, s, g = synth.random_factors(n, p, k, rng)
    return Problem([x], [k], [f], [s], [g])


def assert_small(prob, ctx, n_iters, err_mode):
    worst = compare_trace(prob, ctx, n_iters=n_iters, err_mode=err_mode, impl=L.IMPL_SMALL)
    fit = prob.device_fit(ctx, err_mode=err_mode, impl=L.IMPL_SMALL)
    try:
        fit.run(2)
        c = fit.counters()
        assert c["impl"] == L.IMPL_SMALL and c["kernel_launches"] == 1
    finally:
        fit.close()
    assert worst <= RTOL
    return worst


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("err_mode", [L.ERR_AUTO, L.ERR_DIRECT])
def test_every_k(ctx, k, err_mode):
    assert_small(single(100, 50, k, seed=100 + k), ctx, 6, err_mode)


@pytest.mark.parametrize("shape", [(1, 1), (1, 40), (40, 1), (2, 3), (63, 33), (64, 32), (65, 31), (100, 30),
                                   (180, 180), (1000, 100), (128, 1000), (999, 97), (90, 1024)])
def test_shapes_up_to_the_limits(ctx, shape):
    n, p = shape
    assert_small(single(n, p, min(3, n, p), seed=7 * n + p), ctx, 4, L.ERR_DIRECT)


def test_just_above_the_limits_takes_the_streaming_kernels(ctx):
    """Asked for explicitly it serves views up to 1024 x 1024 padded and 1 MB; AUTO picks it up to 16384 padded entries."""
    for n, p, impl in ((1025, 60, L.IMPL_SMALL), (60, 1025, L.IMPL_SMALL), (1000, 200, L.IMPL_SMALL), (180, 180, L.IMPL_AUTO)):
        prob = single(n, p, 3, seed=n + p)
        fit = prob.device_fit(ctx, impl=impl)
        try:
            fit.run(2)
            assert fit.counters()["impl"] != L.IMPL_SMALL
        finally:
            fit.close()
    fit = single(100, 50, 3, seed=1).device_fit(ctx)  # the README toy's first view: AUTO -> the persistent loop
    try:
        fit.run(2)
        assert fit.counters()["impl"] == L.IMPL_SMALL
    finally:
        fit.close()


@pytest.mark.parametrize("err_mode", [L.ERR_AUTO, L.ERR_ALGEBRAIC, L.ERR_DIRECT])
def test_three_views_phi_psi_xi_partial_overlaps(ctx, err_mode):
    prob, z, V = load("three_views")
    assert_small(prob, ctx, 10, err_mode)


def test_mixed_k_and_na_pairs(ctx):
    rng = np.random.default_rng(3)
    shapes = [(90, 40), (70, 55), (90, 55)]
    ks = [3, 5, 3]
    data = [synth.prep(synth.planted_view(n, p, 3, rng, 0.4, 0.4)[0] + 0.01) for n, p in shapes]
    fs = [synth.random_factors(n, p, k, rng) for (n, p), k in zip(shapes, ks)]
    rn = [[f"r{i}" for i in range(90)], [f"q{i}" for i in range(70)], [f"r{i}" for i in rng.permutation(90)]]
    cn = [[f"c{i}" for i in range(40)], [f"d{i}" for i in range(55)], [f"d{i}" for i in range(55)]]
    phi = np.zeros((3, 3)); phi[0, 2] = 150.0; phi[0, 1] = 10.0   # views 0 and 1 share no row names: an NA pair
    psi = np.zeros((3, 3)); psi[0, 1] = 60.0                      # no shared column names either
    xi = np.zeros((3, 3)); xi[0, 2] = 30.0
    prob = Problem(data, ks, [f[0] for f in fs], [f[1] for f in fs], [f[2] for f in fs],
                   phi=O.init_rest_mats(phi, 3), xi=O.init_rest_mats(xi, 3), psi=O.init_rest_mats(psi, 3),
                   row_names=rn, col_names=cn)
    assert_small(prob, ctx, 6, L.ERR_DIRECT)


def test_converged_toy_same_stop_sweep_and_biclusters(ctx):
    """BASELINE configs[0] (README toy) run to convergence in batches of 32 sweeps per launch."""
    prob, z, V = load("readme_toy")
    fit = prob.device_fit(ctx)
    try:
        done = fit.run(None, 1.0e-6)
        c = fit.counters()
        assert c["impl"] == L.IMPL_SMALL and c["converged"] == 1
        assert done == len(z["all_error"]) and rel_err(fit.errors(), z["all_error"]) <= RTOL
        assert c["kernel_launches"] <= (done + 31) // 32 + 2  # + the AUTO hand-over's residual launches, if any
        fit.normalise()
        outs = [fit.get_factors(v) for v in range(V)]
        rows, cols, _ = O.binarise([o[0] for o in outs], [o[2] for o in outs], [o[1] for o in outs])
        for v in range(V):
            assert np.array_equal(rows[v].astype(np.uint8), z[f"rows_{v}"])
            assert np.array_equal(cols[v].astype(np.uint8), z[f"cols_{v}"])
    finally:
        fit.close()


def test_bitwise_repeatable_and_close_to_the_streaming_kernels(ctx, monkeypatch):
    prob, z, V = load("three_views")
    runs = []
    for i in range(3):
        fit = prob.device_fit(ctx, impl=L.IMPL_SMALL if i < 2 else L.IMPL_AUTO)
        fit.run(25)
        runs.append(([fit.get_factors(v) for v in range(V)], fit.errors().copy(), fit.counters()["impl"]))
        fit.close()
    assert runs[0][2] == L.IMPL_SMALL and runs[2][2] != L.IMPL_SMALL
    assert np.array_equal(runs[0][1], runs[1][1])
    for a, b in zip(runs[0][0], runs[1][0]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    assert rel_err(runs[0][1], runs[2][1]) <= 1e-8
    for a, b in zip(runs[0][0], runs[2][0]):
        assert rel_err(a[0], b[0]) <= 1e-8 and rel_err(a[2], b[2]) <= 1e-8


def test_nan_error_is_reported(ctx):
    views, _ = synth.block_views(2, block=20, n_blocks=3, seed=2)
    data = [synth.prep(x) for x in views]
    n = data[0].shape[0]
    rng = np.random.default_rng(4)
    fs = [synth.random_factors(n, n, 3, rng) for _ in range(2)]
    m = np.zeros((2, 2)); m[0, 1] = 10.0
    prob = Problem(data, [3, 3], [np.zeros_like(f[0]) for f in fs], [f[1] for f in fs], [np.zeros_like(f[2]) for f in fs],
                   phi=O.init_rest_mats(m, 2), psi=O.init_rest_mats(m, 2))
    fit = prob.device_fit(ctx, impl=L.IMPL_SMALL)
    try:
        with pytest.raises(L.ResnmtfNaNError):
            fit.run(None, 1.0e-6, 50)
        assert fit.counters()["impl"] == L.IMPL_SMALL
    finally:
        fit.close()
