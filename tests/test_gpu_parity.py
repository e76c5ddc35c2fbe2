"""Parity of the CUDA path (through the C ABI) against the CPU oracle, sweep by sweep.
Run on the B200 box:  python -m pytest tests -m gpu -x -q"""
import os

import numpy as np
import pytest

from helpers import RTOL, Problem, compare_trace, rel_err
from oracle import resnmtf_oracle as O
from resnmtf_b200 import _lib as L
from resnmtf_b200 import synth

pytestmark = pytest.mark.gpu


def single_view_problem(n, p, k, seed, n_planted=4, sigma=1.0):
    rng = np.random.default_rng(seed)
    x, _, _ = synth.planted_view(n, p, n_planted, rng, row_prob=0.3, col_prob=0.3, sigma=sigma)
    x = synth.prep(x)
    f, s, g = synth.random_factors(n, p, k, rng)
    return Problem([x], [k], [f], [s], [g])


@pytest.mark.parametrize("impl", [L.IMPL_DFMA, L.IMPL_DMMA, L.IMPL_TMA])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8])
def test_single_view_every_k(ctx, k, impl):
    prob = single_view_problem(300, 200, k, seed=100 + k)
    compare_trace(prob, ctx, n_iters=6, err_mode=L.ERR_DIRECT, impl=impl)


@pytest.mark.parametrize("impl", [L.IMPL_AUTO, L.IMPL_DFMA, L.IMPL_TMA])
@pytest.mark.parametrize("k", [9, 10, 11, 12, 13, 14, 15, 16])
def test_single_view_k_above_8(ctx, k, impl):
    """k = 9..16 (the k-extension loop, R/main.r:306-320): the CUDA-core kernels and the two-tile tensor-core TMA pair
    (what AUTO picks) against the oracle."""
    prob = single_view_problem(200, 150, k, seed=200 + k, n_planted=6)
    compare_trace(prob, ctx, n_iters=4, err_mode=L.ERR_DIRECT, impl=impl)
    if impl == L.IMPL_AUTO:
        fit = prob.device_fit(ctx)
        try:
            fit.run(1, 0.0)
            assert fit.counters()["impl"] == L.IMPL_TMA
        finally:
            fit.close()


@pytest.mark.parametrize("split", [(0, 0), (7, 5), (23, 41), (1, 1)])
@pytest.mark.parametrize("k", [9, 16])
def test_k_above_8_tma_straddling_units(ctx, k, split, monkeypatch):
    """Two-tile TMA kernels with row tiles / column groups that straddle CTAs (odd persistent grids, one CTA), on a shape
    with ragged rows and columns."""
    if split[0]:
        monkeypatch.setenv("RESNMTF_F_CTAS", str(split[0]))
        monkeypatch.setenv("RESNMTF_G_CTAS", str(split[1]))
    prob = single_view_problem(1300, 710, k, seed=400 + k, n_planted=8)
    compare_trace(prob, ctx, n_iters=4, err_mode=L.ERR_DIRECT, impl=L.IMPL_TMA)


def test_k_above_8_tma_equals_cuda_core_path_and_is_repeatable(ctx):
    """k = 12 on a mid-size view: the TMA pair against the independently written CUDA-core kernels (1e-9), bit-identical
    run to run, and the algebraic error of the fused finish against the direct residual."""
    prob = single_view_problem(3000, 1100, 12, seed=77, n_planted=10)
    outs = []
    for impl in (L.IMPL_TMA, L.IMPL_TMA, L.IMPL_DFMA):
        fit = prob.device_fit(ctx, err_mode=L.ERR_ALGEBRAIC, impl=impl)
        fit.run(12, 0.0)
        outs.append((fit.get_factors(0), fit.errors()))
        fit.close()
    for a, b in zip(outs[0][0], outs[1][0]):
        assert np.array_equal(a, b)
    assert np.array_equal(outs[0][1], outs[1][1])
    for a, b in zip(outs[0][0], outs[2][0]):
        assert rel_err(a, b) <= RTOL
    assert rel_err(outs[0][1], outs[2][1]) <= RTOL


@pytest.mark.parametrize("impl", [L.IMPL_DFMA, L.IMPL_DMMA, L.IMPL_TMA])
@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (63, 7), (64, 8), (65, 9), (129, 65), (257, 93), (1000, 37)])
def test_ragged_shapes(ctx, shape, impl):
    n, p = shape
    k = min(3, n, p)
    prob = single_view_problem(n, p, k, seed=300 + n + p, n_planted=2)
    compare_trace(prob, ctx, n_iters=4, err_mode=L.ERR_DIRECT, impl=impl)


@pytest.mark.parametrize("err_mode", [L.ERR_AUTO, L.ERR_ALGEBRAIC, L.ERR_DIRECT])
def test_error_modes(ctx, err_mode):
    prob = single_view_problem(500, 260, 4, seed=42)
    compare_trace(prob, ctx, n_iters=8, err_mode=err_mode)


@pytest.mark.parametrize("impl", [L.IMPL_DFMA, L.IMPL_DMMA, L.IMPL_TMA])
@pytest.mark.parametrize("split", [(3, 5, 7, 5), (1, 1, 1, 1), (2, 7, 23, 41), (4, 2, 1000000, 1000000)])
def test_forced_splits(ctx, impl, split, monkeypatch):
    """Work partitions that make tiles / column groups straddle CTAs in every way: column-split F step and
    row-split G stream on the CUDA-core path, odd persistent grid sizes on the stream-K path (the
    cross-CTA partial + last-arriver combine)."""
    monkeypatch.setenv("RESNMTF_F_CS", str(split[0]))
    monkeypatch.setenv("RESNMTF_G_RS", str(split[1]))
    monkeypatch.setenv("RESNMTF_F_CTAS", str(split[2]))
    monkeypatch.setenv("RESNMTF_G_CTAS", str(split[3]))
    prob = single_view_problem(1500, 700, 5, seed=7)
    compare_trace(prob, ctx, n_iters=5, err_mode=L.ERR_DIRECT, impl=impl)


def two_view_problem(seed, k=3, phi=0.0, psi=0.0, xi=0.0, partial=False):
    views, _ = synth.block_views(2, block=40, n_blocks=3, seed=seed)
    n = views[0].shape[0]
    data = [synth.prep(x) for x in views]
    rng = np.random.default_rng(seed + 1)
    inits = [synth.random_factors(n, n, k, rng) for _ in range(2)]
    rn = cn = None
    if partial:
        m = 2 * n // 3
        rn = [[f"row_{i}" for i in range(1, n + 1)],
              [f"row_{i}" for i in range(1, m + 1)] + [f"row_{i}" for i in range(n + 1, 2 * n - m + 1)]]
        cn = [[f"col_{i}" for i in range(1, n + 1)],
              [f"col_{i}" for i in range(1, m + 1)] + [f"col_{i}" for i in range(n + 1, 2 * n - m + 1)]]
        # scramble view 2's order so the gather map is a real permutation
        perm = rng.permutation(n)
        data[1] = np.asfortranarray(data[1][perm][:, perm])
        rn[1] = [rn[1][i] for i in perm]
        cn[1] = [cn[1][i] for i in perm]
    else:
        rn = [[f"row_{i}" for i in range(1, n + 1)]] * 2
        cn = [[f"col_{i}" for i in range(1, n + 1)]] * 2

    def rest(val):
        m_ = np.zeros((2, 2))
        m_[0, 1] = val
        return O.init_rest_mats(m_, 2)

    return Problem(data, [k, k], [i[0] for i in inits], [i[1] for i in inits], [i[2] for i in inits],
                   phi=rest(phi), xi=rest(xi), psi=rest(psi), row_names=rn, col_names=cn)


@pytest.mark.parametrize("impl", [L.IMPL_DFMA, L.IMPL_DMMA, L.IMPL_TMA])
@pytest.mark.parametrize("cfg", [
    dict(phi=200.0), dict(psi=200.0), dict(xi=50.0), dict(phi=200.0, psi=100.0, xi=50.0),
    dict(phi=1000.0, psi=1000.0, partial=True), dict(phi=5.0, partial=True),
])
def test_two_views_coupled(ctx, cfg, impl):
    prob = two_view_problem(seed=11, **cfg)
    compare_trace(prob, ctx, n_iters=6, err_mode=L.ERR_DIRECT, impl=impl)


def test_four_views_mixed_k_and_na_pairs(ctx):
    """Unequal shapes; phi/psi between equal-k views, one pair sharing nothing (R's NA -> skipped), and an
    uncoupled view with another k that still takes the coupled G branch (whole-matrix sum(psi) test)."""
    rng = np.random.default_rng(5)
    shapes = [(120, 80), (120, 50), (90, 80), (70, 40)]
    ks = [3, 3, 3, 5]
    data = [synth.prep(synth.planted_view(n, p, 3, rng, 0.3, 0.3)[0]) for n, p in shapes]
    inits = [synth.random_factors(n, p, k, rng) for (n, p), k in zip(shapes, ks)]
    rn = [[f"r{i}" for i in range(120)], [f"r{i}" for i in range(120)], [f"q{i}" for i in range(90)],
          [f"z{i}" for i in range(70)]]
    cn = [[f"c{i}" for i in range(80)], [f"d{i}" for i in range(50)], [f"c{i}" for i in range(80)],
          [f"y{i}" for i in range(40)]]
    phi = np.zeros((4, 4)); phi[0, 1] = 30.0; phi[0, 2] = 7.0   # (0,2) share no row names -> NA
    psi = np.zeros((4, 4)); psi[0, 2] = 40.0
    prob = Problem(data, ks, [i[0] for i in inits], [i[1] for i in inits], [i[2] for i in inits],
                   phi=O.init_rest_mats(phi, 4), psi=O.init_rest_mats(psi, 4), row_names=rn, col_names=cn)
    for impl in (L.IMPL_DFMA, L.IMPL_DMMA, L.IMPL_TMA):
        compare_trace(prob, ctx, n_iters=6, err_mode=L.ERR_DIRECT, impl=impl)


def test_coupling_views_of_different_k_is_rejected(ctx):
    """phi between views with different k is non-conformable in the reference (R/utils.r:72); E_INVALID."""
    rng = np.random.default_rng(6)
    data = [synth.prep(synth.planted_view(50, 30, 2, rng, 0.3, 0.3)[0]) for _ in range(2)]
    inits = [synth.random_factors(50, 30, k, rng) for k in (2, 3)]
    phi = np.zeros((2, 2)); phi[0, 1] = 5.0
    prob = Problem(data, [2, 3], [i[0] for i in inits], [i[1] for i in inits], [i[2] for i in inits],
                   phi=O.init_rest_mats(phi, 2), row_names=[[f"r{i}" for i in range(50)]] * 2)
    fit = prob.device_fit(ctx)
    try:
        with pytest.raises(L.ResnmtfError) as e:
            fit.run(1)
        assert e.value.code == L.E_INVALID
    finally:
        fit.close()


def test_convergence_matches_oracle(ctx):
    """Stop rule of R/main.r:55: same number of sweeps, same All_Error, identical binarised output."""
    views, _ = synth.block_views(2, seed=3)
    data = [synth.prep(x) for x in views]
    rng = np.random.default_rng(9)
    k = 3
    noise = [np.abs(np.sqrt(0.05) * rng.standard_normal((k, k))) for _ in range(2)]
    fs, ss, gs, _, _ = O.init_mats_inner(data, [k, k], noise)
    prob = Problem(data, [k, k], fs, ss, gs)
    ref = prob.oracle()
    fit = prob.device_fit(ctx)
    try:
        done = fit.run(None, 1.0e-6)
        errs = fit.errors()
        assert done == len(ref["All_Error"]) == len(errs)
        assert rel_err(errs, ref["All_Error"]) <= RTOL
        fit.normalise()
        outs = [fit.get_factors(v) for v in range(2)]
        for v in range(2):
            assert rel_err(outs[v][0], ref["output_f"][v]) <= RTOL
            assert rel_err(outs[v][1], ref["output_s"][v]) <= RTOL
            assert rel_err(outs[v][2], ref["output_g"][v]) <= RTOL
        rows_d, cols_d, _ = O.binarise([o[0] for o in outs], [o[2] for o in outs], [o[1] for o in outs])
        rows_o, cols_o, _ = O.binarise(ref["output_f"], ref["output_g"], ref["output_s"])
        for v in range(2):
            assert np.array_equal(rows_d[v], rows_o[v]) and np.array_equal(cols_d[v], cols_o[v])
            assert sorted(rows_d[v].sum(0)) == [60, 60, 60] and sorted(cols_d[v].sum(0)) == [60, 60, 60]
        c = fit.counters()
        assert c["converged"] == 1 and c["kernel_launches"] > 0
    finally:
        fit.close()


def test_medium_single_view_bitwise_repeatable(ctx):
    """4000 x 1500, k=8: parity with the oracle and run-to-run bit-reproducibility (fixed-order sums)."""
    prob = single_view_problem(4000, 1500, 8, seed=77, n_planted=5)
    compare_trace(prob, ctx, n_iters=3, err_mode=L.ERR_AUTO)
    outs = []
    for _ in range(2):
        fit = prob.device_fit(ctx)
        fit.run(20)
        outs.append((fit.get_factors(0), fit.errors()))
        fit.close()
    for a, b in zip(outs[0][0], outs[1][0]):
        assert np.array_equal(a, b)
    assert np.array_equal(outs[0][1], outs[1][1])


def test_nan_error_is_reported(ctx):
    """An all-zero G makes the coupled update 0/0: the reference's while(NA) throws; we return E_NAN."""
    prob = two_view_problem(seed=2, phi=10.0, psi=10.0)
    prob.init_g = [np.zeros_like(g) for g in prob.init_g]
    prob.init_f = [np.zeros_like(f) for f in prob.init_f]
    with pytest.raises(FloatingPointError):
        prob.oracle(max_iters=5)
    fit = prob.device_fit(ctx)
    try:
        with pytest.raises(L.ResnmtfNaNError):
            fit.run(None, 1.0e-6, max_iters=5)
    finally:
        fit.close()


def test_state_errors(ctx):
    from resnmtf_b200.device import DeviceFit

    fit = DeviceFit(ctx, [10], [5], [2])
    with pytest.raises(L.ResnmtfError) as e:
        fit.run(1)
    assert e.value.code == L.E_STATE
    fit.close()
    with pytest.raises(L.ResnmtfError) as e:
        DeviceFit(ctx, [10], [5], [17])
    assert e.value.code == L.E_UNSUPPORTED


def test_shared_data_handle_matches_private_copy(ctx):
    """resnmtf_data: one upload shared by the fits of a k-sweep gives bit-identical results to set_data, and
    survives being destroyed while fits are still attached (reference counting)."""
    from resnmtf_b200.device import DeviceData, DeviceFit

    prob = single_view_problem(700, 333, 4, seed=55)
    x = prob.data[0]
    data = DeviceData(ctx, x)
    outs = []
    for shared in (False, True, True):
        fit = DeviceFit(ctx, [x.shape[0]], [x.shape[1]], [4])
        if shared:
            fit.attach_data(0, data)
        else:
            fit.set_data(0, x)
        fit.set_factors(0, prob.init_f[0], prob.init_s[0], prob.init_g[0])
        if len(outs) == 1:
            data.close()  # the second fit keeps the buffer alive on its own
        fit.run(12)
        outs.append((fit.get_factors(0), fit.errors(), fit.view_errors()[1]))
        if len(outs) == 2:
            data = DeviceData(ctx, x)  # a fresh handle for the third fit
        fit.close()
    for o in outs[1:]:
        for a, b in zip(outs[0][0], o[0]):
            assert np.array_equal(a, b)
        assert np.array_equal(outs[0][1], o[1]) and np.array_equal(outs[0][2], o[2])
    data.close()


def clean_block_problem(sigma, n=1500, p=600, k=3, seed=11):
    """Three disjoint blocks of height 10 (the reference's test data, test-resnmtf.R:38-52) with little noise, factors
    started near the blocks: the error falls through 1e-3 within a few sweeps."""
    rng = np.random.default_rng(seed)
    x = sigma * np.abs(rng.standard_normal((n, p)))
    for b in range(3):
        x[b * n // 3:(b + 1) * n // 3, b * p // 3:(b + 1) * p // 3] += 10.0
    x = synth.prep(x)
    f = 0.1 * np.abs(rng.standard_normal((n, k)))
    g = 0.1 * np.abs(rng.standard_normal((p, k)))
    for b in range(3):
        f[b * n // 3:(b + 1) * n // 3, b] += 1.0
        g[b * p // 3:(b + 1) * p // 3, b] += 1.0
    f /= f.sum(0)
    g /= g.sum(0)
    s = np.eye(k) + 0.05 * np.abs(rng.standard_normal((k, k)))
    return Problem([x], [k], [f], [s], [g])


def test_auto_error_mode_stays_on_the_one_pass_path_above_1e_4(ctx):
    """AUTO error mode on clean data: between 1e-3 and 1e-4 the algebraic error still holds the 1e-9 bar (deviation
    ~1e-15 / err), so no direct residual pass runs and the fit keeps one launch and one read of X per sweep; the whole
    error history and the stop sweep are the oracle's."""
    prob = clean_block_problem(sigma=0.1)
    ref = prob.oracle()
    assert 1.0e-4 < ref["All_Error"][-1] < 1.0e-3
    fit = prob.device_fit(ctx, err_mode=L.ERR_AUTO)
    try:
        done = fit.run(None, 1.0e-6)
        assert done == len(ref["All_Error"])
        assert rel_err(fit.errors(), ref["All_Error"]) <= RTOL
        c = fit.counters()
        assert c["direct_error_passes"] == 0 and c["impl"] == L.IMPL_FUSED
        assert c["kernel_launches"] <= 32  # one launch per enqueued sweep (one batch of 32), none for the error
    finally:
        fit.close()


def test_auto_error_mode_hands_over_below_1e_4(ctx):
    """Cleaner data: the error drops below 1e-4, where cancellation would cost the algebraic form its ninth digit; the
    direct residual pass takes over from that sweep on and the history still matches to 1e-9."""
    prob = clean_block_problem(sigma=0.01)
    ref = prob.oracle()
    assert ref["All_Error"][-1] < 1.0e-5
    fit = prob.device_fit(ctx, err_mode=L.ERR_AUTO)
    try:
        done = fit.run(None, 1.0e-6)
        assert done == len(ref["All_Error"])
        assert rel_err(fit.errors(), ref["All_Error"]) <= RTOL
        below = int((np.asarray(ref["All_Error"]) < 1.0e-4).sum())
        assert fit.counters()["direct_error_passes"] >= 1 and below >= 1
    finally:
        fit.close()
