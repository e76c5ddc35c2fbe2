"""The BASELINE.json configurations as parity cases: C2 at full size (20000 x 4000), C3 / C4 with the exact
restriction structure of SURVEY 8(d) at reduced size (the oracle materialises X_hat, so full size would not
fit the time budget), plus size-independent properties at full size (two independently written kernel paths
agree; unit column sums after normalisation_check; the algebraic and the direct error agree)."""
import numpy as np
import pytest

from helpers import RTOL, Problem, compare_trace, rel_err
from oracle import resnmtf_oracle as O
from resnmtf_b200 import _lib as L
from resnmtf_b200 import synth

pytestmark = pytest.mark.gpu


def planted(n, p, n_planted, seed, rows=None, cols=None, sigma=1.0):
    rng = np.random.default_rng(seed)
    return synth.planted_view(n, p, n_planted, rng, 0.2, 0.2, 5.0, sigma, rows=rows, cols=cols)


def test_c2_full_size_two_sweeps_vs_oracle_and_paths_agree(ctx):
    n, p, k = 20000, 4000, 5
    x, _, _ = planted(n, p, 5, synth.config_seed(2, 0))
    x = synth.prep(x)
    f, s, g = synth.random_factors(n, p, k, np.random.default_rng(1))
    prob = Problem([x], [k], [f], [s], [g])
    compare_trace(prob, ctx, n_iters=2, err_mode=L.ERR_AUTO, impl=L.IMPL_FUSED)  # the default path at this size
    compare_trace(prob, ctx, n_iters=2, err_mode=L.ERR_AUTO, impl=L.IMPL_TMA)
    outs = {}
    for impl, mode in ((L.IMPL_TMA, L.ERR_ALGEBRAIC), (L.IMPL_FUSED, L.ERR_ALGEBRAIC), (L.IMPL_DMMA, L.ERR_DIRECT),
                       (L.IMPL_DFMA, L.ERR_DIRECT)):
        fit = prob.device_fit(ctx, err_mode=mode, impl=impl)
        fit.run(5)
        errs = fit.errors()
        assert fit.counters()["impl"] == impl
        fit.normalise()
        outs[impl] = (fit.get_factors(0), errs)
        fit.close()
    ref = outs[L.IMPL_TMA]
    for impl in (L.IMPL_FUSED, L.IMPL_DMMA, L.IMPL_DFMA):
        for a, b in zip(outs[impl][0][:3], ref[0][:3]):
            assert rel_err(a, b) <= RTOL
        assert rel_err(outs[impl][1], ref[1]) <= RTOL  # direct vs algebraic error
    np.testing.assert_allclose(ref[0][0].sum(0), np.ones(k), rtol=0, atol=1e-12)
    np.testing.assert_allclose(ref[0][2].sum(0), np.ones(k), rtol=0, atol=1e-12)


@pytest.fixture(params=["two-pass kernels", "fused kernel"])
def kernel_family(request, monkeypatch):
    """The reduced-size views take the one-pass kernel by default (rn_fused2_step deals their columns without padding;
    the full-size configs run rn_fused_step); the other pass of each structure test sends them through the two-pass
    TMA kernels, which serve the views the one-pass kernels do not take (p > 8064, row-sharded, k > 8)."""
    if request.param == "two-pass kernels":
        monkeypatch.setenv("RESNMTF_IMPL", str(L.IMPL_TMA))
    return request.param


def assert_family(prob, ctx, family):
    fit = prob.device_fit(ctx, err_mode=L.ERR_ALGEBRAIC)
    try:
        fit.run(1)
        assert fit.counters()["impl"] == (L.IMPL_FUSED if family == "fused kernel" else L.IMPL_TMA)
    finally:
        fit.close()


def test_c3_structure_four_views_phi_psi(ctx, kernel_family):
    """4 views; 1-2 and 3-4 share rows (phi = 200), 1&3 and 2&4 share columns (psi = 200); k = 5."""
    n, p, k = 12500, 1250, 5
    _, r12, _ = planted(8, 8, 5, 1)  # placeholders replaced below
    rng = np.random.default_rng(synth.config_seed(3, 0))
    rows_a = (rng.random((n, 5)) < 0.2).astype(float)
    rows_b = (rng.random((n, 5)) < 0.2).astype(float)
    cols_a = (rng.random((p, 5)) < 0.2).astype(float)
    cols_b = (rng.random((p, 5)) < 0.2).astype(float)
    layout = [(rows_a, cols_a), (rows_a, cols_b), (rows_b, cols_a), (rows_b, cols_b)]
    data = [synth.prep(planted(n, p, 5, synth.config_seed(3, v), rows=r, cols=c)[0]) for v, (r, c) in enumerate(layout)]
    rn = [[f"a{i}" for i in range(n)], [f"a{i}" for i in range(n)], [f"b{i}" for i in range(n)],
          [f"b{i}" for i in range(n)]]
    cn = [[f"u{i}" for i in range(p)], [f"w{i}" for i in range(p)], [f"u{i}" for i in range(p)],
          [f"w{i}" for i in range(p)]]
    phi = np.zeros((4, 4)); phi[0, 1] = 200.0; phi[2, 3] = 200.0
    psi = np.zeros((4, 4)); psi[0, 2] = 200.0; psi[1, 3] = 200.0
    irng = np.random.default_rng(7)
    inits = [synth.random_factors(n, p, k, irng) for _ in range(4)]
    prob = Problem(data, [k] * 4, [i[0] for i in inits], [i[1] for i in inits], [i[2] for i in inits],
                   phi=O.init_rest_mats(phi, 4), psi=O.init_rest_mats(psi, 4), row_names=rn, col_names=cn)
    assert_family(prob, ctx, kernel_family)
    compare_trace(prob, ctx, n_iters=4, err_mode=L.ERR_AUTO)


def test_c4_structure_eight_views_phi_xi_psi(ctx, kernel_family):
    """8 views, all sharing rows and columns, phi = psi = 200 and xi = 50 on every pair; k = 8."""
    n, p, k, V = 6250, 500, 8, 8
    rng = np.random.default_rng(synth.config_seed(4, 0))
    rows = (rng.random((n, 8)) < 0.2).astype(float)
    cols = (rng.random((p, 8)) < 0.2).astype(float)
    data = [synth.prep(planted(n, p, 8, synth.config_seed(4, v), rows=rows, cols=cols)[0]) for v in range(V)]
    rn = [[f"r{i}" for i in range(n)]] * V
    cn = [[f"c{i}" for i in range(p)]] * V
    up = np.triu(np.ones((V, V)), 1)
    irng = np.random.default_rng(8)
    inits = [synth.random_factors(n, p, k, irng) for _ in range(V)]
    prob = Problem(data, [k] * V, [i[0] for i in inits], [i[1] for i in inits], [i[2] for i in inits],
                   phi=O.init_rest_mats(200.0 * up, V), xi=O.init_rest_mats(50.0 * up, V),
                   psi=O.init_rest_mats(200.0 * up, V), row_names=rn, col_names=cn)
    assert_family(prob, ctx, kernel_family)
    compare_trace(prob, ctx, n_iters=3, err_mode=L.ERR_AUTO)


# ---- BASELINE configs[2] / configs[3] at FULL size ------------------------------------------------------------------
# One sweep against the oracle (which materialises X_hat per view: seconds per sweep at this size), then the kernel
# families against each other over several sweeps, as for C2.


def _cross_paths(prob, ctx, sweeps, impls):
    outs = {}
    for impl in impls:
        fit = prob.device_fit(ctx, err_mode=L.ERR_ALGEBRAIC, impl=impl)
        try:
            fit.run(sweeps)
            assert fit.counters()["impl"] == impl
            errs = fit.errors()
            fit.normalise()
            outs[impl] = ([fit.get_factors(v)[:3] for v in range(len(prob.data))], errs)
        finally:
            fit.close()
    ref = outs[impls[0]]
    for impl in impls[1:]:
        for fa, fb in zip(outs[impl][0], ref[0]):
            for a, b in zip(fa, fb):
                assert rel_err(a, b) <= RTOL
        assert rel_err(outs[impl][1], ref[1]) <= RTOL
    return ref


def test_c3_full_size_vs_oracle_and_paths_agree(ctx):
    """configs[2]: 4 views 50000 x 5000, phi on (1,2), (3,4), psi on (1,3), (2,4), k = 5 -- 8 GB of data, fused kernel on
    5-CTA clusters."""
    n, p, k = 50000, 5000, 5
    rng = np.random.default_rng(synth.config_seed(3, 0))
    rows_a = (rng.random((n, 5)) < 0.2).astype(float)
    rows_b = (rng.random((n, 5)) < 0.2).astype(float)
    cols_a = (rng.random((p, 5)) < 0.2).astype(float)
    cols_b = (rng.random((p, 5)) < 0.2).astype(float)
    layout = [(rows_a, cols_a), (rows_a, cols_b), (rows_b, cols_a), (rows_b, cols_b)]
    data = [synth.prep(planted(n, p, 5, synth.config_seed(3, v), rows=r, cols=c)[0]) for v, (r, c) in enumerate(layout)]
    rn = [[f"a{i}" for i in range(n)]] * 2 + [[f"b{i}" for i in range(n)]] * 2
    cn = [[f"u{i}" for i in range(p)], [f"w{i}" for i in range(p)]] * 2
    phi = np.zeros((4, 4)); phi[0, 1] = 200.0; phi[2, 3] = 200.0
    psi = np.zeros((4, 4)); psi[0, 2] = 200.0; psi[1, 3] = 200.0
    irng = np.random.default_rng(7)
    inits = [synth.random_factors(n, p, k, irng) for _ in range(4)]
    prob = Problem(data, [k] * 4, [i[0] for i in inits], [i[1] for i in inits], [i[2] for i in inits],
                   phi=O.init_rest_mats(phi, 4), psi=O.init_rest_mats(psi, 4), row_names=rn, col_names=cn)
    compare_trace(prob, ctx, n_iters=1, err_mode=L.ERR_AUTO)  # default path = fused
    ref = _cross_paths(prob, ctx, 4, (L.IMPL_FUSED, L.IMPL_TMA))
    for fv in ref[0]:
        np.testing.assert_allclose(fv[0].sum(0), np.ones(k), rtol=0, atol=1e-12)
        np.testing.assert_allclose(fv[2].sum(0), np.ones(k), rtol=0, atol=1e-12)


def test_c4_full_size_vs_oracle_and_paths_agree(ctx):
    """configs[3]: 8 views 100000 x 2000, all sharing rows and columns, phi = psi = 200 and xi = 50 on every pair, k = 8
    -- 12.8 GB of data, 7 phi partners per view, fused kernel on 2-CTA clusters."""
    n, p, k, V = 100000, 2000, 8, 8
    rng = np.random.default_rng(synth.config_seed(4, 0))
    rows = (rng.random((n, 8)) < 0.2).astype(float)
    cols = (rng.random((p, 8)) < 0.2).astype(float)
    data = [synth.prep(planted(n, p, 8, synth.config_seed(4, v), rows=rows, cols=cols)[0]) for v in range(V)]
    rn = [[f"r{i}" for i in range(n)]] * V
    cn = [[f"c{i}" for i in range(p)]] * V
    up = np.triu(np.ones((V, V)), 1)
    irng = np.random.default_rng(8)
    inits = [synth.random_factors(n, p, k, irng) for _ in range(V)]
    prob = Problem(data, [k] * V, [i[0] for i in inits], [i[1] for i in inits], [i[2] for i in inits],
                   phi=O.init_rest_mats(200.0 * up, V), xi=O.init_rest_mats(50.0 * up, V),
                   psi=O.init_rest_mats(200.0 * up, V), row_names=rn, col_names=cn)
    compare_trace(prob, ctx, n_iters=1, err_mode=L.ERR_AUTO)
    _cross_paths(prob, ctx, 3, (L.IMPL_FUSED, L.IMPL_TMA))


def test_c5_shard_width_on_the_sharded_path_vs_oracle():
    """configs[4] shard geometry: a 32768 x 20000 row shard (p = 20000 > 8064: two-pass TMA kernels, 313 column groups,
    5.2 GB) through the row-sharded code path (communicator of one rank: pack -> ncclAllReduce -> stand-alone G
    epilogue), one sweep against the oracle and the sharded path against the unsharded one over three sweeps."""
    from resnmtf_b200.device import Context, DeviceFit

    n, p, k = 32768, 20000, 8
    x = synth.prep(planted(n, p, 6, synth.config_seed(5, 0))[0])
    f, s, g = synth.random_factors(n, p, k, np.random.default_rng(3))
    prob = Problem([x], [k], [f], [s], [g])
    states = []
    prob.oracle(n_iters=1, trace=lambda t, cf, cs, cg, cl, cm, err: states.append((cf[0], cs[0], cg[0], cl[0], cm[0], err)))
    outs = []
    for sharded in (True, False):
        with Context(0) as c:
            if sharded:
                c.join(Context.comm_id_create(), 0, 1)
            fit = DeviceFit(c, [n], [p], [k])
            fit.set_options(err_mode=L.ERR_AUTO, impl=L.IMPL_AUTO)
            fit.set_data(0, x)
            fit.set_factors(0, f, s, g)
            fit.step()
            got = fit.get_factors(0)
            errs, _ = fit.view_errors()
            for a, b in zip(got + (errs,), states[0]):
                assert rel_err(a, np.asarray(b).reshape(np.asarray(a).shape)) <= RTOL
            fit.run(2)
            assert fit.counters()["impl"] == L.IMPL_TMA
            outs.append(fit.get_factors(0)[:3] + (fit.errors(),))
            fit.close()
    for a, b in zip(outs[0], outs[1]):
        assert rel_err(a, b) <= RTOL
