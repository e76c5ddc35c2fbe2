"""CUDA path against the committed golden fixtures: 10 sweeps from the SVD initialisation (1e-9 relative),
then the converged run: same number of sweeps, same All_Error, bit-identical binarised biclusters."""
import numpy as np
import pytest

from golden_util import CASES, N_SWEEPS, load
from helpers import RTOL, rel_err
from oracle import resnmtf_oracle as O
from resnmtf_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("impl", [L.IMPL_DFMA, L.IMPL_DMMA, L.IMPL_TMA, L.IMPL_FUSED, L.IMPL_SMALL])
@pytest.mark.parametrize("name", CASES)
def test_fixed_sweeps_match_golden(ctx, name, impl, monkeypatch):
    if impl == L.IMPL_FUSED:  # the golden views are narrow: lift the fused kernel's padding rule
        monkeypatch.setenv("RESNMTF_FUSED_MAX_PAD", "100000000")
    prob, z, V = load(name)
    fit = prob.device_fit(ctx, err_mode=L.ERR_AUTO, impl=impl)
    try:
        for t in range(N_SWEEPS):
            fit.step()
            errs, _ = fit.view_errors()
            assert rel_err(errs, z["sweep_errors"][t]) <= RTOL, f"sweep {t}"
        for v in range(V):
            f, s, g, lam, mu = fit.get_factors(v)
            assert rel_err(f, z[f"f{N_SWEEPS}_{v}"]) <= RTOL
            assert rel_err(s, z[f"s{N_SWEEPS}_{v}"]) <= RTOL
            assert rel_err(g, z[f"g{N_SWEEPS}_{v}"]) <= RTOL
            assert rel_err(lam, z[f"lam{N_SWEEPS}_{v}"]) <= RTOL
            assert rel_err(mu, z[f"mu{N_SWEEPS}_{v}"]) <= RTOL
    finally:
        fit.close()


@pytest.mark.parametrize("family", ["default kernels", "two-pass kernels", "first one-pass kernel",
                                    "second one-pass kernel"])
@pytest.mark.parametrize("name", CASES)
def test_converged_run_matches_golden(ctx, name, family, monkeypatch):
    """Default for these tiny fits: the persistent single-CTA loop (rn_small_sweeps) up to 16384 padded entries, else
    rn_fused2_step; then the two-pass TMA kernels, rn_fused_step and rn_fused2_step forced."""
    want = L.IMPL_FUSED if name == "three_views" else L.IMPL_SMALL
    if family == "two-pass kernels":
        monkeypatch.setenv("RESNMTF_IMPL", str(L.IMPL_TMA))
        want = L.IMPL_TMA
    elif family == "first one-pass kernel":
        monkeypatch.setenv("RESNMTF_SMALL", "0")
        monkeypatch.setenv("RESNMTF_FUSED_MAX_PAD", "100000000")
        monkeypatch.setenv("RESNMTF_FUSED_KIND", "1")
        want = L.IMPL_FUSED
    elif family == "second one-pass kernel":
        monkeypatch.setenv("RESNMTF_SMALL", "0")
        want = L.IMPL_FUSED
    prob, z, V = load(name)
    fit = prob.device_fit(ctx)
    try:
        done = fit.run(None, 1.0e-6)
        assert fit.counters()["impl"] == want
        assert done == len(z["all_error"])
        assert rel_err(fit.errors(), z["all_error"]) <= RTOL
        fit.normalise()
        outs = [fit.get_factors(v) for v in range(V)]
        for v in range(V):
            assert rel_err(outs[v][0], z[f"of_{v}"]) <= 1e-8  # hundreds of sweeps: allow 10x the per-sweep bar
            assert rel_err(outs[v][2], z[f"og_{v}"]) <= 1e-8
            assert rel_err(outs[v][1], z[f"os_{v}"]) <= 1e-8
        rows, cols, _ = O.binarise([o[0] for o in outs], [o[2] for o in outs], [o[1] for o in outs])
        for v in range(V):
            assert np.array_equal(rows[v].astype(np.uint8), z[f"rows_{v}"])
            assert np.array_equal(cols[v].astype(np.uint8), z[f"cols_{v}"])
    finally:
        fit.close()
