"""Coupled views on different GPUs (resnmtf_fit_create_placed): the partner rows of F / G and the
partner S of the phi / psi / xi terms (star_prod_relevant / star_prod, R/utils.r:39-78) are read from the partner's
memory inside the update kernels and the Gauss-Seidel order of update_matrices() (R/update_steps.r:282-314) is an event
chain between the GPUs' streams.  The whole placed code path (per-context metadata copies, events, per-view streams)
also runs on ONE GPU with two contexts on it -- that is what a 1-GPU box tests; the 2-GPU tests need
gpurun --gpus 2 -- python -m pytest tests/test_gpu_placed.py -m gpu."""
import numpy as np
import pytest

from helpers import RTOL, Problem, compare_trace, rel_err
from oracle import resnmtf_oracle as O
from resnmtf_b200 import _lib as L
from resnmtf_b200 import synth
from resnmtf_b200.device import Context

pytestmark = pytest.mark.gpu


def coupled_problem(n_views=3, n=260, p=150, k=3, seed=41, wide=False):
    """phi, psi and xi all non-zero, partial and permuted overlaps of the names; `wide` views take the one-pass kernels."""
    rng = np.random.default_rng(seed)
    if wide:
        n, p = 200, 800
    data = [synth.prep(synth.planted_view(n, p, 3, rng, 0.35, 0.35)[0]) for _ in range(n_views)]
    rn = [[f"r{i}" for i in range(n)]]
    cn = [[f"c{i}" for i in range(p)]]
    for v in range(1, n_views):
        rn.append([f"r{i}" for i in rng.permutation(n)] if v % 2 else [f"r{i}" for i in range(30, n)] + [f"s{v}_{i}" for i in range(30)])
        cn.append([f"c{i}" for i in range(20, p)] + [f"d{v}_{i}" for i in range(20)] if v % 2 else [f"c{i}" for i in rng.permutation(p)])
    V = n_views
    phi, psi, xi = np.zeros((V, V)), np.zeros((V, V)), np.zeros((V, V))
    for v in range(V - 1):
        phi[v, v + 1] = 200.0 / (v + 1)
        psi[v, v + 1] = 80.0
        xi[v, v + 1] = 25.0
    phi[0, V - 1] = 40.0
    fs = [synth.random_factors(n, p, k, rng) for _ in range(V)]
    return Problem(data, [k] * V, [f[0] for f in fs], [f[1] for f in fs], [f[2] for f in fs],
                   phi=O.init_rest_mats(phi, V), xi=O.init_rest_mats(xi, V), psi=O.init_rest_mats(psi, V),
                   row_names=rn, col_names=cn)


def run_and_fetch(prob, ctxs, sweeps, err_mode=L.ERR_AUTO, impl=L.IMPL_AUTO):
    fit = prob.device_fit(ctxs, err_mode=err_mode, impl=impl)
    try:
        if sweeps is None:
            fit.run(None, 1.0e-6)
        else:
            fit.run(sweeps)
        errs = fit.errors().copy()
        outs = [fit.get_factors(v) for v in range(len(prob.data))]
        launches = fit.counters()["kernel_launches"]
    finally:
        fit.close()
    return errs, outs, launches


@pytest.mark.parametrize("impl", [L.IMPL_AUTO, L.IMPL_TMA, L.IMPL_DFMA])
@pytest.mark.parametrize("err_mode", [L.ERR_AUTO, L.ERR_DIRECT])
def test_placed_views_on_two_contexts_of_one_gpu_match_the_oracle(ctx, impl, err_mode, monkeypatch):
    monkeypatch.setenv("RESNMTF_FUSED_MAX_PAD", "100000000")
    other = Context(ctx.device)
    try:
        prob = coupled_problem()
        worst = compare_trace(prob, [ctx, other, ctx], n_iters=6, err_mode=err_mode, impl=impl)
        assert worst <= RTOL
    finally:
        other.close()


def test_placed_fit_is_bit_identical_to_the_single_context_fit(ctx):
    """Same kernels, same operands, same order: where the views live changes nothing, down to the last bit -- fixed
    sweeps and the converged run (same stop sweep)."""
    other = Context(ctx.device)
    try:
        for wide in (False, True):
            prob = coupled_problem(n_views=4, wide=wide, seed=43)
            for sweeps in (7, None):
                e1, o1, _ = run_and_fetch(prob, ctx, sweeps)
                e2, o2, _ = run_and_fetch(prob, [ctx, other, other, ctx], sweeps)
                assert len(e1) == len(e2) and np.array_equal(e1, e2)
                for a, b in zip(o1, o2):
                    for x, y in zip(a, b):
                        assert np.array_equal(x, y)
    finally:
        other.close()


def test_placed_uncoupled_views_and_error_codes(ctx):
    other = Context(ctx.device)
    try:
        rng = np.random.default_rng(5)
        data = [synth.prep(synth.planted_view(120, 70, 3, rng, 0.3, 0.3)[0]) for _ in range(2)]
        fs = [synth.random_factors(120, 70, 4, rng) for _ in range(2)]
        prob = Problem(data, [4, 4], [f[0] for f in fs], [f[1] for f in fs], [f[2] for f in fs])
        assert compare_trace(prob, [ctx, other], n_iters=5) <= RTOL
        fit = prob.device_fit([ctx, other])
        with pytest.raises(L.ResnmtfError):
            fit.profile(1)
        fit.close()
    finally:
        other.close()


needs_two = pytest.mark.skipif(L.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")


@needs_two
@pytest.mark.parametrize("impl", [L.IMPL_AUTO, L.IMPL_TMA])
def test_placed_views_on_two_gpus_match_the_oracle(impl, monkeypatch):
    monkeypatch.setenv("RESNMTF_FUSED_MAX_PAD", "100000000")
    c0, c1 = Context(0), Context(1)
    try:
        prob = coupled_problem()
        assert compare_trace(prob, [c0, c1, c0], n_iters=6, impl=impl) <= RTOL
        wide = coupled_problem(n_views=4, wide=True, seed=47)
        assert compare_trace(wide, [c0, c1, c1, c0], n_iters=4, impl=impl) <= RTOL
    finally:
        c0.close()
        c1.close()


@needs_two
def test_placed_fit_on_two_gpus_is_bit_identical_to_one_gpu():
    c0, c1 = Context(0), Context(1)
    try:
        prob = coupled_problem(n_views=4, wide=True, seed=49)
        for sweeps in (6, None):
            e1, o1, _ = run_and_fetch(prob, c0, sweeps)
            e2, o2, _ = run_and_fetch(prob, [c0, c1, c0, c1], sweeps)
            assert len(e1) == len(e2) and np.array_equal(e1, e2)
            for a, b in zip(o1, o2):
                for x, y in zip(a, b):
                    assert np.array_equal(x, y)
    finally:
        c0.close()
        c1.close()
