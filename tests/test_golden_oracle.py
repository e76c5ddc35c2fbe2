"""The oracle still reproduces the committed golden fixtures (a change in the oracle is noticed)."""
import numpy as np
import pytest

from golden_util import CASES, N_SWEEPS, load
from helpers import rel_err
from oracle import resnmtf_oracle as O


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    prob, z, V = load(name)
    errs = []
    res = prob.oracle(n_iters=N_SWEEPS, trace=lambda t, cf, cs, cg, cl, cm, e: errs.append(e.copy()))
    assert rel_err(np.array(errs), z["sweep_errors"]) <= 1e-12
    for v in range(V):
        assert rel_err(res["raw_f"][v], z[f"f{N_SWEEPS}_{v}"]) <= 1e-11
        assert rel_err(res["raw_g"][v], z[f"g{N_SWEEPS}_{v}"]) <= 1e-11
        assert rel_err(res["raw_s"][v], z[f"s{N_SWEEPS}_{v}"]) <= 1e-11


def test_golden_convergence_margin_is_comfortable():
    """The stop decision |d err| <= 1e-6 of every fixture is not within rounding of the threshold, so a
    1e-9-accurate implementation must stop on the same sweep."""
    for name in CASES:
        _, z, _ = load(name)
        e = np.concatenate([[0.0], z["all_error"]])
        d = np.abs(np.diff(e))
        assert d[-1] <= 1e-6 and (d[:-1] > 1e-6).all()
        margin = min(abs(d[-1] - 1e-6), abs(d[-2] - 1e-6)) / 1e-6
        assert margin > 1e-6, (name, margin)
