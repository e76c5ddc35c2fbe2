"""Deterministic synthetic inputs for tests and bench.py (planted-bicluster model of the reference's own
test and README data: tests/testthat/test-resnmtf.R:38-52, README.md:37-58), followed by the reference's
prep (column shift to non-negative, L1 column normalisation -- R/utils.r:20-27,86-88).

Everything is seeded NumPy (PCG64), so host and test fixtures agree bit for bit across machines."""
from __future__ import annotations

import numpy as np


def prep(x):
    """make_non_neg_inner + matrix_normalisation (R/utils.r:20-27,86-88) on one view."""
    x = np.asarray(x, dtype=np.float64)
    x = x + np.abs(np.minimum(0.0, x.min(axis=0)))[None, :]
    return np.asfortranarray(x / x.sum(axis=0)[None, :])


def planted_view(n, p, n_planted, rng, row_prob=0.2, col_prob=0.2, height=5.0, sigma=1.0, rows=None, cols=None,
                 chunk=4096):
    """X = R diag(height) C' + sigma |N(0,1)|, column-major float64.  ``rows``/``cols`` reuse a membership
    matrix (shared rows / columns across views).  Built in row chunks so large views stay in budget."""
    if rows is None:
        rows = (rng.random((n, n_planted)) < row_prob).astype(np.float64)
    if cols is None:
        cols = (rng.random((p, n_planted)) < col_prob).astype(np.float64)
    x = np.empty((n, p), dtype=np.float64, order="F")
    ct = np.ascontiguousarray(cols.T) * height
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        x[r0:r1, :] = rows[r0:r1] @ ct + sigma * np.abs(rng.standard_normal((r1 - r0, p)))
    return x, rows, cols


def block_views(n_views=2, block=60, n_blocks=3, height=10.0, sigma=0.1, seed=0):
    """The reference's test data (test-resnmtf.R:38-52): disjoint diagonal blocks + 0.1 |N(0,1)|."""
    rng = np.random.default_rng(seed)
    n = block * n_blocks
    rc = np.zeros((n, n_blocks))
    for i in range(n_blocks):
        rc[i * block:(i + 1) * block, i] = 1.0
    base = rc @ (height * np.eye(n_blocks)) @ rc.T
    return [np.asfortranarray(base + sigma * np.abs(rng.standard_normal((n, n)))) for _ in range(n_views)], rc


def random_factors(n, p, k, rng):
    """Explicit positive inits with unit column sums (what init_mats hands to the loop when init_* are
    given; used wherever an SVD would only slow a test or a benchmark down)."""
    f = rng.random((n, k)) + 0.05
    g = rng.random((p, k)) + 0.05
    f /= f.sum(axis=0)[None, :]
    g /= g.sum(axis=0)[None, :]
    s = np.abs(np.diag(rng.random(k) + 0.5)) + np.abs(np.sqrt(0.05) * rng.standard_normal((k, k)))
    return np.asfortranarray(f), np.asfortranarray(s), np.asfortranarray(g)


def config_seed(config, view=0):
    """Seeds of SURVEY 8(d): 20260000 + 100*config + view."""
    return 20260000 + 100 * int(config) + int(view)
