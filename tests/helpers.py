"""Shared helpers of the parity tests: run the same seeded problem through the oracle and through the
C-ABI (DeviceFit) and compare sweep by sweep."""
from __future__ import annotations

import numpy as np

from oracle import resnmtf_oracle as O

RTOL = 1.0e-9  # BASELINE.json north_star: factors and objective within 1e-9 relative in FP64


def rel_err(a, b):
    """max elementwise relative error (entries that are exactly equal -- incl. both zero -- count 0)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if np.isnan(b).any() or np.isnan(a).any():
        assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
        m = ~np.isnan(b)
        a, b = a[m], b[m]
    d = np.abs(a - b)
    den = np.maximum(np.abs(b), 1e-290)
    r = np.where(d == 0, 0.0, d / den)
    return float(r.max()) if r.size else 0.0


def names_to_maps(names_v, names_w):
    """index pairs of the names two views share (what the host side hands to set_shared_map)."""
    pos_w = {s: i for i, s in enumerate(names_w)}
    iv = [i for i, s in enumerate(names_v) if s in pos_w]
    iw = [pos_w[names_v[i]] for i in iv]
    return np.asarray(iv, dtype=np.int32), np.asarray(iw, dtype=np.int32)


class Problem:
    """A fully specified loop input: views (already prepped), names, symmetrised restrictions, inits."""

    def __init__(self, data, k_vec, init_f, init_s, init_g, phi=None, xi=None, psi=None,
                 row_names=None, col_names=None):
        self.data = [np.asfortranarray(x, dtype=np.float64) for x in data]
        V = len(self.data)
        self.k_vec = list(k_vec)
        self.init_f, self.init_s, self.init_g = init_f, init_s, init_g
        z = np.zeros((V, V))
        self.phi = z.copy() if phi is None else np.asarray(phi, dtype=np.float64)
        self.xi = z.copy() if xi is None else np.asarray(xi, dtype=np.float64)
        self.psi = z.copy() if psi is None else np.asarray(psi, dtype=np.float64)
        dn_r, dn_c = O.default_names(self.data)
        self.row_names = dn_r if row_names is None else row_names
        self.col_names = dn_c if col_names is None else col_names
        self.row_indices = O.shared_names(self.row_names)
        self.col_indices = O.shared_names(self.col_names)

    def oracle(self, n_iters=None, trace=None, max_iters=None):
        return O.res_nmtf_loop(self.data, self.row_indices, self.col_indices, self.row_names, self.col_names,
                               self.init_f, self.init_s, self.init_g, self.k_vec, self.phi, self.xi, self.psi,
                               n_iters=n_iters, trace=trace, max_iters=max_iters)

    def device_fit(self, ctx, err_mode=0, impl=0):
        from resnmtf_b200 import _lib as L
        from resnmtf_b200.device import DeviceFit

        V = len(self.data)
        fit = DeviceFit(ctx, [x.shape[0] for x in self.data], [x.shape[1] for x in self.data], self.k_vec)
        fit.set_options(err_mode=err_mode, impl=impl)
        for v in range(V):
            fit.set_data(v, self.data[v])
            fit.set_factors(v, self.init_f[v], self.init_s[v], self.init_g[v])
        fit.set_restrictions(self.phi, self.xi, self.psi)
        for v in range(V):
            for w in range(V):
                if w == v:
                    continue
                fit.set_shared_map(L.MAP_ROW, v, w, *names_to_maps(self.row_names[v], self.row_names[w]))
                fit.set_shared_map(L.MAP_COL, v, w, *names_to_maps(self.col_names[v], self.col_names[w]))
        return fit


def compare_trace(problem, ctx, n_iters, err_mode=0, impl=0, rtol=RTOL):
    """Steps the device one sweep at a time and compares every sweep with the oracle's trace."""
    states = []
    problem.oracle(n_iters=n_iters,
                   trace=lambda t, cf, cs, cg, cl, cm, err: states.append(
                       ([a.copy() for a in cf], [a.copy() for a in cs], [a.copy() for a in cg],
                        [a.copy() for a in cl], [a.copy() for a in cm], err.copy())))
    fit = problem.device_fit(ctx, err_mode=err_mode, impl=impl)
    worst = 0.0
    try:
        for t in range(n_iters):
            fit.step()
            errs, _ = fit.view_errors()
            cf, cs, cg, cl, cm, oerr = states[t]
            for v in range(len(problem.data)):
                f, s, g, lam, mu = fit.get_factors(v)
                for name, a, b in (("F", f, cf[v]), ("S", s, cs[v]), ("G", g, cg[v]),
                                   ("lambda", lam, cl[v]), ("mu", mu, cm[v]),
                                   ("err", errs[v:v + 1], oerr[v:v + 1])):
                    r = rel_err(a, b)
                    worst = max(worst, r)
                    assert r <= rtol, f"sweep {t} view {v} {name}: rel err {r:.3e} > {rtol:g}"
        mean_hist = fit.errors()
        assert len(mean_hist) == n_iters
    finally:
        fit.close()
    return worst
