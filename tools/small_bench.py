"""Iteration rate on small multi-view problems (C1 toy, the reference's 180x180 test data) and on the C3 / C4
structures, through the C ABI.  Usage: python tools/small_bench.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import _lib as L  # noqa: E402
from resnmtf_b200 import synth  # noqa: E402
from resnmtf_b200.device import Context, DeviceFit  # noqa: E402


def run(ctx, name, shapes, k, phi=None, psi=None, xi=None, iters=300, impl=L.IMPL_AUTO):
    rng = np.random.default_rng(1)
    V = len(shapes)
    fit = DeviceFit(ctx, [s[0] for s in shapes], [s[1] for s in shapes], [k] * V)
    fit.set_options(err_mode=L.ERR_ALGEBRAIC, impl=impl)
    for v, (n, p) in enumerate(shapes):
        x = rng.random((n, p))
        x /= x.sum(0)[None, :]
        fit.set_data(v, x)
        fit.set_factors(v, *synth.random_factors(n, p, k, rng))
    fit.set_restrictions(phi, xi, psi)
    for v in range(V):
        for w in range(V):
            if v != w and shapes[v][0] == shapes[w][0]:
                idx = np.arange(shapes[v][0], dtype=np.int32)
                fit.set_shared_map(L.MAP_ROW, v, w, idx, idx)
            if v != w and shapes[v][1] == shapes[w][1]:
                idx = np.arange(shapes[v][1], dtype=np.int32)
                fit.set_shared_map(L.MAP_COL, v, w, idx, idx)
    fit.run(20)
    t0 = time.perf_counter()
    fit.run(iters)
    dt = time.perf_counter() - t0
    c = fit.counters()
    gb = c["alg_bytes_per_iter"] * iters / (c["device_ms"] * 1e-3) * 1e-9
    print(f"{name:34s} {c['device_ms'] / iters * 1e3:9.1f} us/iter device, {dt / iters * 1e6:9.1f} us/iter host wall, "
          f"{gb:7.0f} GB/s alg, {c['kernel_launches'] // iters} launches/iter", flush=True)
    fit.close()


def sym(V, val, pairs=None):
    m = np.zeros((V, V))
    if pairs is None:
        m[np.triu_indices(V, 1)] = val
    else:
        for a, b in pairs:
            m[a, b] = val
    return m + m.T


def main():
    ctx = Context()
    run(ctx, "C1 toy 100x50 + 100x30, phi", [(100, 50), (100, 30)], 3, phi=sym(2, 200.0), iters=2000)
    run(ctx, "test data 2 x 180x180", [(180, 180)] * 2, 3, iters=2000)
    run(ctx, "1 view 2000x1000 k=5", [(2000, 1000)], 5, iters=1000)
    run(ctx, "C3 4 x 50000x5000 k=5 phi,psi", [(50000, 5000)] * 4, 5, phi=sym(4, 200.0, [(0, 1), (2, 3)]),
        psi=sym(4, 200.0, [(0, 2), (1, 3)]), iters=20)
    run(ctx, "C4 8 x 100000x2000 k=8 phi,psi,xi", [(100000, 2000)] * 8, 8, phi=sym(8, 200.0), psi=sym(8, 200.0),
        xi=sym(8, 50.0), iters=10)
    ctx.close()


if __name__ == "__main__":
    main()
