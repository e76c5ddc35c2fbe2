"""Iteration rate on small multi-view problems (C1 toy, the reference's 180x180 test data) and on the C3 / C4
structures, through the C ABI.  Usage: python tools/small_bench.py [--gpus N] [--only c3,c4]
--gpus N > 1 also runs C3 / C4 as PLACED fits (resnmtf_fit_create_placed): the views of one fit spread over N GPUs, the
coupled factor rows read over NVLink, the Gauss-Seidel order kept by events between the GPUs' streams."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import _lib as L  # noqa: E402
from resnmtf_b200 import synth  # noqa: E402
from resnmtf_b200.device import Context, DeviceFit  # noqa: E402


_CACHE = {}


def run(ctx, name, shapes, k, phi=None, psi=None, xi=None, iters=300, impl=L.IMPL_AUTO, placement=None):
    """placement: list of contexts, one per view (placed fit); None: every view on ctx."""
    rng = np.random.default_rng(1)
    V = len(shapes)
    fit = DeviceFit(placement if placement is not None else ctx, [s[0] for s in shapes], [s[1] for s in shapes], [k] * V)
    fit.set_options(err_mode=L.ERR_ALGEBRAIC, impl=impl)
    for v, (n, p) in enumerate(shapes):
        if (n, p) not in _CACHE:  # one synthetic matrix per shape (timing only: the views of a fit may be equal)
            _CACHE.clear()
            x = rng.random((n, p))
            x /= x.sum(0)[None, :]
            _CACHE[(n, p)] = np.asfortranarray(x)
        fit.set_data(v, _CACHE[(n, p)])
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
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--only", default="c1,test,small,c3,c4")
    a = ap.parse_args()
    only = set(a.only.split(","))
    ctx = Context(0)
    others = [Context(g) for g in range(1, a.gpus)]
    pool = [ctx] + others
    if "c1" in only:
        run(ctx, "C1 toy 100x50 + 100x30, phi", [(100, 50), (100, 30)], 3, phi=sym(2, 200.0), iters=2000)
    if "test" in only:
        run(ctx, "test data 2 x 180x180", [(180, 180)] * 2, 3, iters=2000)
    if "small" in only:
        run(ctx, "1 view 2000x1000 k=5", [(2000, 1000)], 5, iters=1000)
    c3 = dict(shapes=[(50000, 5000)] * 4, k=5, phi=sym(4, 200.0, [(0, 1), (2, 3)]), psi=sym(4, 200.0, [(0, 2), (1, 3)]),
              iters=20)
    c4 = dict(shapes=[(100000, 2000)] * 8, k=8, phi=sym(8, 200.0), psi=sym(8, 200.0), xi=sym(8, 50.0), iters=10)
    if "c3" in only:
        run(ctx, "C3 4 x 50000x5000 k=5 phi,psi", **c3)
        run(ctx, "C3 shapes, uncoupled", **dict(c3, phi=None, psi=None))
        if a.gpus > 1:
            # dependency graph of the sweep: view 1 -> {2, 3} -> 4 (phi 12, 34; psi 13, 24): views 2 and 3 on different GPUs
            place = [pool[0], pool[0], pool[1 % a.gpus], pool[1 % a.gpus]] if a.gpus < 4 else pool[:4]
            run(ctx, f"C3 placed on {min(a.gpus, 4)} GPUs", placement=place, **c3)
    if "c4" in only:
        run(ctx, "C4 8 x 100000x2000 k=8 phi,psi,xi", **c4)
        if a.gpus > 1:
            run(ctx, f"C4 placed on {min(a.gpus, 8)} GPUs (all pairs coupled: a chain)",
                placement=[pool[v % a.gpus] for v in range(8)], **c4)
            c4u = dict(c4, phi=None, psi=None, xi=None)
            run(ctx, "C4 shapes, uncoupled, 1 GPU", **c4u)
            run(ctx, f"C4 shapes, uncoupled, placed on {min(a.gpus, 8)} GPUs",
                placement=[pool[v % a.gpus] for v in range(8)], **c4u)
    for c in others:
        c.close()
    ctx.close()


if __name__ == "__main__":
    main()
