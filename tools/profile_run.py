"""Short single-fit run for ncu: the C2 view (20000 x 4000), one k, a few sweeps; optionally several identical-shape
views coupled by phi / psi / xi on every pair (shared rows and columns = all, identity maps).
Usage: python tools/profile_run.py [--k 8 --iters 5 --n 20000 --p 4000 --views 1 --phi 0 --psi 0 --xi 0 --impl 0]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import _lib as L  # noqa: E402
from resnmtf_b200 import synth  # noqa: E402
from resnmtf_b200.device import Context, DeviceFit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=8)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--p", type=int, default=4000)
ap.add_argument("--err", type=int, default=L.ERR_AUTO)
ap.add_argument("--impl", type=int, default=L.IMPL_AUTO)
ap.add_argument("--views", type=int, default=1)
ap.add_argument("--phi", type=float, default=0.0)
ap.add_argument("--psi", type=float, default=0.0)
ap.add_argument("--xi", type=float, default=0.0)
a = ap.parse_args()
rng = np.random.default_rng(1)
V = a.views
ctx = Context()
fit = DeviceFit(ctx, [a.n] * V, [a.p] * V, [a.k] * V)
fit.set_options(err_mode=a.err, impl=a.impl)
for v in range(V):
    x = np.asfortranarray(rng.random((a.p, a.n)).T)
    x /= x.sum(axis=0)[None, :]
    fit.set_data(v, x)
    fit.set_factors(v, *synth.random_factors(a.n, a.p, a.k, rng))


def full(val):
    m = np.full((V, V), float(val))
    np.fill_diagonal(m, 0.0)
    return m


if V > 1:
    fit.set_restrictions(full(a.phi), full(a.xi), full(a.psi))
    ir = np.arange(a.n, dtype=np.int32)
    ic = np.arange(a.p, dtype=np.int32)
    for v in range(V):
        for w in range(V):
            if v != w:
                fit.set_shared_map(L.MAP_ROW, v, w, ir, ir)
                fit.set_shared_map(L.MAP_COL, v, w, ic, ic)
fit.run(a.iters)
c = fit.counters()
print(f"k={a.k} views={V} iters={a.iters} impl={c['impl']} device_ms={c['device_ms']:.3f} "
      f"launches={c['kernel_launches']} err={fit.errors()[-1]:.6f}")
fit.close()
ctx.close()
