"""Do two independent fits interleave on one GPU?  The one-pass kernel ends every update-iteration with a tail in which
most SMs idle (T publication, cross-cluster reduction, G update, single-CTA view finish); a second fit on another
stream can start its clusters on the SMs the first one has left.  Measures update-iterations/s of the bench workload
(six k-fits on the 20000 x 4000 view) with the fits driven from 1, 2 and 3 host threads / contexts.
Usage: python tools/interleave_bench.py [--iters 300]"""
import argparse
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import _lib as L  # noqa: E402
from resnmtf_b200 import synth  # noqa: E402
from resnmtf_b200.device import Context, DeviceData, DeviceFit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=300)
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--p", type=int, default=4000)
a = ap.parse_args()
rng = np.random.default_rng(1)
x = np.asfortranarray(rng.random((a.p, a.n)).T)
x /= x.sum(axis=0)[None, :]
ks = [3, 4, 5, 6, 7, 8]
inits = {k: synth.random_factors(a.n, a.p, k, rng) for k in ks}
for n_ctx in (1, 2, 3):
    ctxs = [Context(0) for _ in range(n_ctx)]
    datas = [DeviceData(c, x) for c in ctxs]
    fits = {}
    for i, k in enumerate(ks):
        c = i % n_ctx
        f = DeviceFit(ctxs[c], [a.n], [a.p], [k])
        f.set_options(err_mode=L.ERR_ALGEBRAIC, impl=L.IMPL_AUTO)
        f.attach_data(0, datas[c])
        f.set_factors(0, *inits[k])
        f.run(5)
        fits[k] = (c, f)

    def work(c):
        for k in ks:
            if fits[k][0] == c:
                fits[k][1].run(a.iters)

    best = None
    for rep in range(3):
        ths = [threading.Thread(target=work, args=(c,)) for c in range(n_ctx)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(f"{n_ctx} context(s): {len(ks) * a.iters / best:9.1f} update-iterations/s  ({1e6 * best / (len(ks) * a.iters):.1f} us per "
          f"update-iteration, wall clock, best of 3)")
    for k in ks:
        fits[k][1].close()
    for d in datas:
        d.close()
    for c in ctxs:
        c.close()
