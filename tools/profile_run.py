"""Short single-fit run for ncu: the C2 view (20000 x 4000), one k, a few sweeps.
Usage: python tools/profile_run.py [--k 8 --iters 5 --n 20000 --p 4000]"""
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
a = ap.parse_args()
rng = np.random.default_rng(1)
x = np.asfortranarray(rng.random((a.p, a.n)).T)
x /= x.sum(axis=0)[None, :]
f, s, g = synth.random_factors(a.n, a.p, a.k, rng)
ctx = Context()
fit = DeviceFit(ctx, [a.n], [a.p], [a.k])
fit.set_options(err_mode=a.err)
fit.set_data(0, x)
fit.set_factors(0, f, s, g)
fit.run(a.iters)
c = fit.counters()
print(f"k={a.k} iters={a.iters} device_ms={c['device_ms']:.3f} launches={c['kernel_launches']} err={fit.errors()[-1]:.6f}")
fit.close()
ctx.close()
