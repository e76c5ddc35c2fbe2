"""Fused one-pass kernel against the two-pass TMA kernels over view widths (cluster sizes 1..8), k = 5.
Usage: python tools/fused_widths.py [n_rows]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import _lib as L  # noqa: E402
from resnmtf_b200 import synth  # noqa: E402
from resnmtf_b200.device import Context, DeviceFit  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
ctx = Context()
rng = np.random.default_rng(1)
for p in (1000, 2000, 3000, 4000, 5000, 6000, 7000, 8000):
    x = np.asfortranarray(rng.random((p, n)).T)
    x /= x.sum(axis=0)[None, :]
    f, s, g = synth.random_factors(n, p, 5, rng)
    line = f"n={n} p={p}:"
    for impl in (L.IMPL_TMA, L.IMPL_FUSED):
        fit = DeviceFit(ctx, [n], [p], [5])
        fit.set_options(err_mode=L.ERR_ALGEBRAIC, impl=impl)
        fit.set_data(0, x)
        fit.set_factors(0, f, s, g)
        fit.run(5)
        fit.run(30)
        c = fit.counters()
        line += f"  impl {c['impl']}: {c['device_ms'] / 30 * 1e3:8.1f} us/iter"
        fit.close()
    print(line, flush=True)
ctx.close()
