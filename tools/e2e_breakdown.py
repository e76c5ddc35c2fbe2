"""Where the end-to-end k-sweep time goes (host wall clock around each C-ABI call, C2 shape).
Usage: python tools/e2e_breakdown.py [iters]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import synth  # noqa: E402
from resnmtf_b200.device import Context, DeviceData, DeviceFit  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n, p = 20000, 4000
rng = np.random.default_rng(1)
xt = torch.from_numpy(np.ascontiguousarray(rng.random((p, n)))).pin_memory()
x = xt.numpy().T
ctx = Context()
inits = {k: synth.random_factors(n, p, k, rng) for k in (3, 4, 5, 6, 7, 8)}


def sweep(report):
    acc = {}

    def tick(name, t0):
        torch.cuda.synchronize()
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0

    t_all = time.perf_counter()
    t0 = time.perf_counter(); data = DeviceData(ctx, x); tick("data_create (H2D + re-tile + norm)", t0)
    for k in (3, 4, 5, 6, 7, 8):
        f0, s0, g0 = inits[k]
        t0 = time.perf_counter(); fit = DeviceFit(ctx, [n], [p], [k]); tick("fit_create", t0)
        t0 = time.perf_counter(); fit.attach_data(0, data); tick("attach_data", t0)
        t0 = time.perf_counter(); fit.set_factors(0, f0, s0, g0); tick("set_factors", t0)
        t0 = time.perf_counter(); fit.run(1); tick("first run(1): plan + X8 + graph", t0)
        t0 = time.perf_counter(); fit.run(iters - 1); tick(f"run({iters - 1})", t0)
        t0 = time.perf_counter(); fit.errors(); fit.normalise(); tick("errors + normalise", t0)
        t0 = time.perf_counter(); fit.get_factors(0); tick("get_factors", t0)
        t0 = time.perf_counter(); fit.close(); tick("fit_destroy", t0)
    t0 = time.perf_counter(); data.close(); tick("data_destroy", t0)
    total = time.perf_counter() - t_all
    if report:
        for name, v in acc.items():
            print(f"  {name:38s} {v * 1e3:8.2f} ms")
        print(f"  total {total * 1e3:.2f} ms for {6 * iters} update-iterations -> {6 * iters / total:.0f} it/s")


sweep(False)
sweep(True)
ctx.close()
