"""Wall time of the native building blocks of an apply_resnmtf call on the bench view (20000 x 4000): upload + prep,
shuffle, Gram + top-k triplets, a shuffled-refit unit, and a 36-unit k-sweep batch on the visible GPUs.
Usage: python tools/native_profile.py [--gpus N]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import synth  # noqa: E402
from resnmtf_b200.native import NativePool  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=0)
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--p", type=int, default=4000)
a = ap.parse_args()
rng = np.random.default_rng(synth.config_seed(2, 0))
x, _, _ = synth.planted_view(a.n, a.p, 5, rng, row_prob=0.2, col_prob=0.2, height=5.0, sigma=1.0)


def timed(label, fn, reps=1):
    ts = []
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    print(f"{label:52s} " + " ".join(f"{1e3 * t:8.1f}" for t in ts) + " ms", flush=True)
    return out


with NativePool(a.gpus) as pool:
    print(f"pool of {len(pool)} GPU(s), view {a.n} x {a.p}")
    timed("upload (pageable) + prep", lambda: pool.put_host(0, [x], prep=True), reps=2)
    base = pool.get(0, 0, 0)
    sh = timed("shuffle + re-normalise (new handle)", lambda: base.shuffle(5), reps=3)
    timed("top-16 triplets of the shuffled view (Gram + solve)", lambda: base.shuffle(6).svd_topk(8), reps=2)
    timed("top-16 triplets of the planted view", lambda: base.svd_topk(8), reps=2)
    sub = timed("sub-sample 0.9 x 0.9 (new handle)",
                lambda: base.subsample(np.sort(rng.permutation(a.n)[: int(0.9 * a.n)]),
                                       np.sort(rng.permutation(a.p)[: int(0.9 * a.p)])), reps=2)
    timed("sums of the sub-sample", lambda: sub.sums(), reps=2)
    k = 5
    noise = [np.abs(np.sqrt(0.05) * rng.standard_normal((k, k)))]
    r = timed("one shuffled-refit unit (k = 5, to convergence)",
              lambda: pool.run([dict(key=0, k=[k], noise=noise, shuffle_seed=77, n_iters=None, max_iters=5000)]), reps=3)
    print(f"   sweeps {r[0]['iters']}, unit seconds {r[0]['seconds']:.3f}")
    r = timed("one fit unit of the resident view (k = 5)",
              lambda: pool.run([dict(key=0, k=[k], noise=noise, n_iters=None, max_iters=5000)]), reps=2)
    print(f"   sweeps {r[0]['iters']}, unit seconds {r[0]['seconds']:.3f}")
    units = []
    for kk in range(3, 9):
        nz = [np.abs(np.sqrt(0.05) * rng.standard_normal((kk, kk)))]
        units.append(dict(key=0, k=[kk], noise=nz, n_iters=None, max_iters=5000))
        for rep in range(5):
            units.append(dict(key=0, k=[kk], noise=nz, shuffle_seed=100 * kk + rep, n_iters=None, max_iters=5000))
    r = timed(f"k-sweep batch: {len(units)} units on {len(pool)} GPU(s)", lambda: pool.run(units), reps=2)
    per = {}
    for u in r:
        per.setdefault(u["gpu"], []).append(u["seconds"])
    print("   busy seconds per GPU: " + " ".join(f"gpu{g}={sum(v):.2f}({len(v)})" for g, v in sorted(per.items())))
    print("   unit seconds: min %.3f max %.3f mean %.3f" % (min(u["seconds"] for u in r), max(u["seconds"] for u in r),
                                                           np.mean([u["seconds"] for u in r])))
