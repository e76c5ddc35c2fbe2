"""One launch of the JSD pair kernel (resnmtf_jsd_pairs) at the size the spurious-bicluster test of a k = 8 fit on the
C2 view needs it: 48 factor-like columns of 20000 values, 960 pairs.  For ncu and for timing.
Usage: python tools/jsd_run.py [n k repeats]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import bicluster as B  # noqa: E402
from resnmtf_b200.device import default_context  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rng = np.random.default_rng(0)


def factors():
    f = np.abs(rng.standard_normal((n, k))) * 0.05 / n
    for c in range(k):
        rows = rng.random(n) < 0.15
        f[rows, c] += (3.0 + rng.random(int(rows.sum()))) / n
    return np.asfortranarray(f / f.sum(axis=0)[None, :])


ctx = default_context()
f_main = factors()
f_mess = [[factors()] for _ in range(reps)]
B._jsd_scores_device(f_mess, f_main, 0, reps, k, ctx)  # warm-up
t0 = time.perf_counter()
thr, per = B._jsd_scores_device(f_mess, f_main, 0, reps, k, ctx)
dt = time.perf_counter() - t0
print(f"n={n} k={k}: {len(thr) + per.size * reps * k} pairs in {1e3 * dt:.2f} ms (incl. bw.nrd0 of {(reps + 1) * k} "
      f"columns on the host, upload, kernel, download); mean threshold score {np.mean(thr):.4f}")
t0 = time.perf_counter()
want = B.jsd_calc(f_mess[0][0][:, 0], f_mess[1][0][:, 0])
print(f"host jsd_calc: {1e3 * (time.perf_counter() - t0):.2f} ms per pair; first pair {thr[0]:.12f} vs {want:.12f}")
