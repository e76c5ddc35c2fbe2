"""k-sweep wall time through apply_resnmtf on the BASELINE configs[1] view (20000 x 4000, k = 3..8, bisilhouette
selection; spurious-bicluster removal and stability switched on by flags).  Usage: python tools/ksweep_wall.py
[--spurious] [--stability] [--n 20000 --p 4000]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import synth  # noqa: E402
from resnmtf_b200.api import apply_resnmtf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--p", type=int, default=4000)
ap.add_argument("--spurious", action="store_true")
ap.add_argument("--stability", action="store_true")
a = ap.parse_args()
rng = np.random.default_rng(synth.config_seed(2, 0))
x, _, _ = synth.planted_view(a.n, a.p, 5, rng, row_prob=0.2, col_prob=0.2, height=5.0, sigma=1.0)
import torch  # noqa: E402,F401  (import cost kept out of the timed region)
from resnmtf_b200.device import default_context  # noqa: E402

default_context()
t0 = time.perf_counter()
res = apply_resnmtf([x], k_min=3, k_max=8, spurious=a.spurious, stability=a.stability,
                    rng=np.random.default_rng(5), max_iters=5000)
dt = time.perf_counter() - t0
print(f"apply_resnmtf {a.n}x{a.p} k sweep 3..8 spurious={a.spurious} stability={a.stability}: {dt:.2f} s wall, "
      f"selected k = {res['output_f'][0].shape[1]}, bisil = {res['bisil']:.4f}, "
      f"sweeps of the selected fit = {len(res['All_Error'])}")
