"""Where apply_resnmtf's wall time goes on a mid-size view (host SVD initialisation, device loop, bisilhouette,
shuffle refits, stability).  Usage: python tools/apply_profile.py [n p]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import synth  # noqa: E402
from resnmtf_b200.api import apply_resnmtf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
rng = np.random.default_rng(3)
x, _, _ = synth.planted_view(n, p, 4, rng, row_prob=0.2, col_prob=0.2, height=5.0, sigma=1.0)
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
res = apply_resnmtf([x], k_min=3, k_max=6, rng=np.random.default_rng(5), max_iters=2000)
pr.disable()
dt = time.perf_counter() - t0
print(f"apply_resnmtf {n}x{p}: {dt:.2f} s, selected k = {res['output_f'][0].shape[1]}, bisil = {res['bisil']:.4f}")
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(22)
