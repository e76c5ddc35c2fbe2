"""k-sweep wall time through apply_resnmtf on the BASELINE configs[1] view (20000 x 4000, k = 3..8, bisilhouette
selection; spurious-bicluster removal and stability switched on by flags), on 1 GPU and on every GPU count in
--gpus.  The result of the call must not depend on the GPU count; the tool checks that.  RESNMTF_TRACE=1 adds the
wall time and per-GPU busy time of every phase of the pool.
Usage: python tools/ksweep_wall.py [--spurious] [--stability] [--n 20000 --p 4000] [--gpus 1,2,4,8]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import _lib as L  # noqa: E402
from resnmtf_b200 import synth  # noqa: E402
from resnmtf_b200.api import apply_resnmtf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--p", type=int, default=4000)
ap.add_argument("--spurious", action="store_true")
ap.add_argument("--stability", action="store_true")
ap.add_argument("--gpus", default="1")
ap.add_argument("--repeat", type=int, default=1)
a = ap.parse_args()
rng = np.random.default_rng(synth.config_seed(2, 0))
x, _, _ = synth.planted_view(a.n, a.p, 5, rng, row_prob=0.2, col_prob=0.2, height=5.0, sigma=1.0)
import torch  # noqa: E402,F401  (import cost kept out of the timed region)
from resnmtf_b200.device import default_context, device_contexts  # noqa: E402

ctx = default_context()
counts = [g for g in (int(s) for s in a.gpus.split(",")) if g <= L.device_count()]
# warm-up: CUDA contexts, cuSOLVER / cuBLAS handles and the library's memory pools of every GPU (one small call)
os.environ["RESNMTF_MAX_GPUS"] = str(max(counts))
small, _, _ = synth.planted_view(1200, 600, 3, np.random.default_rng(1), row_prob=0.3, col_prob=0.3)
apply_resnmtf([small], k_min=3, k_max=3 + max(counts), spurious=False, stability=False, rng=np.random.default_rng(2),
              max_iters=50)
first = None
for g in counts:
    os.environ["RESNMTF_MAX_GPUS"] = str(g)
    for _ in range(a.repeat):
        t0 = time.perf_counter()
        res = apply_resnmtf([x], k_min=3, k_max=8, spurious=a.spurious, stability=a.stability,
                            rng=np.random.default_rng(5), max_iters=5000)
        dt = time.perf_counter() - t0
        same = ""
        if first is None:
            first = res
        else:
            ok = all(np.array_equal(u, v) for key in ("output_f", "row_clusters", "col_clusters")
                     for u, v in zip(first[key], res[key])) and first["bisil"] == res["bisil"]
            same = ", identical to the first run" if ok else ", DIFFERS from the first run"
        print(f"apply_resnmtf {a.n}x{a.p} k sweep 3..8 spurious={a.spurious} stability={a.stability} on "
              f"{len(device_contexts(ctx))} GPU(s): {dt:.2f} s wall, selected k = {res['output_f'][0].shape[1]}, "
              f"bisil = {res['bisil']:.4f}, biclusters kept = {int((res['row_clusters'][0].sum(0) > 0).sum())}, "
              f"sweeps of the selected fit = {len(res['All_Error'])}{same}", flush=True)
