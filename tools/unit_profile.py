"""Where the wall time of ONE shuffled-refit unit of apply_resnmtf goes on the GPU (R/obtain_bicl.r:33-40 on the device:
permutation, re-normalisation, SVD initialisation, layout conversion, fit), step by step with synchronisation between
the steps.  Usage: python tools/unit_profile.py [n p k]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from resnmtf_b200 import api, synth  # noqa: E402
from resnmtf_b200.device import DeviceFit, default_context  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 5
ctx = default_context()
dev = torch.device("cuda", ctx.device)
x = synth.prep(synth.planted_view(n, p, 5, np.random.default_rng(1), row_prob=0.2, col_prob=0.2, height=5.0, sigma=1.0)[0])
xt = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
api._torch_cuda()
rng = np.random.default_rng(3)


def timed(label, fn, acc):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    ctx.synchronize()
    acc.setdefault(label, []).append(time.perf_counter() - t0)
    return out


acc = {}
for rep in range(4):
    gen = torch.Generator(device=dev)
    gen.manual_seed(rep + 1)
    perm = timed("randperm", lambda: torch.randperm(xt.numel(), generator=gen, device=dev), acc)
    m = timed("gather", lambda: xt.reshape(-1)[perm].reshape(xt.shape), acc)
    ok = timed("zero row/col test", lambda: bool((m.sum(dim=0) == 0).any() or (m.sum(dim=1) == 0).any()), acc)
    del perm
    m = timed("normalise", lambda: m / m.sum(dim=1, keepdim=True), acc)
    gram = timed("gram GEMM", lambda: m @ m.T, acc)
    w, v = timed("eigh", lambda: torch.linalg.eigh(gram), acc)
    u, d, g = timed("triplets (gram + eigh + U) + download",
                    lambda: tuple(t.cpu().numpy() for t in api._gram_topk_torch(torch, m, k)), acc)
    f0, s0, g0, lam, mu = api._init_from_svd(u, d, g, k, rng)
    fit = timed("fit_create", lambda: DeviceFit(ctx, [n], [p], [k]), acc)
    timed("set_data_device (re-tiling)", lambda: fit.set_data_device(0, m.data_ptr(), n), acc)
    timed("set_factors", lambda: fit.set_factors(0, f0, s0, g0, lam, mu), acc)
    its = timed("run (convergence)", lambda: fit.run(None, 1.0e-6, 5000), acc)
    timed("normalise + get_factors", lambda: (fit.normalise(), fit.get_factors(0)), acc)
    timed("fit_close", fit.close, acc)
    acc.setdefault("sweeps", []).append(its)
    del m, gram, w, v
print(f"shuffled-refit unit on {n} x {p}, k = {k} (4 repeats; first includes first-touch costs):")
for label, vals in acc.items():
    if label == "sweeps":
        print(f"  {label:34s} {vals}")
    else:
        print(f"  {label:34s} " + " ".join(f"{1e3 * v:8.1f}" for v in vals) + "  ms")
