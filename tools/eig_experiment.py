"""Experiment: top-16 eigenpairs of the 4000 x 4000 Gram matrix of a fit's initialisation through cuSOLVER's
selected-range solver (cusolverDnDsyevdx, range I) against the full torch.linalg.eigh (syevd) the units use now:
time and agreement (eigenvalues, |eigenvectors|) on a planted and on a shuffled (nearly degenerate) view.
Usage: python tools/eig_experiment.py [n p]"""
import ctypes as C
import glob
import os
import site
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from resnmtf_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
KC = 16
dev = torch.device("cuda", 0)
path = [f for sp in site.getsitepackages() for f in glob.glob(sp + "/nvidia/cusolver/lib/libcusolver.so.*")][0]
lib = C.CDLL(path)
handle = C.c_void_p()
assert lib.cusolverDnCreate(C.byref(handle)) == 0
stream = torch.cuda.current_stream(dev).cuda_stream
assert lib.cusolverDnSetStream(handle, C.c_void_p(stream)) == 0
VEC, RANGE_I, LOWER = 1, 1002, 0


def syevdx_top(gram, kc):
    m = gram.shape[0]
    a = gram.clone()  # symmetric: row-major == column-major; overwritten with the eigenvectors (first meig COLUMNS)
    w = torch.empty(m, dtype=torch.float64, device=dev)
    meig, lwork = C.c_int(0), C.c_int(0)
    args = (handle, VEC, RANGE_I, LOWER, m, C.c_void_p(a.data_ptr()), m, C.c_double(0.0), C.c_double(0.0),
            m - kc + 1, m, C.byref(meig), C.c_void_p(w.data_ptr()))
    st = lib.cusolverDnDsyevdx_bufferSize(*args, C.byref(lwork))
    assert st == 0, st
    work = torch.empty(lwork.value, dtype=torch.float64, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    st = lib.cusolverDnDsyevdx(*args, C.c_void_p(work.data_ptr()), lwork, C.c_void_p(info.data_ptr()))
    assert st == 0, st
    assert int(info.item()) == 0 and meig.value == kc, (int(info.item()), meig.value)
    vals = w[:kc].flip(0)
    vecs = a[:kc, :].T.flip(1)  # column j of the column-major result is row j of the tensor
    return vals, vecs


def timed(fn, reps=4):
    out, ts = None, []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return out, ts


rng = np.random.default_rng(1)
x = synth.prep(synth.planted_view(n, p, 5, rng, row_prob=0.2, col_prob=0.2, height=5.0, sigma=1.0)[0])
xt = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
perm = torch.randperm(xt.numel(), device=dev)
xs = xt.reshape(-1)[perm].reshape(xt.shape)
xs = xs / xs.sum(dim=1, keepdim=True)
for name, m in (("planted", xt), ("shuffled", xs)):
    gram = m @ m.T
    (w_full, v_full), t_full = timed(lambda: torch.linalg.eigh(gram))
    (w_top, v_top), t_top = timed(lambda: syevdx_top(gram, KC))
    wf, vf = w_full[-KC:].flip(0), v_full[:, -KC:].flip(1)
    rel_w = float(((w_top - wf).abs() / wf.abs()).max())
    dv = float((v_top.abs() - vf.abs()).abs().max())
    # what the initialisation uses: U = X V / d, |U|, |V|
    d = torch.sqrt(wf)
    u_full, u_top = (m.T @ vf) / d, (m.T @ v_top) / torch.sqrt(w_top)
    du = float((u_top.abs() - u_full.abs()).abs().max())
    resid = float((gram @ v_top - v_top * w_top).abs().max())
    orth = float((v_top.T @ v_top - torch.eye(KC, dtype=torch.float64, device=dev)).abs().max())
    print(f"{name:9s} eigh {['%.1f' % (1e3 * t) for t in t_full]} ms | syevdx top-{KC} {['%.1f' % (1e3 * t) for t in t_top]} ms | "
          f"rel dW {rel_w:.2e}  max d|V| {dv:.2e}  max d|U| {du:.2e}  residual {resid:.2e}  orth {orth:.2e}")
    print("   top eigenvalues:", [f"{float(v):.6e}" for v in wf[:8]])
    out, t_f = timed(lambda: api._topk_eig_filtered(torch, gram, KC))
    if out is None:
        print(f"   filtered subspace iteration: NOT CONVERGED ({['%.1f' % (1e3 * t) for t in t_f]} ms)")
        continue
    w_i, v_i = out
    u_i = (m.T @ v_i) / torch.sqrt(w_i)
    print(f"   filtered subspace iteration {['%.1f' % (1e3 * t) for t in t_f]} ms | rel dW "
          f"{float(((w_i - wf).abs() / wf.abs()).max()):.2e}  max d|V| {float((v_i.abs() - vf.abs()).abs().max()):.2e}  "
          f"max d|U| {float((u_i.abs() - u_full.abs()).abs().max()):.2e}  residual/lambda_max "
          f"{float((gram @ v_i - v_i * w_i).norm(dim=0).max() / wf[0]):.2e} (eigh: "
          f"{float((gram @ vf - vf * wf).norm(dim=0).max() / wf[0]):.2e})  orth "
          f"{float((v_i.T @ v_i - torch.eye(KC, dtype=torch.float64, device=dev)).abs().max()):.2e}")
