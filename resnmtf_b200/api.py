"""``res_nmtf_inner`` and ``apply_resnmtf`` with the reference's exact signatures, argument meaning and
return structure (R/main.r:32-140, 214-335), with the loop of R/main.r:50-109 replaced by one call into
the device-resident C-ABI fit (include/resnmtf_b200.h).  This image has no R, so the host side above the C
ABI is Python; an R package would keep its own R/main.r and call the same C entry points through .Call
(INTEGRATION.md).

Python-only additions are keyword-only: ``rng`` (a ``numpy.random.Generator`` standing in for R's global
RNG stream), ``ctx`` (a ``resnmtf_b200.device.Context``; default: one shared context on the current GPU)
and ``max_iters`` (a safety cap the reference's ``while`` does not have; 0 = none)."""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import prep, sharding
from .bicluster import obtain_biclusters
from .device import DeviceFit, default_context, device_contexts
from .prep import as_named
from .stability import stability_check


# --------------------------------------------------------------------------------------------------
# initialisation (R/update_steps.r:36-125) -- host side; the truncated-SVD-on-device row is "next" (N3)
# --------------------------------------------------------------------------------------------------


_torch_lock = __import__("threading").Lock()
_torch_warm = False


def _torch_cuda():
    """torch with a CUDA device, or None (torch is plumbing here: device tensors + library GEMM / eigensolver for the
    boundary inputs of a fit, never the update path).  The first caller initialises torch's lazily loaded linear
    algebra backend under a lock: that initialisation is not thread-safe, and the fits of a k-sweep run on one host
    thread per GPU."""
    global _torch_warm
    try:
        import torch
    except Exception:  # pragma: no cover - torch is part of the image
        return None
    if not torch.cuda.is_available():
        return None
    if not _torch_warm:
        with _torch_lock:
            if not _torch_warm:
                for dev in range(torch.cuda.device_count()):
                    e = torch.eye(4, dtype=torch.float64, device=torch.device("cuda", dev))
                    torch.linalg.eigh(e @ e)
                    torch.randperm(4, device=e.device)
                    torch.cuda.synchronize(dev)
                _torch_warm = True
    return torch


_TOPK_COLS = 16  # RESNMTF_MAX_K: the triplets are always computed to this width and sliced, so that a cached set (the
                 # k-sweep fits the same data for every k) and a freshly computed one are the same numbers


def torch_device(index):
    """torch.device of CUDA device ``index`` (< 0 or None: the current one)."""
    import torch

    return torch.device("cuda", int(index) if index is not None and index >= 0 else torch.cuda.current_device())


def _gram_topk_torch(torch, xt, k):
    """Top-k singular triplets of the n x p matrix whose column-major storage is the row-major p x n tensor ``xt``,
    through the Gram matrix of the smaller side.  Returns device tensors |U| (n x k), d (k), |V| (p x k)."""
    u, d, v = _gram_triplets_torch(torch, xt)
    return u[:, :k], d[:k], v[:, :k]


def _topk_eig_filtered(torch, gram, kc, guard=48, tol=1.0e-15, max_outer=40):
    """The ``kc`` largest eigenpairs of the symmetric positive semi-definite ``gram`` by Chebyshev-filtered subspace
    iteration with locking (Zhou & Saad's scheme): a block of kc + guard vectors; per outer iteration a Rayleigh-Ritz
    step, locking of the leading Ritz pairs whose residual is at rounding level of the matrix norm
    (``tol * lambda_max``, what a backward-stable dense solver delivers), explicit deflation of the locked pairs,
    then a Chebyshev filter that damps [0, smallest Ritz value of the block] -- its degree bounded so that the most
    amplified direction of the block gains at most e^23 on the least.  Only the top k <= 16 pairs of the Gram matrix
    are ever used (R/update_steps.r:93-95 keeps the first k singular triplets), and the full eigendecomposition is
    what a fit's unit spends most of its time in.  The start block comes from a fixed seed: same matrix, same
    answer.  Returns (eigenvalues descending, eigenvectors) or None when the iteration did not converge -- the
    caller then takes the dense solver."""
    import math

    m = gram.shape[0]
    b = min(m, kc + guard)
    dev, dt = gram.device, gram.dtype
    gen = torch.Generator(device=dev)
    gen.manual_seed(20260000)
    q, _ = torch.linalg.qr(torch.randn((m, b), generator=gen, device=dev, dtype=dt))
    work = gram  # deflated copy once something is locked: locked pairs then sit at eigenvalue 0, inside the damped range
    vlock = torch.empty((m, 0), device=dev, dtype=dt)
    lam_lock = []

    def off_locked(y):
        return y - vlock @ (vlock.T @ y) if vlock.shape[1] else y

    for _ in range(max_outer):
        aq = off_locked(work @ q)
        h = (q.T @ aq).cpu().numpy()  # Rayleigh-Ritz on the host: the b x b problem is too small for the GPU
        th, y = np.linalg.eigh(0.5 * (h + h.T))
        th, y = th[::-1].copy(), np.ascontiguousarray(y[:, ::-1])
        y = torch.from_numpy(y).to(dev)
        theta = torch.from_numpy(th).to(dev)
        q, aq = q @ y, aq @ y
        rs = torch.linalg.vector_norm(aq - q * theta[None, :], dim=0).cpu().numpy()
        scale = max(abs(float(th[0])), lam_lock[0] if lam_lock else 0.0)
        n_new = 0  # leading pairs only, in order: nothing above a locked pair is still moving
        while n_new < len(th) and len(lam_lock) + n_new < kc and rs[n_new] <= tol * scale:
            n_new += 1
        if n_new:
            vn = q[:, :n_new]
            vlock = torch.cat([vlock, vn], dim=1)
            lam_lock += [float(v) for v in th[:n_new]]
            if work is gram:
                work = gram.clone()
            work.addmm_(vn * theta[:n_new][None, :], vn.T, alpha=-1.0)
            q, th = q[:, n_new:], th[n_new:]
        if len(lam_lock) >= kc:
            return torch.tensor(lam_lock, device=dev, dtype=dt), vlock
        pos = th[th > 0.0]
        if q.shape[1] < 2 or pos.size == 0:
            return None
        c = float(pos[-1])  # the unwanted part of the (deflated) spectrum lies in [0, c]
        t_max = max((2.0 * float(th[0]) - c) / c, 1.0 + 1.0e-12)
        deg = int(min(40, max(2, math.floor(23.0 / math.acosh(t_max)))))
        e = 0.5 * c
        # T_j((W - e I) / e) q by the three-term recurrence, one fused GEMM per degree on the shifted matrix
        shifted = work.clone()
        shifted.diagonal().sub_(e)
        y0, y1 = q, (shifted @ q) / e
        for _ in range(2, deg + 1):
            y0, y1 = y1, torch.addmm(y0, shifted, y1, beta=-1.0, alpha=2.0 / e)
        del shifted
        q, _ = torch.linalg.qr(off_locked(y1))
    return None


def _topk_eigh(torch, gram, kc):
    """(eigenvalues descending, eigenvectors) of the kc largest eigenpairs of ``gram``: the Chebyshev-filtered subspace
    iteration for matrices of order >= 1024 (23-28 ms against 84 ms for the dense solver at order 4000 on B200, same
    residual level, tools/eig_experiment.py), the dense solver for smaller ones, as the fallback when the iteration
    does not converge, and always with RESNMTF_TOPK=eigh."""
    import os

    if gram.shape[0] >= 1024 and os.environ.get("RESNMTF_TOPK", "filtered") == "filtered":
        out = _topk_eig_filtered(torch, gram, kc)
        if out is not None:
            return out
    w, v = torch.linalg.eigh(gram)
    return w[-kc:].flip(0), v[:, -kc:].flip(1)


def _gram_triplets_torch(torch, xt):
    p, n = xt.shape
    kc = min(_TOPK_COLS, p, n)
    if p <= n:
        w, v = _topk_eigh(torch, xt @ xt.T, kc)
        d = torch.sqrt(torch.clamp(w, min=0.0))
        u = (xt.T @ v) / d[None, :]
    else:
        w, u = _topk_eigh(torch, xt.T @ xt, kc)
        d = torch.sqrt(torch.clamp(w, min=0.0))
        v = (xt @ u) / d[None, :]
    return u.abs(), d, v.abs()


def _svd_topk_device(x, k, device):
    """Top-k singular triplets through the Gram matrix of the smaller side, ON THE GPU: X'X (or XX') is one FP64 GEMM,
    its eigendecomposition one cuSOLVER call, U = X V / d one more GEMM -- library code through torch, this is the
    boundary input of a fit, not the update path.  The reference's full svd(x) (LAPACK,
    R/update_steps.r:92) is what dominates the wall time of apply_resnmtf once the loop runs on the device (host
    profile on 4000 x 1500: 33 of 45 s); the Gram route agrees with it to ~1e-14 on planted and on shuffled (nearly
    degenerate) spectra.  Returns None when torch / CUDA is unavailable.
    Default (native route): the library's own resnmtf_data_svd_topk on a temporary data handle -- hand-written FP64
    tensor-core Gram / projection kernels and a filtered subspace iteration on the device, no torch."""
    if native_route():
        if L.device_count() <= 0:
            return None
        from .device import Context, DeviceData

        ctx = default_context() if device is None or device < 0 or device == default_context().device else Context(device)
        h = DeviceData(ctx, np.asfortranarray(x, dtype=np.float64))
        try:
            return h.svd_topk(k)
        finally:
            h.close()
            if ctx is not default_context():
                ctx.close()
    torch = _torch_cuda()
    if torch is None:
        return None
    dev = torch_device(device)
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)  # p x n, row-major (== column-major n x p)
    u, d, v = _gram_topk_torch(torch, xt, k)
    return u.cpu().numpy(), d.cpu().numpy(), v.cpu().numpy()


def _init_from_svd(f, d, g, k, rng, sigma=0.05):
    """R/update_steps.r:93-105 on the top-k triplets: S = |diag(d)| + |noise|, columns rescaled like the final
    normalisation (quirk Q7), lambda / mu = column sums of the normalised factors."""
    s = np.abs(np.diag(d)) + np.abs(np.sqrt(sigma) * rng.standard_normal((k, k)))
    csf, csg = f.sum(axis=0), g.sum(axis=0)
    s = s * (csf * csg)[None, :]
    f = f / csf[None, :]
    g = g / csg[None, :]
    return f, s, g, f.sum(axis=0), g.sum(axis=0)


def native_route():
    """The matrix-sized steps run on the library alone (native_route.py) unless RESNMTF_ROUTE=torch asks for the first
    version of these steps (library GEMM / eigensolver / gathers through torch; kept for comparison runs)."""
    import os

    return os.environ.get("RESNMTF_ROUTE", "native") != "torch"


def device_route(data):
    """Matrix-sized views (>= 250k entries each) keep every matrix-sized step on the GPU -- initialisation, shuffles,
    sub-samples, bisilhouette distance blocks, JSD thresholds, all behind the C ABI
    (native_route.py); small ones take the reference's host route around the device loop."""
    if min(int(np.prod(m.shape)) for m in data) < 250_000:
        return False
    if native_route():
        return L.device_count() > 0
    return _torch_cuda() is not None


def _pool_devices(ctx, use_parallel):
    """GPUs of a call: the context's GPU first, then -- with use_parallel -- the other visible ones, at most
    RESNMTF_MAX_GPUS (0: all)."""
    import os

    if not use_parallel:
        return [ctx.device]
    n = L.device_count()
    cap = int(os.environ.get("RESNMTF_MAX_GPUS", "0") or 0)
    devs = [ctx.device] + [d for d in range(n) if d != ctx.device]
    return devs[:cap] if cap > 0 else devs


def _shuffle_refit_device(torch, xts, k, rng, ctx, max_iters=0):
    """One repeat of obtain_shuffled_f (R/obtain_bicl.r:33-40) with every matrix-sized step on the device (SURVEY 8f
    row N2): per view the full random permutation of the entries (shuffle_view, :11-22, incl. its rejection of
    all-zero rows / columns), the re-normalisation apply_resnmtf applies to the shuffled data (check_inputs,
    R/utils.r:20-27, 86-88), the SVD initialisation and the fit itself; only the n x k factors come back.  The shuffle
    refits run with phi = xi = psi = NULL (R/obtain_bicl.r:35-39), so no shared-index maps are needed.  Randomness:
    one seed per shuffled view and the k x k initialisation noise are drawn from ``rng``; the permutation itself comes
    from a device generator keyed by that seed."""
    dev = xts[0].device
    n_v = len(xts)
    shapes = [(int(x.shape[1]), int(x.shape[0])) for x in xts]  # (n, p)
    with torch.cuda.device(dev):
        messed = []
        for xd in xts:
            gen = torch.Generator(device=dev)
            gen.manual_seed(int(rng.integers(0, 2 ** 62)))
            while True:  # the storage order of xd is R's column-major vector order
                perm = torch.randperm(xd.numel(), generator=gen, device=dev)
                m = xd.reshape(-1)[perm].reshape(xd.shape)
                if not bool((m.sum(dim=0) == 0).any() or (m.sum(dim=1) == 0).any()):
                    break
            del perm
            messed.append(m / m.sum(dim=1, keepdim=True))  # L1 column normalisation (columns of X = rows of m)
        fit = DeviceFit(ctx, [s_[0] for s_ in shapes], [s_[1] for s_ in shapes], [k] * n_v)
        try:
            for v_, m in enumerate(messed):
                u, d, g = _gram_topk_torch(torch, m, k)
                f0, s0, g0, lam, mu = _init_from_svd(u.cpu().numpy(), d.cpu().numpy(), g.cpu().numpy(), k, rng)
                torch.cuda.current_stream(dev).synchronize()  # the library reads m on its own stream
                fit.set_data_device(v_, m.data_ptr(), shapes[v_][0])
                fit.set_factors(v_, f0, s0, g0, lam, mu)
            fit.run(None, 1.0e-6, max_iters)
            fit.normalise()
            return [fit.get_factors(v_)[0] for v_ in range(n_v)]
        finally:
            fit.close()


def shuffled_fits_device(data, n_clusts, num_repeats, rng, ctx, max_iters=0, resident=None):
    """obtain_shuffled_f (R/obtain_bicl.r:31-42) on the device, one repeat after the other: every repeat draws from its
    own child generator of ``rng`` (the pool of apply_resnmtf runs the same repeats as independent units, possibly on
    other GPUs, with the same generators).  Returns None when torch / CUDA is unavailable."""
    torch = _torch_cuda()
    if torch is None:
        return None
    if resident is None:
        dev = torch_device(ctx.device)
        resident = [torch.from_numpy(np.ascontiguousarray((m.x if hasattr(m, "x") else m).T)).to(dev) for m in data]
    return [_shuffle_refit_device(torch, resident, int(n_clusts), r, ctx, max_iters)
            for r in rng.spawn(int(num_repeats))]


def _svd_topk(x, k, device=None):
    """|U_k|, d_k, |V_k| of x.  The reference calls LAPACK's full svd(x) (R/update_steps.r:92); only the
    top-k triplets are used and abs() removes the sign ambiguity.  Small views use the same full SVD on the host;
    larger ones use the Gram matrix of the smaller side (top-k eigenpairs) -- on the GPU when one is there
    (``_svd_topk_device``), else on the host, which is what makes 20000 x 4000 tractable."""
    n, p = x.shape
    if min(n, p) > 512:
        out = _svd_topk_device(x, k, device)
        if out is not None:
            return out
    if min(n, p) <= 1536:
        u, d, vt = np.linalg.svd(x, full_matrices=False)
        return np.abs(u[:, :k]), d[:k], np.abs(vt[:k, :].T)
    from scipy.linalg import eigh

    if p <= n:
        gram = x.T @ x
        w, v = eigh(gram, subset_by_index=[p - k, p - 1])
        w, v = w[::-1], v[:, ::-1]
        d = np.sqrt(np.maximum(w, 0.0))
        u = (x @ v) / d[None, :]
    else:
        gram = x @ x.T
        w, u = eigh(gram, subset_by_index=[n - k, n - 1])
        w, u = w[::-1], u[:, ::-1]
        d = np.sqrt(np.maximum(w, 0.0))
        v = (x.T @ u) / d[None, :]
    return np.abs(u), d, np.abs(v)


def init_mats_inner(x, k_vec, rng, sigma=0.05, device=None):
    """R/update_steps.r:78-125.  The noise term abs(MASS::mvrnorm(k, 0, sigma I_k)) is drawn from ``rng``."""
    fs, ss, gs, lams, mus = [], [], [], [], []
    for i, xi in enumerate(x):
        k = int(k_vec[i])
        f, d, g = _svd_topk(xi, k, device)
        f, s, g, lam, mu = _init_from_svd(f, d, g, k, rng, sigma)
        fs.append(f)
        ss.append(s)
        gs.append(g)
        lams.append(lam)
        mus.append(mu)
    return fs, ss, gs, lams, mus


def init_mats(x, n_v, k_vec, init_f, init_g, init_s, rng, device=None):
    """R/update_steps.r:36-66."""
    if init_f is None or init_g is None or init_s is None:
        return init_mats_inner(x, k_vec, rng, device=device)
    cf = [np.asarray(a, dtype=np.float64) for a in init_f]
    cs = [np.asarray(a, dtype=np.float64) for a in init_s]
    cg = [np.asarray(a, dtype=np.float64) for a in init_g]
    return cf, cs, cg, [a.sum(axis=0) for a in cf], [a.sum(axis=0) for a in cg]


# --------------------------------------------------------------------------------------------------
# one fit = core (initialisation + the loop of R/main.r:50-109 on the device + normalisation) and post-processing
# (obtain_biclusters, R/main.r:122).  The two halves are separate so that the pool can run the cores and the shuffled
# refits of many fits side by side and the post-processing once their inputs exist.
# --------------------------------------------------------------------------------------------------


def _names_or_default(data):
    """Row / column names for the shared-index maps; unnamed views get view-unique placeholders."""
    rn, cn = [], []
    for v, m in enumerate(data):
        rn.append(m.rownames if m.rownames is not None else [f"__v{v}_r{i}" for i in range(m.shape[0])])
        cn.append(m.colnames if m.colnames is not None else [f"__v{v}_c{i}" for i in range(m.shape[1])])
    return rn, cn


def _init_resident(torch, resident, k_vec, rng, cache):
    """init_mats_inner (R/update_steps.r:78-125) on views resident on the GPU; ``cache`` (a dict, or None) keeps the
    triplets of data that is fitted for several k."""
    out = [[], [], [], [], []]
    for v, xt in enumerate(resident):
        k = int(k_vec[v])
        with torch.cuda.device(xt.device):
            if cache is not None and v in cache:
                trip = cache[v]
            else:
                trip = tuple(t.cpu().numpy() for t in _gram_triplets_torch(torch, xt))
                if cache is not None:
                    cache[v] = trip
        u, d, g = trip
        for lst, val in zip(out, _init_from_svd(u[:, :k], d[:k], g[:, :k], k, rng)):
            lst.append(val)
    return tuple(out)


def _fit_core(data, row_indices, column_indices, init_f, init_s, init_g, k_vec, phi, xi, psi, n_iters, *, rng, ctx,
              max_iters=0, err_mode=L.ERR_AUTO, impl=L.IMPL_AUTO, device_data=None, resident=None, eig_cache=None):
    """R/main.r:38-110: initial factors, the update loop (one C-ABI call), normalisation_check."""
    n_v = len(data)
    phi = np.zeros((n_v, n_v)) if phi is None else np.asarray(phi, dtype=np.float64)
    xi = np.zeros((n_v, n_v)) if xi is None else np.asarray(xi, dtype=np.float64)
    psi = np.zeros((n_v, n_v)) if psi is None else np.asarray(psi, dtype=np.float64)
    xs = None
    if resident is not None and (init_f is None or init_g is None or init_s is None):
        cf, cs, cg, clam, cmu = _init_resident(_torch_cuda(), resident, k_vec, rng, eig_cache)
    else:
        xs = [np.asfortranarray(m.x, dtype=np.float64) for m in data] if resident is None else None
        cf, cs, cg, clam, cmu = init_mats(xs, n_v, k_vec, init_f, init_g, init_s, rng, device=ctx.device)
    k_used = [int(f.shape[1]) for f in cf]
    fit = DeviceFit(ctx, [m.shape[0] for m in data], [m.shape[1] for m in data], k_used)
    try:
        fit.set_options(err_mode=err_mode, impl=impl)
        for v in range(n_v):
            if device_data is not None:
                fit.attach_data(v, device_data[v])
            elif resident is not None:
                import torch

                torch.cuda.current_stream(resident[v].device).synchronize()  # the library reads on its own stream
                fit.set_data_device(v, resident[v].data_ptr(), data[v].shape[0])
            else:
                fit.set_data(v, xs[v])  # data_norms (R/main.r:48) are computed on the device
            fit.set_factors(v, cf[v], cs[v], cg[v], clam[v], cmu[v])
        fit.set_restrictions(phi, xi, psi)
        rn, cn = _names_or_default(data)
        for (v, w), (iv, iw) in prep.shared_maps(row_indices, rn).items():
            fit.set_shared_map(L.MAP_ROW, v, w, iv, iw)
        for (v, w), (iv, iw) in prep.shared_maps(column_indices, cn).items():
            fit.set_shared_map(L.MAP_COL, v, w, iv, iw)
        fit.run(n_iters, 1.0e-6, max_iters)  # the loop of R/main.r:50-109
        total_err = fit.errors()
        lam_mu = [fit.get_factors(v)[3:] for v in range(n_v)]
        fit.normalise()  # normalisation_check, R/main.r:110
        outs = [fit.get_factors(v) for v in range(n_v)]
        counters = fit.counters()
    finally:
        fit.close()
    return {"output_f": [o[0] for o in outs], "output_s": [o[1] for o in outs], "output_g": [o[2] for o in outs],
            "total_err": total_err, "lambda": [lm[0] for lm in lam_mu], "mu": [lm[1] for lm in lam_mu],
            "counters": counters}


def _fit_post(core, data, n_iters, num_repeats, spurious, distance, no_clusts, *, rng, ctx, shuffled_f=None,
              resident=None, want_bisil=True):
    """R/main.r:111-139: the result list of res_nmtf_inner from the normalised factors."""
    if no_clusts:
        return {"output_f": core["output_f"], "output_s": core["output_s"], "output_g": core["output_g"]}
    clusters = obtain_biclusters(data, core["output_f"], core["output_g"], core["output_s"], num_repeats, spurious,
                                 distance, rng=rng, ctx=ctx, shuffled_f=shuffled_f, resident=resident,
                                 want_bisil=want_bisil)
    total_err = core["total_err"]
    error = float(np.mean(total_err[-10:])) if n_iters is None else float(total_err[-1])
    return {
        "output_f": core["output_f"], "output_s": core["output_s"], "output_g": core["output_g"],
        "Error": error, "All_Error": total_err, "bisil": clusters["bisil"],
        "row_clusters": clusters["row_clustering"], "col_clusters": clusters["col_clustering"],
        "lambda": core["lambda"], "mu": core["mu"],
        # not in the reference's list: names (R carries them as dimnames) and device counters
        "row_names": [m.rownames for m in data], "col_names": [m.colnames for m in data],
        "counters": core["counters"],
    }


def run_fits(pool, specs, phi, xi, psi, n_iters, num_repeats, spurious, distance, no_clusts, max_iters=0,
             err_mode=L.ERR_AUTO, impl=L.IMPL_AUTO):
    """Runs the res_nmtf_inner calls described by ``specs`` on the pool and returns their result lists in order.

    A spec: dict(key=<data key placed on the pool>, data=[views], row_indices, col_indices, k_vec, rng, and optionally
    init_f / init_s / init_g, shared=<the data is fitted by several specs>, want_bisil).  Units: the core of every spec
    and -- when spurious biclusters are to be removed -- its ``num_repeats`` shuffled refits, all independent
    (R/obtain_bicl.r:31-42 runs them inside the fit; they depend on k and on the data only); then, per spec, the
    post-processing (JSD thresholds, binarisation, bisilhouette), which waits for exactly those."""
    need_shuffles = bool(spurious) and not no_clusts
    tasks, core_at, post_at = [], [], []
    for si, sp in enumerate(specs):
        data, key = sp["data"], sp["key"]
        on_dev = key in pool.loaders
        k_vec = [int(k) for k in sp["k_vec"]]
        cost = sharding.fit_cost([m.shape for m in data], max(k_vec))
        shuffle_rngs = sp["rng"].spawn(int(num_repeats)) if need_shuffles else []

        def core(worker, sp=sp, data=data, key=key, on_dev=on_dev, k_vec=k_vec):
            shared = sp.get("shared", False)
            return _fit_core(data, sp["row_indices"], sp["col_indices"], sp.get("init_f"), sp.get("init_s"),
                             sp.get("init_g"), k_vec, phi, xi, psi, n_iters, rng=sp["rng"], ctx=worker.ctx,
                             max_iters=max_iters, err_mode=err_mode, impl=impl,
                             device_data=worker.get_handles(key) if shared else None,
                             resident=worker.get_views(key) if on_dev else None,
                             eig_cache=worker.eig_cache(key) if (shared and on_dev) else None)

        core_at.append(len(tasks))
        tasks.append((cost * 1.01, core, (), "fit", key if sp.get("shared", False) else None))
        shuffle_at = []
        for srng in shuffle_rngs:
            def shuffle(worker, data=data, key=key, on_dev=on_dev, srng=srng,
                        k=k_vec[0] if sp.get("init_f") is None else int(np.asarray(sp["init_f"][0]).shape[1])):
                from .bicluster import shuffle_refit

                return shuffle_refit(data, k, srng, worker.ctx, resident=worker.get_views(key) if on_dev else None,
                                     max_iters=max_iters)

            shuffle_at.append(len(tasks))
            tasks.append((cost, shuffle, (), "shuffled refit"))

        def post(worker, sp=sp, on_dev=on_dev, ci=core_at[-1], shuffle_at=tuple(shuffle_at)):
            return _fit_post(worker.results[ci], sp["data"], n_iters, num_repeats, spurious, distance, no_clusts,
                             rng=sp["rng"], ctx=worker.ctx,
                             shuffled_f=[worker.results[i] for i in shuffle_at] if need_shuffles else None,
                             resident=worker.get_views(sp["key"]) if (on_dev and not no_clusts) else None,
                             want_bisil=sp.get("want_bisil", True))

        # the post-processing of a fit is ready as soon as its core and its shuffled refits are; it is cheap next to
        # them but sits on the critical path of the call, so a free GPU takes it before starting another fit
        post_at.append(len(tasks))
        tasks.append((cost * 2.0, post, (core_at[-1],) + tuple(shuffle_at), "post-processing"))
    done = pool.run(tasks)
    return [done[i] for i in post_at]


def _single_pool(ctx):
    from .fitpool import FitPool

    return FitPool([ctx])


def _place(pool, key, data):
    """Puts the views of ``data`` on every GPU of the pool (resident, for matrix-sized views) or registers them as
    host data (small views: every fit uploads through the library)."""
    if device_route(data):
        pool.place_resident(key, data)
    else:
        pool.place_host(key, data)


# --------------------------------------------------------------------------------------------------
# res_nmtf_inner
# --------------------------------------------------------------------------------------------------


def res_nmtf_inner(data, row_indices, column_indices, init_f=None, init_s=None, init_g=None, k_vec=None,
                   phi=None, xi=None, psi=None, n_iters=None, num_repeats=5, spurious=True,
                   distance="euclidean", no_clusts=False, *, rng=None, ctx=None, max_iters=0,
                   err_mode=L.ERR_AUTO, impl=L.IMPL_AUTO):
    """R/main.r:32-140.  ``data``: list of (already prepped) views; ``row_indices`` / ``column_indices``:
    per view a dict {other view: shared names or None}, as produced by ``prep.reorder_data``."""
    rng = np.random.default_rng() if rng is None else rng
    ctx = default_context() if ctx is None else ctx
    data = [as_named(m) for m in data]
    if native_route() and device_route(data):
        from .native_route import NativeRunner

        with NativeRunner(devices=[ctx.device]) as runner:
            runner.put(data, prep_values=False)
            spec = dict(data=data, row_indices=row_indices, col_indices=column_indices, k_vec=k_vec, rng=rng,
                        init_f=init_f, init_s=init_s, init_g=init_g)
            return runner.run_fits([spec], phi, xi, psi, n_iters, num_repeats, spurious, distance, no_clusts,
                                   max_iters=max_iters, err_mode=err_mode, impl=impl)[0]
    with _single_pool(ctx) as pool:
        _place(pool, "data", data)
        spec = dict(key="data", data=data, row_indices=row_indices, col_indices=column_indices, k_vec=k_vec, rng=rng,
                    init_f=init_f, init_s=init_s, init_g=init_g)
        if k_vec is None:  # only legal with explicit initial factors: k comes from them
            spec["k_vec"] = [np.asarray(f).shape[1] for f in init_f]
        return run_fits(pool, [spec], phi, xi, psi, n_iters, num_repeats, spurious, distance, no_clusts,
                        max_iters=max_iters, err_mode=err_mode, impl=impl)[0]


# --------------------------------------------------------------------------------------------------
# apply_resnmtf
# --------------------------------------------------------------------------------------------------


def _apply_native(data, reordering, init_f, init_s, init_g, k_vec, phi, xi, psi, n_iters, k_min, k_max, distance, spurious,
                  num_repeats, no_clusts, sample_rate, n_stability, stability, stab_thres, remove_unstable, devices, rng,
                  max_iters):
    """R/main.r:258-334 for matrix-sized views on the native pool (native_route.py): the raw views go to the first GPU
    once (prepped there), every convergence loop of the call is a unit of resnmtf_batch_run."""
    from .native_route import NativeRunner, ResidentView

    n_v = len(data)
    with NativeRunner(devices=devices) as runner:
        runner.put(data, prep_values=True)
        data = [ResidentView(m.shape, m.rownames, m.colnames) for m in data]  # the values live on the GPUs only
        base = dict(data=data, row_indices=reordering["row_indices"], init_f=init_f, init_s=init_s, init_g=init_g)
        run = lambda specs: runner.run_fits(specs, phi, xi, psi, n_iters, num_repeats, spurious, distance, no_clusts,  # noqa: E731
                                            max_iters=max_iters)
        stab = lambda results, k: runner.stability_check(data, results, k, phi, xi, psi, n_iters, spurious, num_repeats,  # noqa: E731
                                                         no_clusts, distance, sample_rate, n_stability, stab_thres,
                                                         remove_unstable, rng=rng, max_iters=max_iters)
        if k_vec is not None:
            results = run([dict(base, col_indices=reordering["col_indices"], k_vec=k_vec, rng=rng)])[0]
            return stab(results, k_vec) if stability else results
        ks = list(range(int(k_min), int(k_max) + 1))
        child_rngs = rng.spawn(len(ks))
        res_list = run([dict(base, col_indices=reordering["col_indices"], k_vec=[k] * n_v, rng=child_rngs[i])
                        for i, k in enumerate(ks)])
        err_list = extract_bisils(res_list, ks)
        test = ks[int(np.argmax(err_list))]
        max_k = int(k_max)
        if k_min != k_max:
            while test == max_k:
                max_k += 1
                ks.append(max_k)
                # quirk Q3 (R/main.r:312): the extension fits run with column_indices = NULL.  Reproduced.
                res_list.append(run([dict(base, col_indices=None, k_vec=[max_k] * n_v, rng=rng)])[0])
                err_list.append(res_list[-1]["bisil"])
                test = ks[int(np.argmax(err_list))]
        best = int(np.argmax(err_list))
        results = res_list[best]
        return stab(results, ks[best]) if stability else results


def extract_bisils(res_list, k_vec):
    """R/utils.r:203-210."""
    return [res_list[i]["bisil"] for i in range(len(k_vec))]


def apply_resnmtf(data, init_f=None, init_s=None, init_g=None, k_val=None, phi=None, xi=None, psi=None,
                  n_iters=None, k_min=3, k_max=8, distance="euclidean", spurious=True, num_repeats=5,
                  no_clusts=False, sample_rate=0.9, n_stability=5, stability=True, stab_thres=0.4,
                  remove_unstable=True, use_parallel=True, *, rng=None, ctx=None, max_iters=0):
    """R/main.r:214-335."""
    from .fitpool import FitPool

    rng = np.random.default_rng() if rng is None else rng
    ctx = default_context() if ctx is None else ctx
    if not isinstance(data, (list, tuple)):
        # quirk Q1 of the reference (length(data) is taken before the matrix -> list wrap) is not
        # reproduced: a bare matrix is treated as the single-view list the documentation promises
        data = [data]
    n_v = len(data)
    k_vec = None if k_val is None else [int(np.atleast_1d(k_val)[0])] * n_v
    phi = None if phi is None else np.asarray(phi, dtype=np.float64)
    xi = None if xi is None else np.asarray(xi, dtype=np.float64)
    psi = None if psi is None else np.asarray(psi, dtype=np.float64)
    named = prep.give_names(data, n_v, phi, psi)
    reordering = prep.reorder_data(named["data"], n_v, named["row_names"], named["col_names"])
    phi = prep.init_rest_mats(phi, n_v)
    psi = prep.init_rest_mats(psi, n_v)
    xi = prep.init_rest_mats(xi, n_v)
    on_device = device_route(named["data"])  # matrix-sized views are prepped on the GPU on the way in
    data = prep.check_inputs(named["data"], init_f, init_s, init_g, k_vec, phi, xi, psi, n_iters, k_min, k_max,
                             distance, num_repeats, no_clusts, sample_rate, n_stability, stability, stab_thres,
                             remove_unstable, spurious, prep_values=not on_device)
    if k_vec is None and no_clusts:
        # the reference fails here too: results carry no bisil, which.max(NULL) is integer(0) and the
        # `while (test == max_k)` at R/main.r:307 stops with "argument is of length zero"
        raise ValueError("argument is of length zero")
    # The fits of one call are independent units: every fit gets its own child generator -- so the result
    # does not depend on how many GPUs run them -- and, with use_parallel and more than one visible GPU, they are
    # spread over one context per GPU (the reference's %dopar% intent, R/main.r:288-299, which is unreachable there).
    if on_device and native_route():
        return _apply_native(data, reordering, init_f, init_s, init_g, k_vec, phi, xi, psi, n_iters, k_min, k_max, distance,
                             spurious, num_repeats, no_clusts, sample_rate, n_stability, stability, stab_thres,
                             remove_unstable, _pool_devices(ctx, use_parallel), rng, max_iters)
    contexts = device_contexts(ctx) if use_parallel else [ctx]
    fit_kw = dict(max_iters=max_iters)
    with FitPool(contexts) as pool:
        if on_device:
            from .fitpool import ResidentMatrix

            pool.place_resident("data", data, prep=True)
            data = [ResidentMatrix(m.shape, m.rownames, m.colnames) for m in data]  # the values live on the GPUs only
        else:
            pool.place_host("data", data)
        base = dict(key="data", data=data, row_indices=reordering["row_indices"], init_f=init_f, init_s=init_s,
                    init_g=init_g, shared=True)
        if k_vec is not None:
            results = run_fits(pool, [dict(base, col_indices=reordering["col_indices"], k_vec=k_vec, rng=rng)], phi, xi,
                               psi, n_iters, num_repeats, spurious, distance, no_clusts, **fit_kw)[0]
            if stability:
                results = stability_check(data, results, k_vec, phi, xi, psi, n_iters, spurious, num_repeats,
                                          no_clusts, distance, sample_rate, n_stability, stab_thres, rng=rng,
                                          pool=pool, max_iters=max_iters)
            return results
        ks = list(range(int(k_min), int(k_max) + 1))
        child_rngs = rng.spawn(len(ks))
        res_list = run_fits(pool, [dict(base, col_indices=reordering["col_indices"], k_vec=[k] * n_v, rng=child_rngs[i])
                                   for i, k in enumerate(ks)],
                            phi, xi, psi, n_iters, num_repeats, spurious, distance, no_clusts, **fit_kw)
        err_list = extract_bisils(res_list, ks)
        test = ks[int(np.argmax(err_list))]
        max_k = int(k_max)
        if k_min != k_max:
            while test == max_k:
                max_k += 1
                ks.append(max_k)
                # quirk Q3 (R/main.r:312): `reordering$column_indices` does not exist, so the extension fits run
                # with column_indices = NULL -- nothing is overwritten in the psi coupling.  Reproduced.
                res_list.append(run_fits(pool, [dict(base, col_indices=None, k_vec=[max_k] * n_v, rng=rng)], phi, xi, psi,
                                         n_iters, num_repeats, spurious, distance, no_clusts, **fit_kw)[0])
                err_list.append(res_list[-1]["bisil"])
                test = ks[int(np.argmax(err_list))]
        best = int(np.argmax(err_list))
        results = res_list[best]
        if stability:
            results = stability_check(data, results, ks[best], phi, xi, psi, n_iters, spurious, num_repeats,
                                      no_clusts, distance, sample_rate, n_stability, stab_thres, remove_unstable,
                                      rng=rng, pool=pool, max_iters=max_iters)
        return results
