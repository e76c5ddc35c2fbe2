"""Host post-processing of a fit, mirroring R/obtain_bicl.r (binarisation, S-based row/column pairing,
spurious-bicluster removal through shuffled refits and JSD thresholds, bisilhouette aggregation).

These run on n x k / p x k outputs, not on the hot path.  Two of the reference's
dependencies are not in the reference tree and are restated here from their published definitions --
**parity unpinned**: `bisilhouette::bisilhouette` (Remotes: eso28599/bisilhouette, no version pinned) and
`philentropy::JSD` on `stats::density` estimates.  In the R drop-in the R host keeps calling the real
packages on the (bit-identical) binary matrices, so this module only matters for the Python mirror."""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------------------
# stats::density (gaussian kernel, bw.nrd0, linear binning + FFT, as density.default does it)
# --------------------------------------------------------------------------------------------------


def bw_nrd0(x):
    x = np.asarray(x, dtype=np.float64)
    hi = np.std(x, ddof=1) if x.size > 1 else 0.0
    q75, q25 = np.percentile(x, [75, 25])
    lo = min(hi, (q75 - q25) / 1.34)
    if not lo:
        lo = hi or abs(x[0]) or 1.0
    return 0.9 * lo * x.size ** (-0.2)


def bw_nrd0_columns(cols):
    """bw.nrd0 of every column of ``cols`` at once (the same formulas as ``bw_nrd0``, one pass per statistic)."""
    cols = np.asarray(cols, dtype=np.float64)
    n = cols.shape[0]
    hi = np.std(cols, axis=0, ddof=1) if n > 1 else np.zeros(cols.shape[1])
    # along the contiguous axis of the column-major matrix: same values, half the time of axis=0 on this layout
    if cols.flags.f_contiguous:
        q75, q25 = np.percentile(cols.T, [75, 25], axis=1)
    else:
        q75, q25 = np.percentile(cols, [75, 25], axis=0)
    lo = np.minimum(hi, (q75 - q25) / 1.34)
    first = np.abs(cols[0, :])
    lo = np.where(lo != 0, lo, np.where(hi != 0, hi, np.where(first != 0, first, 1.0)))
    return 0.9 * lo * n ** (-0.2)


def r_density(x, from_=None, to=None, n=512, cut=3.0):
    """(x grid, density) like stats::density(x, from=, to=) with the default gaussian kernel."""
    x = np.asarray(x, dtype=np.float64)
    N = x.size
    bw = bw_nrd0(x)
    if from_ is None:
        from_ = x.min() - cut * bw
    if to is None:
        to = x.max() + cut * bw
    n_user = n
    n = max(n, 512)
    if n > 512:
        n = 2 ** int(np.ceil(np.log2(n)))
    lo, up = from_ - 4 * bw, to + 4 * bw
    delta = (up - lo) / (n - 1)
    xpos = (x - lo) / delta
    ix = np.floor(xpos).astype(np.int64)
    fx = xpos - ix
    w = 1.0 / N
    inside = (ix >= 0) & (ix <= n - 2)
    left = ix == -1
    right = ix == n - 1
    # linear binning (BinDist): one ordered accumulation -- np.bincount adds in input order, i.e. exactly the sequence
    # of four np.add.at passes this replaces (bit-identical), at a fraction of their cost
    idx = np.concatenate([ix[inside], ix[inside] + 1, np.zeros(int(left.sum()), dtype=np.int64), ix[right]])
    wts = np.concatenate([w * (1 - fx[inside]), w * fx[inside], w * fx[left], w * (1 - fx[right])])
    y = np.bincount(idx, weights=wts, minlength=2 * n).astype(np.float64)
    kords = np.linspace(0, 2 * (up - lo), 2 * n)
    kords[n + 1:2 * n] = -kords[n - 1:0:-1]
    kords = np.exp(-0.5 * (kords / bw) ** 2) / (bw * np.sqrt(2 * np.pi))
    conv = np.fft.ifft(np.fft.fft(y) * np.conj(np.fft.fft(kords))).real  # numpy's ifft divides by len
    dens = np.maximum(0.0, conv[:n])
    xords = np.linspace(lo, up, n)
    xout = np.linspace(from_, to, n_user)
    return xout, np.interp(xout, xords, dens)


def jsd(p, q):
    """philentropy::JSD(rbind(p, q), unit='log2', est.prob='empirical')."""
    p = np.asarray(p, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    p = p / p.sum()
    q = q / q.sum()
    m = 0.5 * (p + q)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.where(p > 0, p * np.log2(p / m), 0.0)
        b = np.where(q > 0, q * np.log2(q / m), 0.0)
    return 0.5 * a.sum() + 0.5 * b.sum()


def jsd_calc(x1, x2):
    """R/utils.r:95-106."""
    max_val = max(np.max(x1), np.max(x2))
    gx, d1 = r_density(x1, 0.0, max_val)
    _, d2 = r_density(x2, 0.0, max_val)
    d1 = np.where(gx > np.max(x1), 0.0, d1)
    d2 = np.where(gx > np.max(x2), 0.0, d2)
    return jsd(d1, d2)


# --------------------------------------------------------------------------------------------------
# bisilhouette (restated; parity unpinned)
# --------------------------------------------------------------------------------------------------


def _pairwise(a, b, method):
    if method == "euclidean":  # the direct form sqrt(sum (a - b)^2), as stats::dist computes it (no Gram-form cancellation)
        from scipy.spatial.distance import cdist

        return cdist(a, b, "euclidean")
    if method == "manhattan":
        from scipy.spatial.distance import cdist

        return cdist(a, b, "cityblock")
    if method == "cosine":
        na = np.linalg.norm(a, axis=1)[:, None]
        nb = np.linalg.norm(b, axis=1)[None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            c = (a @ b.T) / (na * nb)
        return 1.0 - np.nan_to_num(c)
    raise ValueError("distance must be one of 'euclidean', 'manhattan' or 'cosine'.")


def _pairwise_torch(torch, a, b, method):
    if method == "euclidean":
        d2 = (a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * (a @ b.T)
        return torch.sqrt(torch.clamp(d2, min=0.0))
    if method == "manhattan":
        return torch.cdist(a, b, p=1.0)
    if method == "cosine":
        na = torch.linalg.norm(a, dim=1)[:, None]
        nb = torch.linalg.norm(b, dim=1)[None, :]
        return 1.0 - torch.nan_to_num((a @ b.T) / (na * nb), nan=0.0, posinf=0.0, neginf=0.0)
    raise ValueError("distance must be one of 'euclidean', 'manhattan' or 'cosine'.")


def bisilhouette_device(data, row_clustering, col_clustering, method="euclidean", device=None, xt=None):
    """``bisilhouette`` with the pairwise-distance work on the GPU: the same definition, the same
    formulas, FP64 throughout; the |R_k| x |R_l| distance blocks are Gram contractions (library GEMM through torch --
    this is post-processing of a finished fit, not the update path).  On the k-sweep of the 20000 x 4000 view the
    host version is 53 of 59 s of wall time.  Returns None when torch / CUDA is unavailable."""
    from .api import _torch_cuda

    torch = _torch_cuda()
    if torch is None:
        return None
    if xt is not None:  # the view is already resident (p x n row-major == column-major n x p): no upload
        dev = xt.device
        x = xt.T
    else:
        from .api import torch_device

        dev = torch_device(device)
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(data, dtype=np.float64))).to(dev)
    rc = np.asarray(row_clustering) > 0
    cc = np.asarray(col_clustering) > 0
    k = rc.shape[1]
    live = [j for j in range(k) if rc[:, j].any() and cc[:, j].any()]
    vals = []
    for j in live:
        rows = np.flatnonzero(rc[:, j])
        sub = x[:, torch.from_numpy(np.flatnonzero(cc[:, j])).to(dev)]
        mine = sub[torch.from_numpy(rows).to(dev)]
        if len(rows) > 1:
            d_in = _pairwise_torch(torch, mine, mine, method)
            d_in.fill_diagonal_(0.0)  # d(i, i) = 0 exactly (the Gram form leaves sqrt(rounding noise) there)
            a = d_in.sum(1) / max(len(rows) - 1, 1)
        else:
            a = torch.zeros(len(rows), dtype=torch.float64, device=dev)
        others = []
        for l in live:
            if l == j:
                continue
            rl = np.flatnonzero(rc[:, l] & ~rc[:, j])
            if rl.size:
                others.append(rl)
        if not others:
            rest = np.flatnonzero(~rc[:, j])
            if rest.size:
                others.append(rest)
        if not others:
            vals.append(0.0)
            continue
        b = torch.stack([_pairwise_torch(torch, mine, sub[torch.from_numpy(rl).to(dev)], method).mean(1)
                         for rl in others]).min(dim=0).values
        den = torch.maximum(a, b)
        s_ = torch.where(den > 0, (b - a) / den, torch.zeros_like(den))
        vals.append(float(s_.mean().item()))
    return {"bisil": float(np.mean(vals)) if vals else 0.0, "vals": vals}


def bisilhouette(data, row_clustering, col_clustering, method="euclidean"):
    """Bisilhouette score of a biclustering (Orme et al.): for every non-empty bicluster (R_k, C_k) the
    silhouette of the rows of R_k computed on the columns C_k only, against the other row clusters (or, when
    there is no other non-empty cluster, against the rows outside R_k); the score is the mean over
    biclusters of the mean row silhouette.  Returns dict(bisil=..., vals=[per-bicluster])."""
    x = np.asarray(data, dtype=np.float64)
    rc = np.asarray(row_clustering) > 0
    cc = np.asarray(col_clustering) > 0
    k = rc.shape[1]
    live = [j for j in range(k) if rc[:, j].any() and cc[:, j].any()]
    vals = []
    for j in live:
        rows = np.flatnonzero(rc[:, j])
        sub = x[:, cc[:, j]]
        d_in = _pairwise(sub[rows], sub[rows], method)
        np.fill_diagonal(d_in, 0.0)  # d(i, i) = 0 exactly (the Gram form leaves sqrt(rounding noise) there)
        a = d_in.sum(1) / max(len(rows) - 1, 1) if len(rows) > 1 else np.zeros(len(rows))
        others = []
        for l in live:
            if l == j:
                continue
            rl = np.flatnonzero(rc[:, l] & ~rc[:, j])
            if rl.size:
                others.append(rl)
        if not others:
            rest = np.flatnonzero(~rc[:, j])
            if rest.size:
                others.append(rest)
        if not others:
            vals.append(0.0)
            continue
        b = np.min(np.stack([_pairwise(sub[rows], sub[rl], method).mean(1) for rl in others]), axis=0)
        den = np.maximum(a, b)
        with np.errstate(divide="ignore", invalid="ignore"):
            s = np.where(den > 0, (b - a) / den, 0.0)
        vals.append(float(s.mean()))
    return {"bisil": float(np.mean(vals)) if vals else 0.0, "vals": vals}


# --------------------------------------------------------------------------------------------------
# R/obtain_bicl.r
# --------------------------------------------------------------------------------------------------


def shuffle_view(x_i, rng):
    """R/obtain_bicl.r:11-22: permute all entries until no row / column is all zero."""
    x = np.asarray(x_i, dtype=np.float64)
    while True:
        m = rng.permutation(x.ravel(order="F")).reshape(x.shape, order="F")
        if not ((m.sum(0) == 0).any() or (m.sum(1) == 0).any()):
            return m


def shuffle_refit(data, n_clusts, rng, ctx, resident=None, max_iters=0):
    """One repeat of obtain_shuffled_f (R/obtain_bicl.r:33-40): shuffle every view, refit with k = n_clusts and no
    restrictions, return the F factors.  ``resident``: the views as device tensors -- then the shuffle, the
    re-normalisation, the initialisation and the fit never leave the GPU; otherwise the reference's
    route literally (shuffle on the host, apply_resnmtf with k_val, no_clusts, no stability)."""
    from .api import _shuffle_refit_device, _torch_cuda, apply_resnmtf

    if resident is not None:
        return _shuffle_refit_device(_torch_cuda(), resident, int(n_clusts), rng, ctx, max_iters)
    messed = [shuffle_view(m.x if hasattr(m, "x") else m, rng) for m in data]
    return apply_resnmtf(messed, k_val=n_clusts, no_clusts=True, stability=False, rng=rng, ctx=ctx,
                         use_parallel=False, max_iters=max_iters)["output_f"]


def obtain_shuffled_f(data, n_views, num_repeats, n_clusts, rng, ctx, resident=None):
    """R/obtain_bicl.r:31-42, one repeat after the other.  Every repeat draws from its own child generator of ``rng``;
    apply_resnmtf's pool runs the same repeats, with the same generators, as independent units on any GPU."""
    return [shuffle_refit(data, n_clusts, r, ctx, resident) for r in rng.spawn(int(num_repeats))]


def calculate_f_shuffle_jsd(f_mess, i, j, num_repeats, n_clusts):
    """R/obtain_bicl.r:55-68 (j is 0-based here)."""
    scores = []
    for k in range(n_clusts):
        x1 = f_mess[j][i][:, k]
        for l in range(j + 1, num_repeats):
            for m in range(n_clusts):
                scores.append(jsd_calc(x1, f_mess[l][i][:, m]))
    return scores


def _jsd_scores_device(f_mess, output_f_i, i, num_repeats, n_clusts, ctx):
    """All jsd_calc values the spurious-bicluster test needs for view i -- the pairs of calculate_f_shuffle_jsd
    (R/obtain_bicl.r:55-68, as get_thresholds loops over it, :87-92) and of check_biclusters (:122-129) -- in ONE
    launch of the library's pair kernel (resnmtf_jsd_pairs).  Column layout of the batch: the fitted
    factor first, then the shuffled factors repeat by repeat.  Returns (threshold scores in the reference's order,
    scores[k] = mean JSD of fitted column k against every shuffled column)."""
    from .device import jsd_pairs

    k = int(n_clusts)
    cols = np.concatenate([np.asarray(output_f_i, dtype=np.float64)] +
                          [np.asarray(f_mess[r][i], dtype=np.float64) for r in range(num_repeats)], axis=1)
    cols = np.asfortranarray(cols)
    bw = bw_nrd0_columns(cols)
    vmax = cols.max(axis=0)
    col_of = lambda r, c: k + r * k + c  # noqa: E731 - column of shuffled repeat r, cluster c
    pa, pb = [], []
    for j in range(max(num_repeats - 1, 1)):
        for kk in range(k):
            for l in range(j + 1, num_repeats):
                for m in range(k):
                    pa.append(col_of(j, kk))
                    pb.append(col_of(l, m))
    n_thr = len(pa)
    noise = [col_of(j, c) for j in range(max(num_repeats - 1, 1)) for c in range(k)]
    noise += [col_of(num_repeats - 1, c) for c in range(k)]
    for kk in range(k):
        for c in noise:
            pa.append(kk)
            pb.append(c)
    vals = jsd_pairs(ctx, cols, bw, vmax, pa, pb)
    return list(vals[:n_thr]), vals[n_thr:].reshape(k, len(noise)).mean(axis=1)


def get_thresholds(x, output_f, num_repeats, n_views, n_clusts, rng, ctx, shuffled_f=None, resident=None):
    """R/obtain_bicl.r:80-102.  ``shuffled_f``: the shuffled refits when the caller already has them."""
    f_mess = shuffled_f if shuffled_f is not None else obtain_shuffled_f(x, n_views, num_repeats, n_clusts, rng, ctx,
                                                                         resident)
    on_device = ctx is not None and resident is not None
    avg_score, max_score, shuffled, direct = [], [], [], []
    for i in range(n_views):
        scores, cols = [], []
        for j in range(max(num_repeats - 1, 1)):
            cols.append(f_mess[j][i])
            if not on_device:
                scores += calculate_f_shuffle_jsd(f_mess, i, j, num_repeats, n_clusts)
        cols.append(f_mess[num_repeats - 1][i])
        shuffled.append(np.concatenate(cols, axis=1))
        if on_device:
            scores, per_cluster = _jsd_scores_device(f_mess, output_f[i], i, num_repeats, n_clusts, ctx)
            direct.append(per_cluster)
        avg_score.append(float(np.mean(scores)))
        gx, gy = r_density(np.asarray(scores))
        max_score.append(float(gx[int(np.argmax(gy))]))
    return {"avg_score": avg_score, "max_score": max_score, "shuffled_f": shuffled,
            "score": np.stack(direct) if on_device else None}


def check_biclusters(data, output_f, num_repeats, rng, ctx, shuffled_f=None, resident=None):
    """R/obtain_bicl.r:113-133."""
    n_views = len(data)
    n_clusts = output_f[0].shape[1]
    th = get_thresholds(data, output_f, num_repeats, n_views, n_clusts, rng, ctx, shuffled_f, resident)
    scores = th["score"]
    if scores is None:
        scores = np.zeros((n_views, n_clusts))
        for i in range(n_views):
            noise = th["shuffled_f"][i]
            for k in range(n_clusts):
                xk = output_f[i][:, k]
                scores[i, k] = np.mean([jsd_calc(xk, noise[:, c]) for c in range(noise.shape[1])])
    return {"score": scores, "avg_threshold": th["avg_score"], "max_threshold": th["max_score"]}


def binarise(output_f, output_g):
    """R/obtain_bicl.r:162-173: 1[F > 1/n], 1[G > 1/p] as 0/1 float matrices (bit-exact bar of north_star)."""
    rows = [(f > (1.0 / f.shape[0])).astype(np.float64) for f in output_f]
    cols = [(g > (1.0 / g.shape[0])).astype(np.float64) for g in output_g]
    return rows, cols


def obtain_biclusters(data, output_f, output_g, output_s, num_repeats, remove_spurious=True,
                      distance="euclidean", rng=None, ctx=None, shuffled_f=None, resident=None, want_bisil=True):
    """R/obtain_bicl.r:151-204.  ``shuffled_f`` / ``resident``: shuffled refits already computed by the caller / the
    views as device tensors (both None: the reference's serial route).  ``want_bisil=False`` skips the bisilhouette
    of callers that never read it (the resample fits of the stability analysis use the clusters only)."""
    import os as _os
    import sys as _sys
    import time as _time

    _trace = _os.environ.get("RESNMTF_TRACE", "0") not in ("", "0")
    _t0 = _time.perf_counter()
    n_views = len(output_f)
    biclusts = (check_biclusters(data, output_f, num_repeats, rng, ctx, shuffled_f, resident)
                if remove_spurious else None)
    _t1 = _time.perf_counter()
    row_clustering, col_clustering = binarise(output_f, output_g)
    bisil = []
    for i in range(n_views):
        relations = np.argmax(output_s[i], axis=0)  # which.max per column: first maximum
        row_clustering[i] = row_clustering[i][:, relations]
        if remove_spurious:
            indices = (biclusts["score"][i] < biclusts["max_threshold"][i]) | (biclusts["score"][i] == 0)
            new_indices = indices[relations]
            row_clustering[i][:, new_indices] = 0.0
            col_clustering[i][:, new_indices] = 0.0
        if not want_bisil:
            bisil.append(0.0)
            continue
        score = None
        xi = data[i].x if hasattr(data[i], "x") else data[i]
        if resident is not None and hasattr(resident[i], "bisil"):
            # matrix-sized view resident in the library's layout: distance blocks by resnmtf_data_bisil
            score = resident[i].bisil(row_clustering[i], col_clustering[i], method=distance)
        elif resident is not None and resident[i] is not None:  # first version: a torch tensor of the view
            score = bisilhouette_device(None, row_clustering[i], col_clustering[i], method=distance, xt=resident[i])
        elif xi is not None and int(np.prod(xi.shape)) >= 250_000:
            score = bisilhouette_device(xi, row_clustering[i], col_clustering[i], method=distance,
                                        device=getattr(ctx, "device", None))
        if score is None:
            score = bisilhouette(xi, row_clustering[i], col_clustering[i], method=distance)
        bisil.append(score["bisil"])
    bisil = np.asarray(bisil)
    if _trace:
        print(f"  [resnmtf trace]   post of a k = {output_f[0].shape[1]} fit: spurious test {_t1 - _t0:.3f} s, binarise + "
              f"bisilhouette {_time.perf_counter() - _t1:.3f} s", file=_sys.stderr, flush=True)
    overall = 0.0 if bisil.sum() == 0 else float(bisil[bisil != 0].mean())
    return {"row_clustering": row_clustering, "col_clustering": col_clustering, "bisil": overall}
