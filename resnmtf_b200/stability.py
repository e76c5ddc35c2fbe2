"""Stability analysis, mirroring R/stability_analysis.r: sub-sample rows/columns at ``sample_rate``, refit
with the same k through res_nmtf_inner (i.e. through the device loop), score the new biclusters against
the original ones with the Jaccard relevance, and blank the unstable ones.

R's ``sample()`` is replaced by a NumPy ``Generator`` (``rng.permutation``); the Cartesian-product string
sets of ``cart_prod`` / ``jaccard_func`` (R/utils.r:117-145) are replaced by the equivalent closed form
|A x B  n  C x D| = |A n C| * |B n D|."""
from __future__ import annotations

import numpy as np

from .prep import NamedMatrix, reorder_data


def jaccard_main(row_c, col_c, true_r, true_c, m, n):
    """R/stability_analysis.r:16-33."""
    jac = np.zeros((m, n))
    rc, cc = row_c > 0, col_c > 0
    tr, tc = true_r > 0, true_c > 0
    for i in range(m):
        for j in range(n):
            inter = int((rc[:, i] & tr[:, j]).sum()) * int((cc[:, i] & tc[:, j]).sum())
            size_i = int(rc[:, i].sum()) * int(cc[:, i].sum())
            size_j = int(tr[:, j].sum()) * int(tc[:, j].sum())
            union = size_i + size_j - inter
            jac[i, j] = 0.0 if union == 0 else inter / union
    return jac


def relevance_results(row_c, col_c, true_r, true_c):
    """R/stability_analysis.r:45-67."""
    m, n = row_c.shape[1], true_r.shape[1]
    m_0 = int((row_c.sum(0) != 0).sum())
    n_0 = int((true_r.sum(0) != 0).sum())
    if (m_0 == 0 and n_0 != 0) or (n_0 == 0 and m_0 != 0):
        return 0.0
    if m_0 == 0 and n_0 == 0:
        return 1.0
    return jaccard_main(row_c, col_c, true_r, true_c, m, n).max(axis=0)


def test_cond(data, attempt):
    """R/stability_analysis.r:75-87: TRUE on the first attempt, else whether any view has a zero row/col."""
    if attempt == 1:
        return True
    for x in data:
        if x is None:
            return True
        a = x.x if isinstance(x, NamedMatrix) else np.asarray(x)
        if (a.sum(0) == 0).any() or (a.sum(1) == 0).any():
            return True
    return False


test_cond.__test__ = False  # not a pytest test


def number_biclusters(results):
    """R/stability_analysis.r:94-99.  With no_clusts the result list carries no row_clusters (R/main.r:115-120):
    results$row_clusters is NULL there, lapply over NULL is empty and the count is 0."""
    return float(sum(x.sum() for x in (results.get("row_clusters") or [])))


def _sample(rng, n, size):
    """R's sample(n, size): `size` is truncated towards zero."""
    return rng.permutation(int(n))[: int(size)]


def _subset(view, rows, cols):
    return NamedMatrix(np.asfortranarray(view.x[np.ix_(rows, cols)]),
                       [view.rownames[i] for i in rows], [view.colnames[j] for j in cols])


def host_sums(data):
    """sums(i, rows, cols) -> (column sums, row sums) of data[i][rows, cols] on the host."""
    def sums(i, rows, cols):
        sub = data[i].x[np.ix_(rows, cols)]
        return sub.sum(0), sub.sum(1)

    return sums


def resident_sums(xts):
    """The same on views resident on a GPU (p x n row-major tensors): the sub-sample is gathered and reduced there
    (SURVEY 8f N2), only the two sum vectors come back."""
    import torch

    def sums(i, rows, cols):
        xt = xts[i]
        r = torch.from_numpy(np.asarray(rows, dtype=np.int64)).to(xt.device)
        c = torch.from_numpy(np.asarray(cols, dtype=np.int64)).to(xt.device)
        sub = xt.index_select(0, c).index_select(1, r)
        return sub.sum(dim=1).cpu().numpy(), sub.sum(dim=0).cpu().numpy()

    return sums


def draw_subsample(sums, shapes, dim_1, n_views, sample_rate, rng):
    """The sampling loop of stability_repeat (R/stability_analysis.r:222-253; initial_shuffle :111-132, sample_view
    :157-192): rows / columns sampled at ``sample_rate`` (views with view 1's dimensions reuse its sample), all-zero
    rows / columns of the sub-sample dropped, up to 20 attempts.  Only the column and row sums of a sub-sample are
    ever looked at, so the data is reached through ``sums(i, rows, cols)``.  Returns (row_samples, col_samples), or
    None when no admissible sub-sample was found."""
    row_samples = [None] * n_views
    col_samples = [None] * n_views

    def keep(i):  # (rows with a non-zero sum, columns with a non-zero sum) of the current sub-sample of view i
        cs, rs = sums(i, row_samples[i], col_samples[i])
        return rs != 0, cs != 0

    attempt = 1
    while True:
        if attempt > 1:  # test_cond (:75-87): another attempt while any view has an all-zero row or column
            if not any((~kr).any() or (~kc).any() for kr, kc in (keep(i) for i in range(n_views))):
                break
        if attempt == 20:
            print("Unable to perform stability analysis due to sparsity of data.")
            return None
        row_samples[0] = _sample(rng, dim_1[0], dim_1[0] * sample_rate)
        col_samples[0] = _sample(rng, dim_1[1], dim_1[1] * sample_rate)
        keep_r, keep_c = keep(0)
        if (~keep_r).any() or (~keep_c).any():
            row_samples[0] = row_samples[0][keep_r]
            col_samples[0] = col_samples[0][keep_c]
        for i in range(1, n_views):
            dims = shapes[i]
            # initial_shuffle (R/stability_analysis.r:111-132): reuse view 1's sample when the dims agree
            row_samples[i] = row_samples[0] if dims[0] == dim_1[0] else _sample(rng, dims[0], dims[0] * sample_rate)
            col_samples[i] = col_samples[0] if dims[1] == dim_1[1] else _sample(rng, dims[1], dims[1] * sample_rate)
            keep_r, keep_c = keep(i)
            if (~keep_r).any() or (~keep_c).any():  # sample_view (R/stability_analysis.r:157-192)
                if dims[0] == dim_1[0]:
                    for q in range(i + 1):
                        row_samples[q] = row_samples[q][keep_r]
                else:
                    row_samples[i] = row_samples[i][keep_r]
                if dims[1] == dim_1[1]:
                    for q in range(i + 1):
                        col_samples[q] = col_samples[q][keep_c]
                else:
                    col_samples[i] = col_samples[i][keep_c]
        attempt += 1
    return row_samples, col_samples


def stability_check(data, results, k, phi, xi, psi, n_iters, spurious, num_repeats, no_clusts, distance,
                    sample_rate=0.9, n_stability=5, stab_thres=0.6, remove_unstable=True, rng=None, ctx=None,
                    use_parallel=True, pool=None, max_iters=0):
    """R/stability_analysis.r:302-338 with stability_repeat (:215-278) unrolled over the pool.  ``k`` may be a vector
    (the fixed-k caller passes k_vec; base R's matrix(ncol = k) then uses its first element -- quirk Q9).  The
    ``n_stability`` repeats are independent: each draws its sub-sample and fits from its own child
    generator -- so the result does not depend on how many GPUs run them -- and its fit and shuffled refits are
    units of the pool like those of the k-sweep.  ``pool``: the FitPool of the calling apply_resnmtf with the views
    placed under the key "data"; without one a pool is made from ``ctx`` (one context per GPU when ``use_parallel``)."""
    if number_biclusters(results) == 0:
        print("No biclusters detected!")
        return results
    from .api import _place, run_fits
    from .device import default_context, device_contexts
    from .fitpool import FitPool, ResidentMatrix

    own_pool = pool is None
    if own_pool:
        ctx = default_context() if ctx is None else ctx
        pool = FitPool(device_contexts(ctx) if use_parallel else [ctx])
        _place(pool, "data", data)
    try:
        k = int(np.atleast_1d(k)[0])
        n_views = len(data)
        dim_1 = data[0].shape
        shapes = [m.shape for m in data]
        n_rep = int(n_stability)
        rng = np.random.default_rng() if rng is None else rng
        child_rngs = rng.spawn(n_rep)
        on_dev = "data" in pool.loaders

        def draw(worker, i):
            sums = resident_sums(worker.get_views("data")) if on_dev else host_sums(data)
            return draw_subsample(sums, shapes, dim_1, n_views, sample_rate, child_rngs[i])

        samples = pool.run([(1.0, lambda w, i=i: draw(w, i), (), "sub-sample") for i in range(n_rep)])
        if any(smp is None for smp in samples):
            return results
        specs = []
        for i, (row_s, col_s) in enumerate(samples):
            key = ("sub", i)
            names = [([data[v].rownames[r] for r in row_s[v]], [data[v].colnames[c] for c in col_s[v]])
                     for v in range(n_views)]
            if on_dev:
                new_data = [ResidentMatrix((len(row_s[v]), len(col_s[v])), *names[v]) for v in range(n_views)]
                pool.place_gather(key, "data", row_s, col_s)
            else:
                new_data = [_subset(data[v], row_s[v], col_s[v]) for v in range(n_views)]
                pool.place_host(key, new_data)
            reordered = reorder_data(new_data, n_views, [m.rownames for m in new_data], [m.colnames for m in new_data])
            specs.append(dict(key=key, data=new_data, row_indices=reordered["row_indices"],
                              col_indices=reordered["col_indices"], k_vec=[k] * n_views, rng=child_rngs[i],
                              want_bisil=False))  # the resample fits are read for their clusters only (:268-276)
        new_results = run_fits(pool, specs, phi, xi, psi, n_iters, num_repeats, spurious, distance, False,
                               max_iters=max_iters)
        relevance = np.zeros((n_views, k))
        for (row_s, col_s), new_res in zip(samples, new_results):  # in repeat order, as the reference accumulates them
            for i in range(n_views):
                relevance[i, :] += relevance_results(new_res["row_clusters"][i], new_res["col_clusters"][i],
                                                     results["row_clusters"][i][row_s[i], :],
                                                     results["col_clusters"][i][col_s[i], :])
    finally:
        if own_pool:
            pool.close()
    relevance = relevance / n_stability
    if not remove_unstable:
        return {"res": results, "relevance": relevance}
    for i in range(n_views):
        unstable = relevance[i, :] < stab_thres
        results["row_clusters"][i][:, unstable] = 0.0
        results["col_clusters"][i][:, unstable] = 0.0
    return results
