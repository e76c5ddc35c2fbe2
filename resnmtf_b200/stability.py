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
    """R/stability_analysis.r:94-99."""
    return float(sum(x.sum() for x in results["row_clusters"]))


def _sample(rng, n, size):
    """R's sample(n, size): `size` is truncated towards zero."""
    return rng.permutation(int(n))[: int(size)]


def _check_empty(m):
    return (m.x.sum(0) == 0).any() or (m.x.sum(1) == 0).any()


def _subset(view, rows, cols):
    return NamedMatrix(np.asfortranarray(view.x[np.ix_(rows, cols)]),
                       [view.rownames[i] for i in rows], [view.colnames[j] for j in cols])


def stability_repeat(results, data, dim_1, k, phi, xi, psi, n_iters, num_repeats, distance, spurious,
                     n_views, sample_rate, rng, ctx):
    """R/stability_analysis.r:215-278."""
    from .api import res_nmtf_inner

    new_data = [None] * n_views
    row_samples = [None] * n_views
    col_samples = [None] * n_views
    relevance = np.zeros((n_views, k))
    attempt = 1
    while test_cond(new_data, attempt):
        if attempt == 20:
            print("Unable to perform stability analysis due to sparsity of data.")
            return {"stability_performed": False}
        row_samples[0] = _sample(rng, dim_1[0], dim_1[0] * sample_rate)
        col_samples[0] = _sample(rng, dim_1[1], dim_1[1] * sample_rate)
        new_data[0] = _subset(data[0], row_samples[0], col_samples[0])
        if _check_empty(new_data[0]):
            keep_c = new_data[0].x.sum(0) != 0
            keep_r = new_data[0].x.sum(1) != 0
            row_samples[0] = row_samples[0][keep_r]
            col_samples[0] = col_samples[0][keep_c]
            new_data[0] = _subset(data[0], row_samples[0], col_samples[0])
        for i in range(1, n_views):
            dims = data[i].shape
            # initial_shuffle (R/stability_analysis.r:111-132): reuse view 1's sample when the dims agree
            row_samples[i] = row_samples[0] if dims[0] == dim_1[0] else _sample(rng, dims[0], dims[0] * sample_rate)
            col_samples[i] = col_samples[0] if dims[1] == dim_1[1] else _sample(rng, dims[1], dims[1] * sample_rate)
            new_data[i] = _subset(data[i], row_samples[i], col_samples[i])
            if _check_empty(new_data[i]):  # sample_view (R/stability_analysis.r:157-192)
                keep_c = new_data[i].x.sum(0) != 0
                keep_r = new_data[i].x.sum(1) != 0
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
                for q in range(i + 1):
                    new_data[q] = _subset(data[q], row_samples[q], col_samples[q])
        attempt += 1
    reordered = reorder_data(new_data, n_views, [m.rownames for m in new_data], [m.colnames for m in new_data])
    new_results = res_nmtf_inner(new_data, reordered["row_indices"], reordered["col_indices"],
                                 k_vec=[k] * n_views, phi=phi, xi=xi, psi=psi, n_iters=n_iters,
                                 num_repeats=num_repeats, spurious=spurious, distance=distance, rng=rng, ctx=ctx)
    for i in range(n_views):
        relevance[i, :] += relevance_results(new_results["row_clusters"][i], new_results["col_clusters"][i],
                                             results["row_clusters"][i][row_samples[i], :],
                                             results["col_clusters"][i][col_samples[i], :])
    return {"relevance": relevance, "stability_performed": True}


def stability_check(data, results, k, phi, xi, psi, n_iters, spurious, num_repeats, no_clusts, distance,
                    sample_rate=0.9, n_stability=5, stab_thres=0.6, remove_unstable=True, rng=None, ctx=None,
                    use_parallel=True):
    """R/stability_analysis.r:302-338.  ``k`` may be a vector (the fixed-k caller passes k_vec; base R's
    matrix(ncol = k) then uses its first element -- quirk Q9).  The ``n_stability`` repeats are independent fits
    (SURVEY 8e): each gets its own child generator -- so the result does not depend on how many GPUs run them --
    and, with ``use_parallel`` and more than one visible GPU, they are dealt round-robin to one context (and one
    host thread) per GPU, like the fits of the k-sweep."""
    if number_biclusters(results) == 0:
        print("No biclusters detected!")
        return results
    from .device import device_contexts

    k = int(np.atleast_1d(k)[0])
    n_views = len(data)
    dim_1 = data[0].shape
    n_rep = int(n_stability)
    rng = np.random.default_rng() if rng is None else rng
    child_rngs = rng.spawn(n_rep)
    contexts = device_contexts(ctx) if (use_parallel and ctx is not None) else [ctx]
    reps = [None] * n_rep

    def run_rank(r):
        for i in range(r, n_rep, len(contexts)):
            reps[i] = stability_repeat(results, data, dim_1, k, phi, xi, psi, n_iters, num_repeats, distance, spurious,
                                       n_views, sample_rate, child_rngs[i], contexts[r])

    if len(contexts) == 1 or n_rep <= 1:
        contexts = contexts[:1]
        run_rank(0)
    else:
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=len(contexts)) as pool:  # ctypes releases the GIL inside the library
            list(pool.map(run_rank, range(len(contexts))))
    relevance = np.zeros((n_views, k))
    for rep in reps:  # in repeat order, as the reference accumulates them
        if not rep["stability_performed"]:
            return results
        relevance = relevance + rep["relevance"]
    relevance = relevance / n_stability
    if not remove_unstable:
        return {"res": results, "relevance": relevance}
    for i in range(n_views):
        unstable = relevance[i, :] < stab_thres
        results["row_clusters"][i][:, unstable] = 0.0
        results["col_clusters"][i][:, unstable] = 0.0
    return results
