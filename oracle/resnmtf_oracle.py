"""CPU oracle for the ResNMTF multiplicative-update loop.  TEST INFRASTRUCTURE ONLY.

This file restates, in NumPy FP64, the algorithm of the reference R package
eso28599/resnmtf for the hot path named by BASELINE.json (the F/S/G updates
driven by ``res_nmtf_inner``).  It is the *checker* for the CUDA path, never the
thing shipped: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package ``resnmtf_b200`` must never import anything from ``oracle/``.

Pinning status
--------------
The reference is pure R; R is not installed in this image, and the reference's
tests hold no golden numbers (no seeds anywhere upstream).  The oracle is pinned
against what the reference's tests *do* hold for this path -- the property tests
of tests/testthat/test-resnmtf.R:63-184 (cluster sizes {60,60,60}, colSums(F) =
colSums(G) = 1, reconstruction column sums, phi/psi pulling shared rows/cols
together) -- see tests/test_oracle_reference_properties.py.  Numeric parity
against an actual R run is **unpinned** (no R here to produce vectors).

Every function cites the reference file:line it follows (paths relative to the
reference root).  Operation order inside each expression follows R's
left-to-right evaluation of ``%*%`` so that the rounding is as close to the
reference as a different BLAS allows.

Conventions: a "view" is a float64 matrix (any memory order; results do not depend
on it), row/col names are lists of ``str``.  ``row_indices[v]`` is a dict that
maps ``w`` (0-based view index) to the list of row names shared by views v and w,
or ``None`` where R stores ``NA`` (produce_indices, R/utils.r:560-601).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# restriction matrices, naming, shared index sets
# --------------------------------------------------------------------------------------


def init_rest_mats(mat, n_v):
    """R/update_steps.r:12-24 -- NULL -> zeros; else zero the diagonal and return mat + t(mat)."""
    if mat is None:
        return np.zeros((n_v, n_v))
    m = np.array(mat, dtype=np.float64, copy=True)
    np.fill_diagonal(m, 0.0)
    return m + m.T


def default_names(data):
    """Auto names of give_names (R/utils.r:469-542) for fully unnamed data and no restrictions:
    row_1.. continuing across views, col_1.. likewise."""
    rn, cn = [], []
    r = c = 1
    for x in data:
        rn.append([f"row_{i}" for i in range(r, r + x.shape[0])])
        cn.append([f"col_{i}" for i in range(c, c + x.shape[1])])
        r += x.shape[0]
        c += x.shape[1]
    return rn, cn


def shared_names(names):
    """reorder_data + produce_indices (R/utils.r:560-662) reduced to what the update rules consume:
    for every ordered pair (v, w), w != v, the names present in both views (None when there are none,
    R's NA).  The reference builds these through the power set of views; the union over all view
    subsets containing {v, w} of "names in exactly that subset" is exactly names(v) & names(w)."""
    n_v = len(names)
    out = []
    for v in range(n_v):
        d = {}
        for w in range(n_v):
            if w == v:
                continue
            sw = set(names[w])
            common = [s for s in names[v] if s in sw]
            d[w] = common if len(common) else None
        out.append(d)
    return out


# --------------------------------------------------------------------------------------
# coupling sums
# --------------------------------------------------------------------------------------


def star_prod(vec, mat_list):
    """R/utils.r:39-47 -- sum_i vec[i] * mat_list[i] over non-zero vec[i]."""
    acc = 0.0
    for i in range(len(vec)):
        if vec[i] != 0:
            acc = acc + vec[i] * mat_list[i]
    return acc


def star_prod_relevant(vec, mat_list, current_mat, indices, own_names, all_names):
    """R/utils.r:63-78.

    For every view i with vec[i] != 0 whose shared-name set with the current view is not NA:
    take a copy of ``current_mat``, overwrite the rows *named* in the shared set by the same-named
    rows of ``mat_list[i]``, and accumulate vec[i] * masked * nrow(mat_list[i]).  Finally divide by
    nrow(current_mat).  ``own_names`` / ``all_names[i]`` give the row names used for the lookup."""
    acc = 0.0
    pos_own = {s: j for j, s in enumerate(own_names)}
    for i in range(len(vec)):
        if vec[i] != 0:
            rows = indices.get(i) if indices is not None else None
            if i in (indices or {}) and rows is None:
                continue  # NA: pair shares nothing -> skipped (utils.r:70)
            masked = np.array(current_mat, dtype=np.float64, copy=True)
            if rows is not None:
                pos_other = {s: j for j, s in enumerate(all_names[i])}
                dst = np.fromiter((pos_own[s] for s in rows), dtype=np.int64, count=len(rows))
                src = np.fromiter((pos_other[s] for s in rows), dtype=np.int64, count=len(rows))
                masked[dst, :] = mat_list[i][src, :]
            # rows is None and i not in indices: R's `indices = NULL` case (quirk Q3) -> nothing overwritten
            acc = acc + vec[i] * masked * mat_list[i].shape[0]
    return acc / current_mat.shape[0]


# --------------------------------------------------------------------------------------
# update rules
# --------------------------------------------------------------------------------------


def update_f(x, input_f, input_s, input_g, lambda_in, phi, v, row_indices, row_names):
    """R/update_steps.r:141-165."""
    current_f = input_f[v]
    numerator = (x @ input_g) @ input_s.T
    denominator = (current_f @ input_s) @ ((input_g.T @ input_g) @ input_s.T)
    phi_vec = phi[:, v]
    lambda_mat = 0.5 * np.broadcast_to(lambda_in, current_f.shape)
    with np.errstate(divide="ignore", invalid="ignore"):
        if phi_vec.sum() == 0:
            ratio = numerator / (denominator + lambda_mat)
            ratio[np.isnan(ratio)] = 1.0
            out = current_f * ratio
        else:
            num_prod = star_prod_relevant(phi_vec, input_f, current_f, row_indices, row_names[v], row_names)
            den_prod = phi_vec.sum() * current_f
            out = current_f * ((numerator + num_prod) / (denominator + den_prod + lambda_mat))
    return np.abs(out)


def update_g(x, input_f, input_s, input_g, mu_in, psi, v, col_indices, col_names):
    """R/update_steps.r:180-207 (note the whole-matrix sum(psi) test, :190)."""
    current_g = input_g[v]
    numerator = (x.T @ input_f) @ input_s
    denominator = (current_g @ input_s.T) @ ((input_f.T @ input_f) @ input_s)
    mu_mat = 0.5 * np.broadcast_to(mu_in, current_g.shape)
    with np.errstate(divide="ignore", invalid="ignore"):
        if psi.sum() == 0:
            ratio = numerator / (denominator + mu_mat)
            ratio[np.isnan(ratio)] = 1.0
            out = current_g * ratio
        else:
            psi_vec = psi[:, v]
            num_prod = star_prod_relevant(psi_vec, input_g, current_g, col_indices, col_names[v], col_names)
            den_prod = psi_vec.sum() * current_g
            out = current_g * ((numerator + num_prod) / (denominator + den_prod + mu_mat))
    return np.abs(out)


def update_s(x, input_f, input_s, input_g, xi, v):
    """R/update_steps.r:220-240 (whole-matrix sum(xi) test, :226)."""
    current_s = input_s[v]
    numerator = (input_f.T @ x) @ input_g
    denominator = ((input_f.T @ input_f) @ current_s) @ (input_g.T @ input_g)
    with np.errstate(divide="ignore", invalid="ignore"):
        if xi.sum() == 0:
            ratio = numerator / denominator
            ratio[np.isnan(ratio)] = 1.0
            out = current_s * ratio
        else:
            xi_vec = xi[:, v]
            num_prod = star_prod(xi_vec, input_s)
            den_prod = xi_vec.sum() * current_s
            out = current_s * ((numerator + num_prod) / (denominator + den_prod))
    return np.abs(out)


def update_lm(vec, matrix):
    """R/update_steps.r:249-251."""
    return matrix.sum(axis=0) * vec


def update_matrices(x, input_f, input_s, input_g, lam, mu, phi, xi, psi,
                    row_indices, col_indices, row_names, col_names):
    """R/update_steps.r:272-319 -- Gauss-Seidel sweep over the views, lists updated in place."""
    n_v = len(x)
    cf, cs, cg = list(input_f), list(input_s), list(input_g)
    cl, cm = list(lam), list(mu)
    for v in range(n_v):
        cf[v] = update_f(x[v], cf, cs[v], cg[v], cl[v], phi, v,
                         row_indices[v] if row_indices is not None else None, row_names)
        cg[v] = update_g(x[v], cf[v], cs[v], cg, cm[v], psi, v,
                         col_indices[v] if col_indices is not None else None, col_names)
        cs[v] = update_s(x[v], cf[v], cs, cg[v], xi, v)
        cl[v] = update_lm(cl[v], cf[v])
        cm[v] = update_lm(cm[v], cg[v])
    return cf, cs, cg, cl, cm


def calculate_error(data, cf, cs, cg, data_norms):
    """R/utils.r:157-166 -- materialises X_hat like the reference does."""
    err = np.zeros(len(data))
    for v in range(len(data)):
        x_hat = (cf[v] @ cs[v]) @ cg[v].T
        err[v] = np.linalg.norm(data[v] - x_hat, "fro") ** 2
    return err / data_norms


def normalisation_check(cf, cg, cs):
    """R/utils.r:176-195 -- S columns scaled by cs_F*cs_G (quirk Q7), then F, G to unit column sums."""
    cf, cg, cs = list(cf), list(cg), list(cs)
    for v in range(len(cf)):
        csf = cf[v].sum(axis=0)
        csg = cg[v].sum(axis=0)
        cs[v] = cs[v] * (csf * csg)[None, :]
        cf[v] = cf[v] / csf[None, :]
        cg[v] = cg[v] / csg[None, :]
    return cf, cg, cs


# --------------------------------------------------------------------------------------
# initialisation
# --------------------------------------------------------------------------------------


def init_mats_inner(x, k_vec, noise):
    """R/update_steps.r:78-125.  ``noise[v]`` stands in for abs(MASS::mvrnorm(k, 0, 0.05 I))[1:k,1:k]
    (the R RNG stream cannot be reproduced here; callers pass |sqrt(0.05) * N(0,1)| draws)."""
    fs, ss, gs, lams, mus = [], [], [], [], []
    for i, xi_ in enumerate(x):
        k = int(k_vec[i])
        u, d, vt = np.linalg.svd(xi_, full_matrices=False)
        f = np.abs(u[:, :k])
        g = np.abs(vt[:k, :].T)
        s = np.abs(np.diag(d[:k])) + np.abs(noise[i])
        csf = f.sum(axis=0)
        csg = g.sum(axis=0)
        s = s * (csf * csg)[None, :]
        f = f / csf[None, :]
        g = g / csg[None, :]
        fs.append(f)
        ss.append(s)
        gs.append(g)
        lams.append(f.sum(axis=0))
        mus.append(g.sum(axis=0))
    return fs, ss, gs, lams, mus


def init_mats(x, k_vec, init_f, init_g, init_s, noise=None):
    """R/update_steps.r:36-66."""
    if init_f is None or init_g is None or init_s is None:
        return init_mats_inner(x, k_vec, noise)
    cf = [np.array(a, dtype=np.float64) for a in init_f]
    cs = [np.array(a, dtype=np.float64) for a in init_s]
    cg = [np.array(a, dtype=np.float64) for a in init_g]
    return cf, cs, cg, [a.sum(axis=0) for a in cf], [a.sum(axis=0) for a in cg]


# --------------------------------------------------------------------------------------
# the loop
# --------------------------------------------------------------------------------------


def res_nmtf_loop(data, row_indices, col_indices, row_names, col_names,
                  init_f, init_s, init_g, k_vec, phi, xi, psi, n_iters=None,
                  noise=None, tol=1.0e-6, max_iters=None, trace=None):
    """R/main.r:38-114: init, data norms, convergence or fixed loop, final normalisation.

    Returns a dict with the normalised factors, lambda/mu, ``All_Error`` and ``Error`` exactly as
    R/main.r:126-139 defines them, plus the un-normalised factors (for per-iteration parity).
    ``trace(t, cf, cs, cg, cl, cm, err_vec)`` is called after every sweep when given.
    ``max_iters`` is a safety net the reference does not have (R/main.r:55 has no cap)."""
    n_v = len(data)
    data = [np.asarray(x, dtype=np.float64) for x in data]
    phi = np.asarray(phi, dtype=np.float64)
    xi = np.asarray(xi, dtype=np.float64)
    psi = np.asarray(psi, dtype=np.float64)
    cf, cs, cg, cl, cm = init_mats(data, k_vec, init_f, init_g, init_s, noise)
    data_norms = np.array([np.linalg.norm(x, "fro") ** 2 for x in data])  # main.r:48
    total_err = []

    def sweep(t):
        nonlocal cf, cs, cg, cl, cm
        cf, cs, cg, cl, cm = update_matrices(data, cf, cs, cg, cl, cm, phi, xi, psi,
                                             row_indices, col_indices, row_names, col_names)
        err = calculate_error(data, cf, cs, cg, data_norms)
        if trace is not None:
            trace(t, cf, cs, cg, cl, cm, err)
        return float(np.mean(err))

    if n_iters is None:
        err_diff, err_temp, t = 1.0, 0.0, 0
        while err_diff > tol:  # main.r:55 (NaN here makes R throw; we raise likewise)
            mean_err = sweep(t)
            total_err.append(mean_err)
            if np.isnan(mean_err):
                raise FloatingPointError("missing value where TRUE/FALSE needed")
            err_diff = abs(mean_err - err_temp)
            err_temp = mean_err
            t += 1
            if max_iters is not None and t >= max_iters:
                break
        error = float(np.mean(total_err[-10:]))  # main.r:126-127
    else:
        for t in range(int(n_iters)):
            total_err.append(sweep(t))
        error = total_err[-1]  # main.r:129
    raw = (list(cf), list(cs), list(cg))
    nf, ng, ns = normalisation_check(cf, cg, cs)
    return {
        "output_f": nf, "output_s": ns, "output_g": ng,
        "Error": error, "All_Error": np.array(total_err),
        "lambda": cl, "mu": cm,
        "raw_f": raw[0], "raw_s": raw[1], "raw_g": raw[2],
    }


# --------------------------------------------------------------------------------------
# binarisation (obtain_biclusters without spurious removal and without the bisilhouette call)
# --------------------------------------------------------------------------------------


def binarise(output_f, output_g, output_s):
    """R/obtain_bicl.r:162-180: row_cl = 1[F > 1/n], col_cl = 1[G > 1/p];
    relations[j] = which.max(S[, j]) (first maximum); row_cl <- row_cl[, relations]."""
    rows, cols, rels = [], [], []
    for f, g, s in zip(output_f, output_g, output_s):
        rc = (f > (1.0 / f.shape[0])).astype(np.float64)
        cc = (g > (1.0 / g.shape[0])).astype(np.float64)
        rel = np.argmax(s, axis=0)  # first max, like which.max
        rows.append(rc[:, rel])
        cols.append(cc)
        rels.append(rel)
    return rows, cols, rels


# --------------------------------------------------------------------------------------
# host prep the reference applies before the loop (needed to feed the loop the same bytes)
# --------------------------------------------------------------------------------------


def make_non_neg(x):
    """R/utils.r:20-27 -- per-column shift by |min(0, min(col))|; returns (matrix, warned)."""
    x = np.asarray(x, dtype=np.float64)
    shift = np.abs(np.minimum(0.0, x.min(axis=0)))
    return x + shift[None, :], bool((x < 0).any())


def matrix_normalisation(x):
    """R/utils.r:86-88 -- L1 column normalisation."""
    return x / x.sum(axis=0)[None, :]
