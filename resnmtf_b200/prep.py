"""Host-side input preparation, mirroring the reference's R helpers one for one (same names, argument
meaning, error and warning strings) so the drop-in keeps the reference's behaviour around the device loop:

  init_rest_mats        R/update_steps.r:12-24
  make_non_neg(_inner)  R/utils.r:9-27
  matrix_normalisation  R/utils.r:86-88
  check_*               R/utils.r:220-454
  give_names            R/utils.r:469-542
  reorder_data          R/utils.r:619-662 (+ produce_indices :560-601)

The one structural change: the reference stores shared row/column *names* in ``hash`` objects and matches
character vectors on every update; here ``shared_maps`` turns them into int32 index pairs once, which is
what the C ABI (resnmtf_fit_set_shared_map) consumes.
"""
from __future__ import annotations

import warnings

import numpy as np


class NamedMatrix:
    """A float64 matrix with optional row / column names (R's dimnames).  ``rownames`` / ``colnames`` are
    lists of ``str`` (an entry may be ``None`` = R's NA) or ``None`` (R's NULL)."""

    __slots__ = ("x", "rownames", "colnames")

    def __init__(self, x, rownames=None, colnames=None):
        self.x = np.asarray(x)
        if self.x.ndim != 2:
            raise ValueError("a view must be a 2-d matrix")
        self.rownames = None if rownames is None else list(rownames)
        self.colnames = None if colnames is None else list(colnames)
        if self.rownames is not None and len(self.rownames) != self.x.shape[0]:
            raise ValueError("length of rownames does not match the number of rows")
        if self.colnames is not None and len(self.colnames) != self.x.shape[1]:
            raise ValueError("length of colnames does not match the number of columns")

    @property
    def shape(self):
        return self.x.shape

    def copy(self):
        return NamedMatrix(self.x, self.rownames, self.colnames)


def as_named(m):
    """numpy array / pandas DataFrame / NamedMatrix -> NamedMatrix (names kept where they exist)."""
    if isinstance(m, NamedMatrix):
        return m.copy()
    if hasattr(m, "index") and hasattr(m, "columns") and hasattr(m, "to_numpy"):  # pandas.DataFrame
        return NamedMatrix(m.to_numpy(), [str(s) for s in m.index], [str(s) for s in m.columns])
    return NamedMatrix(np.asarray(m))


# --------------------------------------------------------------------------------------------------
# restriction matrices
# --------------------------------------------------------------------------------------------------


def init_rest_mats(mat, n_v):
    """R/update_steps.r:12-24: NULL -> zeros; else zero the diagonal and return mat + t(mat) (a symmetric
    input is therefore doubled, exactly like the reference)."""
    if mat is None:
        return np.zeros((n_v, n_v))
    m = np.array(mat, dtype=np.float64, copy=True)
    np.fill_diagonal(m, 0.0)
    return m + m.T


# --------------------------------------------------------------------------------------------------
# non-negativity and normalisation
# --------------------------------------------------------------------------------------------------


def make_non_neg_inner(matrix):
    """R/utils.r:20-27: per-column shift by |min(0, min(col))|; warns when any element is negative."""
    x = np.asarray(matrix, dtype=np.float64)
    non_neg = x + np.abs(np.minimum(0.0, x.min(axis=0)))[None, :]
    if (x < 0).any():
        warnings.warn("Matrix is not non-negative. Has been made non-negative.")
    return non_neg


def make_non_neg(x):
    return [make_non_neg_inner(m) for m in x]


def matrix_normalisation(matrix):
    """R/utils.r:86-88: L1 column normalisation."""
    matrix = np.asarray(matrix, dtype=np.float64)
    return matrix / matrix.sum(axis=0)[None, :]


# --------------------------------------------------------------------------------------------------
# argument checks (error strings as in the reference)
# --------------------------------------------------------------------------------------------------


def _is_numeric(x):
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)


def check_whole_number(x, name):
    if not _is_numeric(x):
        raise ValueError(f"{name}  must be a numeric.")
    if np.floor(x) != x or x <= 0:
        raise ValueError(f"{name}  must be a positive integer.")


def check_integers(n_iters, k_min, k_max, num_repeats, n_stability):
    if n_iters is not None:
        check_whole_number(n_iters, "n_iters")
    check_whole_number(num_repeats, "num_repeats")
    check_whole_number(n_stability, "n_stability")
    check_whole_number(k_min, "k_min")
    check_whole_number(k_max, "k_max")
    if k_max <= k_min:
        raise ValueError("k_max must be greater than k_min.")


def check_boolean(no_clusts, stability, remove_unstable, spurious):
    for val, name in ((no_clusts, "no_clusts"), (stability, "stability"),
                      (remove_unstable, "remove_unstable"), (spurious, "spurious")):
        if not isinstance(val, (bool, np.bool_)):
            raise ValueError(f"{name} must be a boolean.")


def check_numeric(sample_rate, stab_thres):
    if not _is_numeric(sample_rate):
        raise ValueError("sample_rate must be a numeric.")
    if not _is_numeric(stab_thres):
        raise ValueError("stab_thres must be a numeric.")
    if stab_thres < 0 or stab_thres > 1:
        raise ValueError("stab_thres must be between 0 and 1.")
    if sample_rate <= 0 or sample_rate > 1:
        raise ValueError("sample_rate must be greater than 0 and less than or equal to 1.")


def check_lists(data, init_f, init_s, init_g):
    """R/utils.r:311-337.  Like the reference, a *list* init is rejected here (the reference's test is
    inverted), so explicit inits can only be given to res_nmtf_inner directly."""
    if not isinstance(data, (list, tuple)):
        if isinstance(data, (np.ndarray, NamedMatrix)) or hasattr(data, "to_numpy"):
            data = [data]
        else:
            raise ValueError("Data must be a list of matrices or a matrix.")
    for val, name in ((init_f, "init_f"), (init_s, "init_s"), (init_g, "init_g")):
        if val is not None and isinstance(val, (list, tuple)):
            raise ValueError(f"{name} must be a list of matrices or NULL.")
    return list(data)


def check_restriction_mat(data, matrix, name):
    if matrix is not None:
        if not isinstance(matrix, np.ndarray) or matrix.ndim != 2:
            raise ValueError(f"{name} must be a matrix or NULL.")
        if (matrix < 0).any():
            raise ValueError(f"{name}  must be a non-negative matrix.")
        if matrix.shape != (len(data), len(data)):
            raise ValueError(f"{name}  must be of the same dimensions as data.")


def check_inputs(data, init_f, init_s, init_g, k_vec, phi, xi, psi, n_iters, k_min, k_max, distance,
                 num_repeats, no_clusts, sample_rate, n_stability, stability, stab_thres, remove_unstable,
                 spurious, prep_values=True):
    """R/utils.r:391-454.  ``data`` is a list of NamedMatrix; returns the prepped list (non-negative, L1
    column-normalised float64, names kept).  ``prep_values=False`` leaves the two matrix-sized passes (the shift to
    non-negative values and the normalisation) to the caller, who runs them on the GPU on the way in
    (``FitPool.place_resident``); everything else -- checks, error strings -- is unchanged."""
    data = check_lists(data, init_f, init_s, init_g)
    check_integers(n_iters, k_min, k_max, num_repeats, n_stability)
    check_boolean(no_clusts, stability, remove_unstable, spurious)
    check_numeric(sample_rate, stab_thres)
    named = [as_named(m) for m in data]
    if named[0].x.dtype != np.float64:
        warnings.warn("Data is not a double matrix. Converting to double.")
    out = []
    for m in named:
        if prep_values:
            x = np.asfortranarray(matrix_normalisation(make_non_neg_inner(m.x)))
        else:
            x = np.asarray(m.x, dtype=np.float64)
        out.append(NamedMatrix(x, m.rownames, m.colnames))
    ranks = [m.shape[1] for m in out]
    if distance not in ("euclidean", "manhattan", "cosine"):
        raise ValueError("distance must be one of 'euclidean', 'manhattan' or 'cosine'.")
    check_restriction_mat(out, phi, "phi")
    check_restriction_mat(out, xi, "xi")
    check_restriction_mat(out, psi, "psi")
    if k_vec is not None:
        kv = np.asarray(k_vec)
        if not np.issubdtype(kv.dtype, np.number):
            raise ValueError("k_vec must be a vector of integers.")
        if (kv < 1).any():
            raise ValueError("k_vec must be a vector of integers greater than 1.")
        if kv.size != len(out):
            raise ValueError("k_vec must be a vector of the same length as the number of views.")
        if (kv > np.asarray(ranks)).any():
            raise ValueError("k_vec must be a vector of integers less than or equal to the\n            ranks of the views.")
    else:
        if k_max > min(ranks):
            raise ValueError("k_max must be less than or equal to the minimum rank of the views.")
    return out


# --------------------------------------------------------------------------------------------------
# naming and shared index sets
# --------------------------------------------------------------------------------------------------


def give_names(data, n_views, phi=None, psi=None):
    """R/utils.r:469-542.  Returns dict(data=[NamedMatrix], row_names=[...], col_names=[...])."""
    data = [as_named(m) for m in data]
    missing_rows = [m.rownames is None for m in data]
    missing_cols = [m.colnames is None for m in data]
    if all(missing_rows):
        n = 1
        for i in range(n_views):
            if data[i].rownames is None:
                data[i].rownames = [f"row_{j}" for j in range(n, n + data[i].shape[0])]
                n += data[i].shape[0]
            if phi is not None:
                for j in range(min(i + 1, n_views - 1), n_views):
                    if phi[i, j] > 0 and data[i].shape[0] != data[j].shape[0]:
                        raise ValueError("Row restriction matrices implies shared rows between views\n"
                                         "               with differing number of unnamed rows. Please name rows.")
                    elif phi[i, j] > 0:
                        data[j].rownames = list(data[i].rownames)
    elif any(missing_rows):
        raise ValueError("At least one view is missing row names. Please name missing rows.")
    else:
        if any(any(s is None for s in m.rownames) for m in data):
            raise ValueError("Some rows missing names. Check row names.")
    if all(missing_cols):
        n = 1
        for i in range(n_views):
            if data[i].colnames is None:
                data[i].colnames = [f"col_{j}" for j in range(n, n + data[i].shape[1])]
                n += data[i].shape[1]
            if psi is not None:
                for j in range(min(i + 1, n_views - 1), n_views):
                    if psi[i, j] > 0 and data[i].shape[1] != data[j].shape[1]:
                        raise ValueError("Column restriction matrices implies shared columns between\n"
                                         "            views with differing number of unnamed columns.\n"
                                         "            Please name columns.")
                    elif psi[i, j] > 0:
                        data[j].colnames = list(data[i].colnames)
    elif any(missing_cols):
        raise ValueError("At least one view is missing column names. Please name missing columns.")
    else:
        if any(any(s is None for s in m.colnames) for m in data):
            raise ValueError("Some columns missing names. Check columns names.")
    return {"data": data, "row_names": [m.rownames for m in data], "col_names": [m.colnames for m in data]}


def reorder_data(data, n_views, row_names, col_names):
    """R/utils.r:619-662 + produce_indices (:560-601): for every view v a dict {w: shared names or None}.

    The reference enumerates the power set of views and collects, per view subset A, the names that occur
    in exactly the views of A; the names shared by v and w are then the union over all A containing both.
    That union is simply names(v) & names(w) (kept here in view v's order), which is what is computed --
    without the 2^n_views enumeration.  ``None`` stands for R's NA (nothing shared)."""
    row_sets = [set(r) for r in row_names]
    col_sets = [set(c) for c in col_names]
    row_indices, col_indices = [], []
    for v in range(n_views):
        rd, cd = {}, {}
        for w in range(n_views):
            if w == v:
                continue
            cr = [s for s in row_names[v] if s in row_sets[w]]
            cc = [s for s in col_names[v] if s in col_sets[w]]
            rd[w] = cr if cr else None
            cd[w] = cc if cc else None
        row_indices.append(rd)
        col_indices.append(cd)
    return {"row_indices": row_indices, "col_indices": col_indices}


def shared_maps(indices, names):
    """Turns ``indices`` (list over v of {w: names or None}, or None for R's NULL) into what the C ABI
    takes: {(v, w): (idx_v, idx_w)} with int32 arrays; an NA pair maps to two empty arrays; pairs absent
    from the result are 'never set' (R's NULL quirk, see resnmtf_fit_set_shared_map)."""
    out = {}
    if indices is None:
        return out
    pos = [{s: i for i, s in enumerate(nm)} for nm in names]
    for v, d in enumerate(indices):
        if d is None:
            continue
        for w, shared in d.items():
            if shared is None:
                out[(v, w)] = (np.empty(0, np.int32), np.empty(0, np.int32))
            else:
                iv = np.fromiter((pos[v][s] for s in shared), dtype=np.int32, count=len(shared))
                iw = np.fromiter((pos[w][s] for s in shared), dtype=np.int32, count=len(shared))
                out[(v, w)] = (iv, iw)
    return out
