"""Thin object layer over the C ABI: a ``Context`` (one GPU, one stream) and a ``DeviceFit`` that owns the
device-resident X, F, S, G, lambda, mu of all views of one ``res_nmtf_inner`` call."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L


def _f64(a):
    """Column-major float64 copy/view of ``a`` (what R would hand over)."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    def __init__(self, device=-1):
        lib = L.require_device()
        self._lib = lib
        h = C.c_void_p()
        L.check(lib.resnmtf_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(lib.resnmtf_ctx_device(h))

    @property
    def stream(self):
        """cudaStream_t (int) the context launches on."""
        return int(self._lib.resnmtf_ctx_stream(self._h) or 0)

    def synchronize(self):
        L.check(self._lib.resnmtf_ctx_synchronize(self._h))

    # ---- row-sharded views over ranks (one process per GPU) ---------------------------------------
    @staticmethod
    def comm_id_create():
        """Opaque NCCL unique id (bytes) created on one rank; broadcast it to the others by any means."""
        lib = L.require_device()
        buf = C.create_string_buffer(lib.resnmtf_comm_id_size())
        L.check(lib.resnmtf_comm_id_create(buf))
        return buf.raw

    def join(self, comm_id, rank, n_ranks):
        """After this, every fit created on the context is row-sharded: n[v] is the local row count, F rows
        are local, G / S / lambda / mu are replicated and [X'F | F'F | colSums(F)] is all-reduced per sweep."""
        buf = C.create_string_buffer(bytes(comm_id), len(comm_id))
        L.check(self._lib.resnmtf_ctx_join(self._h, buf, int(rank), int(n_ranks)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.resnmtf_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


_device_ctx = {}


def device_contexts(first):
    """One context per visible GPU, ``first`` (and its device) leading: the placement targets of the
    independent fits of an apply_resnmtf call."""
    n = L.device_count()
    cap = int(os.environ.get("RESNMTF_MAX_GPUS", "0") or 0)  # 0: all visible GPUs
    if cap > 0:
        n = min(n, max(cap, first.device + 1))
    out = [first]
    for dev in range(n):
        if dev == first.device:
            continue
        if dev not in _device_ctx:
            _device_ctx[dev] = Context(dev)
        out.append(_device_ctx[dev])
    return out[:cap] if cap > 0 else out


def jsd_pairs(ctx, vecs, bw, vmax, pair_a, pair_b):
    """jsd_calc (R/utils.r:95-106) of the column pairs (pair_a[i], pair_b[i]) of the n x m matrix ``vecs`` on the
    GPU (``resnmtf_jsd_pairs``); ``bw`` / ``vmax``: bw.nrd0 and maximum of every column."""
    lib = L.require_device()
    vecs = _f64(vecs)
    bw = np.ascontiguousarray(bw, dtype=np.float64)
    vmax = np.ascontiguousarray(vmax, dtype=np.float64)
    pa = np.ascontiguousarray(pair_a, dtype=np.int32)
    pb = np.ascontiguousarray(pair_b, dtype=np.int32)
    if pa.shape != pb.shape or bw.size != vecs.shape[1] or vmax.size != vecs.shape[1]:
        raise ValueError("jsd_pairs: inconsistent arguments")
    out = np.empty(pa.size, dtype=np.float64)
    lock = getattr(ctx, "lock", None)  # a context shared by several host threads (native_route) carries one
    if lock is not None:
        lock.acquire()
    try:
        L.check(lib.resnmtf_jsd_pairs(ctx._h, _ptr(vecs), vecs.shape[0], vecs.shape[1], vecs.shape[0], _ptr(bw),
                                      _ptr(vmax), _ptr(pa), _ptr(pb), pa.size, _ptr(out)))
    finally:
        if lock is not None:
            lock.release()
    return out


class DeviceData:
    """One view uploaded once into the device layout and shared by several fits (the k-sweep of
    apply_resnmtf fits the same data for every k).  Reference counted on the C side."""

    def __init__(self, ctx, x):
        self._lib = L.require_device()
        self.ctx = ctx
        x = _f64(x)
        self.shape = x.shape
        h = C.c_void_p()
        L.check(self._lib.resnmtf_data_create(ctx._h, x.shape[0], x.shape[1], _ptr(x), x.shape[0], C.byref(h)))
        self._h = h

    @classmethod
    def from_device(cls, ctx, dev_ptr, n, p, ld=None):
        """The n x p column-major float64 matrix at DEVICE address ``dev_ptr`` on the context's GPU."""
        self = cls.__new__(cls)
        self._lib = L.require_device()
        self.ctx = ctx
        self.shape = (int(n), int(p))
        h = C.c_void_p()
        L.check(self._lib.resnmtf_data_create_device(ctx._h, int(n), int(p), C.c_void_p(int(dev_ptr)),
                                                     int(n if ld is None else ld), C.byref(h)))
        self._h = h
        return self

    @classmethod
    def _wrap(cls, ctx, handle, shape):
        self = cls.__new__(cls)
        self._lib = L.require_device()
        self.ctx = ctx
        self.shape = (int(shape[0]), int(shape[1]))
        self._h = handle
        return self

    @classmethod
    def prepped(cls, ctx, x):
        """Upload + make_non_neg_inner + matrix_normalisation (R/utils.r:20-27, 86-88) on the device.  Returns
        (handle, was_negative)."""
        lib = L.require_device()
        x = _f64(x)
        h, neg = C.c_void_p(), C.c_int32(0)
        L.check(lib.resnmtf_data_create_prepped(ctx._h, x.shape[0], x.shape[1], _ptr(x), x.shape[0], C.byref(neg),
                                                C.byref(h)))
        return cls._wrap(ctx, h, x.shape), bool(neg.value)

    def download(self):
        out = np.empty(self.shape, dtype=np.float64, order="F")
        L.check(self._lib.resnmtf_data_download(self._h, _ptr(out), self.shape[0]))
        return out

    def sums(self):
        """(colSums, rowSums) of the view."""
        cs = np.empty(self.shape[1], dtype=np.float64)
        rs = np.empty(self.shape[0], dtype=np.float64)
        L.check(self._lib.resnmtf_data_sums(self._h, _ptr(cs), _ptr(rs)))
        return cs, rs

    def shuffle(self, seed, renormalise=True):
        """shuffle_view (R/obtain_bicl.r:11-22) on the device: a new handle with all entries permuted (keyed by
        ``seed``), reshuffled until no row / column is all zero; ``renormalise``: the prep apply_resnmtf applies to
        the shuffled data."""
        h, tries = C.c_void_p(), C.c_int64(0)
        L.check(self._lib.resnmtf_data_shuffle(self._h, C.c_uint64(int(seed) & (2 ** 64 - 1)), int(bool(renormalise)),
                                               C.byref(tries), C.byref(h)))
        out = DeviceData._wrap(self.ctx, h, self.shape)
        out.attempts = int(tries.value)
        return out

    def subsample(self, rows, cols):
        """x[rows, cols] as a new handle (0-based indices), gathered on the device."""
        r = np.ascontiguousarray(rows, dtype=np.int32)
        c = np.ascontiguousarray(cols, dtype=np.int32)
        h = C.c_void_p()
        L.check(self._lib.resnmtf_data_subsample(self._h, _ptr(r), r.size, _ptr(c), c.size, C.byref(h)))
        return DeviceData._wrap(self.ctx, h, (r.size, c.size))

    def copy_to(self, ctx):
        """The view (and its cached SVD triplets) on another context's GPU, device to device."""
        h = C.c_void_p()
        L.check(self._lib.resnmtf_data_copy(self._h, ctx._h, C.byref(h)))
        return DeviceData._wrap(ctx, h, self.shape)

    def svd_topk(self, k):
        """|U[:, :k]|, d[:k], |V[:, :k]| of the view (what init_mats_inner takes from svd(x), R/update_steps.r:92-95),
        computed on the device and cached on the handle."""
        k = int(k)
        u = np.empty((self.shape[0], k), dtype=np.float64, order="F")
        d = np.empty(k, dtype=np.float64)
        v = np.empty((self.shape[1], k), dtype=np.float64, order="F")
        L.check(self._lib.resnmtf_data_svd_topk(self._h, k, _ptr(u), _ptr(d), _ptr(v)))
        return u, d, v

    def bisil(self, row_clustering, col_clustering, method="euclidean"):
        """Bisilhouette score of this view's biclustering with the distance blocks on the GPU (resnmtf_data_bisil).
        Returns dict(bisil=..., vals=[per bicluster that is non-empty])."""
        if method not in L.DISTANCES:
            raise ValueError("distance must be one of 'euclidean', 'manhattan' or 'cosine'.")
        rc = _f64(np.asarray(row_clustering) > 0)
        cc = _f64(np.asarray(col_clustering) > 0)
        k = rc.shape[1]
        vals = np.zeros(k, dtype=np.float64)
        out = C.c_double(0.0)
        L.check(self._lib.resnmtf_data_bisil(self._h, _ptr(rc), _ptr(cc), k, L.DISTANCES[method], _ptr(vals),
                                             C.byref(out)))
        live = [j for j in range(k) if rc[:, j].any() and cc[:, j].any()]
        return {"bisil": float(out.value), "vals": [float(vals[j]) for j in live]}

    def bisil_part(self, rc, cc, want, method="euclidean"):
        """Per-bicluster bisilhouette values of the biclusters flagged in ``want`` only (resnmtf_data_bisil_part); ``rc``
        / ``cc`` are the 0/1 float64 column-major cluster matrices.  Returns (vals[k], number of non-empty biclusters)."""
        k = rc.shape[1]
        vals = np.zeros(k, dtype=np.float64)
        flags = np.ascontiguousarray(want, dtype=np.int32)
        n_live = C.c_int32(0)
        L.check(self._lib.resnmtf_data_bisil_part(self._h, _ptr(rc), _ptr(cc), k, L.DISTANCES[method], _ptr(flags),
                                                  _ptr(vals), C.byref(n_live)))
        return vals, int(n_live.value)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.resnmtf_data_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceFit:
    """Device-resident state of one fit.  Shapes: view v is n[v] x p[v] with k[v] clusters.  ``ctx`` is a Context, or a
    list with one Context per view: the views are then placed on those GPUs (resnmtf_fit_create_placed) and the coupled
    factor rows travel over NVLink inside the update kernels."""

    def __init__(self, ctx, n, p, k):
        self._lib = L.require_device()
        self.view_ctx = list(ctx) if isinstance(ctx, (list, tuple)) else None
        if self.view_ctx is not None:
            ctx = self.view_ctx[0]
        self.ctx = ctx
        self.n = [int(x) for x in n]
        self.p = [int(x) for x in p]
        self.k = [int(x) for x in k]
        self.n_views = len(self.n)
        if not (len(self.p) == self.n_views == len(self.k)):
            raise ValueError("n, p, k must have one entry per view")
        V = self.n_views
        an = (C.c_int64 * V)(*self.n)
        ap = (C.c_int64 * V)(*self.p)
        ak = (C.c_int32 * V)(*self.k)
        h = C.c_void_p()
        if self.view_ctx is not None:
            if len(self.view_ctx) != V:
                raise ValueError("one context per view")
            actx = (C.c_void_p * V)(*[c._h.value for c in self.view_ctx])
            L.check(self._lib.resnmtf_fit_create_placed(actx, V, an, ap, ak, C.byref(h)))
        else:
            L.check(self._lib.resnmtf_fit_create(ctx._h, V, an, ap, ak, C.byref(h)))
        self._h = h

    # ---- inputs ------------------------------------------------------------------------------
    def set_data(self, v, x):
        """Host matrix (any layout; converted to column-major float64 if it is not already)."""
        x = _f64(x)
        if x.shape != (self.n[v], self.p[v]):
            raise ValueError(f"view {v}: expected shape {(self.n[v], self.p[v])}, got {x.shape}")
        L.check(self._lib.resnmtf_fit_set_data(self._h, v, _ptr(x), x.shape[0]))

    def attach_data(self, v, data):
        """Use a shared DeviceData as view v (no copy)."""
        if tuple(data.shape) != (self.n[v], self.p[v]):
            raise ValueError(f"view {v}: expected shape {(self.n[v], self.p[v])}, got {tuple(data.shape)}")
        L.check(self._lib.resnmtf_fit_attach_data(self._h, v, data._h))

    def set_data_device(self, v, dev_ptr, ld):
        """Column-major float64 matrix already resident on this context's GPU."""
        L.check(self._lib.resnmtf_fit_set_data_device(self._h, v, C.c_void_p(int(dev_ptr)), int(ld)))

    def set_factors(self, v, f, s, g, lam=None, mu=None):
        f, s, g = _f64(f), _f64(s), _f64(g)
        k = self.k[v]
        if f.shape != (self.n[v], k) or g.shape != (self.p[v], k) or s.shape != (k, k):
            raise ValueError(f"view {v}: factor shapes do not match n={self.n[v]}, p={self.p[v]}, k={k}")
        lam = None if lam is None else np.ascontiguousarray(lam, dtype=np.float64)
        mu = None if mu is None else np.ascontiguousarray(mu, dtype=np.float64)
        L.check(self._lib.resnmtf_fit_set_factors(self._h, v, _ptr(f), _ptr(s), _ptr(g), _ptr(lam), _ptr(mu)))

    def set_restrictions(self, phi=None, xi=None, psi=None):
        mats = [None if m is None else _f64(m) for m in (phi, xi, psi)]
        for m in mats:
            if m is not None and m.shape != (self.n_views, self.n_views):
                raise ValueError("restriction matrices must be n_views x n_views")
        L.check(self._lib.resnmtf_fit_set_restrictions(self._h, *[_ptr(m) for m in mats]))

    def set_shared_map(self, kind, v, w, idx_v, idx_w):
        iv = np.ascontiguousarray(idx_v, dtype=np.int32)
        iw = np.ascontiguousarray(idx_w, dtype=np.int32)
        if iv.shape != iw.shape:
            raise ValueError("idx_v and idx_w must have the same length")
        L.check(self._lib.resnmtf_fit_set_shared_map(self._h, int(kind), int(v), int(w), _ptr(iv), _ptr(iw), iv.size))

    def set_options(self, err_mode=L.ERR_AUTO, impl=L.IMPL_AUTO):
        L.check(self._lib.resnmtf_fit_set_options(self._h, int(err_mode), int(impl)))

    # ---- the loop ----------------------------------------------------------------------------
    def run(self, n_iters=None, tol=1.0e-6, max_iters=0):
        """n_iters=None: until |d mean err| <= tol (R/main.r:55-81); else exactly n_iters sweeps."""
        done = C.c_int64(0)
        ni = -1 if n_iters is None else int(n_iters)
        L.check(self._lib.resnmtf_fit_run(self._h, ni, float(tol), int(max_iters), C.byref(done)))
        return int(done.value)

    def step(self):
        L.check(self._lib.resnmtf_fit_step(self._h))

    def profile(self, n_iters):
        ms = (C.c_double * 5)()
        cnt = (C.c_int64 * 5)()
        L.check(self._lib.resnmtf_fit_profile(self._h, int(n_iters), ms, cnt))
        names = ("f_step", "g_stream", "fused_step", "residual", "finish")
        return {nm: {"ms": float(ms[i]), "intervals": int(cnt[i])} for i, nm in enumerate(names)}

    # ---- outputs -----------------------------------------------------------------------------
    def get_factors(self, v):
        k = self.k[v]
        f = np.empty((self.n[v], k), dtype=np.float64, order="F")
        g = np.empty((self.p[v], k), dtype=np.float64, order="F")
        s = np.empty((k, k), dtype=np.float64, order="F")
        lam = np.empty(k, dtype=np.float64)
        mu = np.empty(k, dtype=np.float64)
        L.check(self._lib.resnmtf_fit_get_factors(self._h, v, _ptr(f), _ptr(s), _ptr(g), _ptr(lam), _ptr(mu)))
        return f, s, g, lam, mu

    def normalise(self):
        L.check(self._lib.resnmtf_fit_normalise(self._h))

    def errors(self):
        cnt = C.c_int64(0)
        L.check(self._lib.resnmtf_fit_get_errors(self._h, None, 0, C.byref(cnt)))
        out = np.empty(int(cnt.value), dtype=np.float64)
        if out.size:
            L.check(self._lib.resnmtf_fit_get_errors(self._h, _ptr(out), out.size, C.byref(cnt)))
        return out

    def view_errors(self):
        err = np.empty(self.n_views, dtype=np.float64)
        norms = np.empty(self.n_views, dtype=np.float64)
        L.check(self._lib.resnmtf_fit_get_view_errors(self._h, _ptr(err), _ptr(norms)))
        return err, norms

    def counters(self):
        c = L.Counters()
        L.check(self._lib.resnmtf_fit_get_counters(self._h, C.byref(c)))
        return {name: getattr(c, name) for name, _ in L.Counters._fields_}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.resnmtf_fit_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
