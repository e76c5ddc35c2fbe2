"""Thin object layer over the pool / unit entry points of the C ABI (``resnmtf_pool_*``, ``resnmtf_batch_run``): the
independent fits of one ``apply_resnmtf`` call -- k-sweep fits, shuffled refits, stability resamples -- described as
units and run by the library's own worker threads, one per GPU.  Nothing here computes: units go in, factors come out."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .device import Context, DeviceData, _f64


class _BorrowedContext(Context):
    """A context owned by a pool: same interface, never destroyed from here."""

    def __init__(self, lib, handle):  # noqa: D401 - no resnmtf_ctx_create on purpose
        self._lib = lib
        self._h = C.c_void_p(handle)
        self.device = int(lib.resnmtf_ctx_device(self._h))

    def close(self):
        self._h = C.c_void_p()


def _ptr_array(arrays):
    """(void* array, keep-alive list) for a list of numpy arrays (None entries become NULL)."""
    arr = (C.c_void_p * len(arrays))(*[None if a is None else a.ctypes.data for a in arrays])
    return arr, arrays


class NativePool:
    def __init__(self, n_gpus=0, devices=None):
        self._lib = L.require_device()
        h = C.c_void_p()
        if devices is not None:
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            L.check(self._lib.resnmtf_pool_create(devs, len(devices), C.byref(h)))
        else:
            L.check(self._lib.resnmtf_pool_create(None, int(n_gpus), C.byref(h)))
        self._h = h
        self.size = int(self._lib.resnmtf_pool_size(h))
        self.shapes = {}  # key -> [(n, p)] per view
        self.contexts = [_BorrowedContext(self._lib, self._lib.resnmtf_pool_ctx(h, g)) for g in range(self.size)]

    def __len__(self):
        return self.size

    # ---- data sets ----------------------------------------------------------------------------------
    def put_host(self, key, views, prep=False):
        """Uploads the views (host matrices) to the first GPU, with ``prep`` through make_non_neg_inner and
        matrix_normalisation on the device.  Returns whether a negative entry was seen (the reference's warning)."""
        xs = [_f64(x) for x in views]
        V = len(xs)
        n = (C.c_int64 * V)(*[x.shape[0] for x in xs])
        p = (C.c_int64 * V)(*[x.shape[1] for x in xs])
        ptrs, _keep = _ptr_array(xs)
        neg = C.c_int32(0)
        L.check(self._lib.resnmtf_pool_put_host(self._h, int(key), V, n, p, ptrs, None, int(bool(prep)), C.byref(neg)))
        self.shapes[key] = [x.shape for x in xs]
        return bool(neg.value)

    def put(self, key, handles):
        arr = (C.c_void_p * len(handles))(*[h._h.value for h in handles])
        L.check(self._lib.resnmtf_pool_put(self._h, int(key), len(handles), arr))
        self.shapes[key] = [tuple(h.shape) for h in handles]

    def get(self, key, gpu=0, view=0):
        """The pool's handle of a view on one of its GPUs (borrowed: do not close)."""
        h = C.c_void_p()
        L.check(self._lib.resnmtf_pool_get(self._h, int(key), int(gpu), int(view), C.byref(h)))
        d = DeviceData._wrap(self.contexts[gpu], h, self.shapes[key][view])
        d.close = lambda: None  # owned by the pool
        return d

    def drop(self, key):
        L.check(self._lib.resnmtf_pool_drop(self._h, int(key)))
        self.shapes.pop(key, None)

    # ---- units ----------------------------------------------------------------------------------------
    def run(self, units):
        """``units``: list of dicts -- key, k (per view), and optionally rows / cols (per view index arrays: sub-sample),
        shuffle_seed (int: shuffle every view), renormalise, init_f / init_s / init_g (per view) or noise (per view
        k x k), phi / xi / psi, maps [(kind, v, w, idx_v, idx_w)], n_iters (None: convergence), tol, max_iters,
        err_mode, impl, errors_cap.  Returns one dict per unit: output_f / output_s / output_g (normalised), lambda,
        mu (as the loop left them), total_err, iters, gpu, seconds."""
        n_units = len(units)
        if n_units == 0:
            return []
        arr = (L.Unit * n_units)()
        keep, outs = [], []
        for i, u in enumerate(units):
            cu = arr[i]
            key = u["key"]
            shapes = list(self.shapes[key])
            V = len(shapes)
            k = np.ascontiguousarray(u["k"], dtype=np.int32)
            if k.size != V:
                raise ValueError("unit: k must have one entry per view")
            cu.data_key = int(key)
            derive = 0
            if u.get("rows") is not None:
                rows = [np.ascontiguousarray(r, dtype=np.int32) for r in u["rows"]]
                cols = [np.ascontiguousarray(c, dtype=np.int32) for c in u["cols"]]
                rp, _ = _ptr_array(rows)
                cp, _ = _ptr_array(cols)
                nr = (C.c_int64 * V)(*[r.size for r in rows])
                nc = (C.c_int64 * V)(*[c.size for c in cols])
                cu.rows, cu.cols = C.cast(rp, C.c_void_p), C.cast(cp, C.c_void_p)
                cu.n_rows, cu.n_cols = C.cast(nr, C.c_void_p), C.cast(nc, C.c_void_p)
                keep += [rows, cols, rp, cp, nr, nc]
                shapes = [(r.size, c.size) for r, c in zip(rows, cols)]
                derive |= L.DERIVE_SUBSAMPLE
            if u.get("shuffle_seed") is not None:
                derive |= L.DERIVE_SHUFFLE
                cu.seed = int(u["shuffle_seed"]) & (2 ** 64 - 1)
                cu.renormalise = int(bool(u.get("renormalise", True)))
            cu.derive = derive
            cu.k = k.ctypes.data
            keep.append(k)
            if u.get("init_f") is not None:
                for name in ("init_f", "init_s", "init_g"):
                    mats = [_f64(m) for m in u[name]]
                    pa, _ = _ptr_array(mats)
                    setattr(cu, name, C.cast(pa, C.c_void_p))
                    keep += [mats, pa]
            elif u.get("noise") is not None:
                mats = [_f64(m) for m in u["noise"]]
                pa, _ = _ptr_array(mats)
                cu.noise = C.cast(pa, C.c_void_p)
                keep += [mats, pa]
            for name in ("phi", "xi", "psi"):
                m = u.get(name)
                if m is not None:
                    m = _f64(m)
                    setattr(cu, name, m.ctypes.data)
                    keep.append(m)
            maps = u.get("maps") or []
            if maps:
                marr = (L.Map * len(maps))()
                for j, (kind, v, w, iv, iw) in enumerate(maps):
                    iv = np.ascontiguousarray(iv, dtype=np.int32)
                    iw = np.ascontiguousarray(iw, dtype=np.int32)
                    marr[j].kind, marr[j].v, marr[j].w = int(kind), int(v), int(w)
                    marr[j].idx_v, marr[j].idx_w, marr[j].len = iv.ctypes.data, iw.ctypes.data, iv.size
                    keep += [iv, iw]
                cu.maps = C.cast(marr, C.c_void_p)
                cu.n_maps = len(maps)
                keep.append(marr)
            n_iters = u.get("n_iters")
            cu.n_iters = -1 if n_iters is None else int(n_iters)
            cu.tol = float(u.get("tol", 1.0e-6))
            cu.max_iters = int(u.get("max_iters", 0) or 0)
            cu.err_mode = int(u.get("err_mode", L.ERR_AUTO))
            cu.impl = int(u.get("impl", L.IMPL_AUTO))
            f = [np.empty((n, int(kk)), dtype=np.float64, order="F") for (n, _), kk in zip(shapes, k)]
            g = [np.empty((p, int(kk)), dtype=np.float64, order="F") for (_, p), kk in zip(shapes, k)]
            s = [np.empty((int(kk), int(kk)), dtype=np.float64, order="F") for kk in k]
            lam = [np.empty(int(kk), dtype=np.float64) for kk in k]
            mu = [np.empty(int(kk), dtype=np.float64) for kk in k]
            cap = int(u.get("errors_cap", 0) or (cu.n_iters if cu.n_iters > 0 else max(cu.max_iters, 20000)))
            errs = np.empty(max(cap, 1), dtype=np.float64)
            for name, mats in (("out_f", f), ("out_s", s), ("out_g", g), ("out_lambda", lam), ("out_mu", mu)):
                pa, _ = _ptr_array(mats)
                setattr(cu, name, C.cast(pa, C.c_void_p))
                keep.append(pa)
            cu.errors = errs.ctypes.data
            cu.errors_cap = errs.size
            outs.append((f, s, g, lam, mu, errs))
        rc = self._lib.resnmtf_batch_run(self._h, arr, n_units)
        L.check(rc)
        results = []
        for i, (f, s, g, lam, mu, errs) in enumerate(outs):
            cu = arr[i]
            results.append({"output_f": f, "output_s": s, "output_g": g, "lambda": lam, "mu": mu,
                            "total_err": errs[:min(int(cu.n_errors), errs.size)].copy(), "iters": int(cu.iters),
                            "gpu": int(cu.gpu), "seconds": float(cu.seconds)})
        del keep
        return results

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            for c in self.contexts:
                c.close()
            self._lib.resnmtf_pool_destroy(self._h)
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
