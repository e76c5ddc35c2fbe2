"""The matrix-sized route of ``apply_resnmtf`` / ``res_nmtf_inner`` on the library alone (no torch): the views live on the
GPUs of a native pool (``resnmtf_pool_*``), every convergence loop of the call -- k-sweep fits, shuffled refits,
stability resamples (R/main.r:270-299, R/obtain_bicl.r:31-42, R/stability_analysis.r:302-338) -- is a *unit* of
``resnmtf_batch_run`` (derive the data on the device, SVD initialisation on the device, the update loop, normalisation;
one native worker thread per GPU), and a fit's post-processing (R/obtain_bicl.r:151-204) takes its JSD thresholds from
``resnmtf_jsd_pairs`` and its bisilhouette distance blocks from ``resnmtf_data_bisil`` on the resident view.  The host
keeps what the reference keeps on k x k / n x k outputs: naming, index maps, threshold densities, Jaccard relevance.

Every unit draws from its own child generator (per k, per shuffled repeat, per resample), in a fixed order on the host,
so the result of a call does not depend on the number of GPUs or on which GPU ran which unit."""
from __future__ import annotations

import os
import sys
import time
import warnings
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib as L
from . import prep
from .native import NativePool
from .prep import NamedMatrix

TRACE = os.environ.get("RESNMTF_TRACE", "0") not in ("", "0")


def _trace(label, t0, units=None):
    """RESNMTF_TRACE=1: wall time of a phase of the call and, for a batch of units, the busy seconds of every GPU."""
    if not TRACE:
        return
    msg = f"  [resnmtf trace] {label:44s} {time.perf_counter() - t0:7.3f} s"
    if units:
        per = {}
        for u in units:
            per.setdefault(u["gpu"], []).append(u["seconds"])
        msg += f"  {len(units)} units, longest {max(u['seconds'] for u in units):.3f} s, busy " + " ".join(
            f"g{g}={sum(v):.2f}({len(v)})" for g, v in sorted(per.items()))
    print(msg, file=sys.stderr, flush=True)


KEY_DATA = 0
SIGMA = 0.05  # variance of the noise term of init_mats_inner (R/update_steps.r:96-99)


class ResidentView(NamedMatrix):
    """A view that exists on the GPUs only: shape and names, no host values."""

    __slots__ = ("_shape",)

    def __init__(self, shape, rownames=None, colnames=None):  # noqa: D401 - no host array on purpose
        self.x = None
        self._shape = (int(shape[0]), int(shape[1]))
        self.rownames = None if rownames is None else list(rownames)
        self.colnames = None if colnames is None else list(colnames)

    @property
    def shape(self):
        return self._shape

    def copy(self):
        return ResidentView(self._shape, self.rownames, self.colnames)


def _noise(rng, k_vec):
    """abs(MASS::mvrnorm(k, 0, sigma I_k)) per view, in view order (R/update_steps.r:96-99)."""
    return [np.abs(np.sqrt(SIGMA) * rng.standard_normal((int(k), int(k)))) for k in k_vec]


def _names_or_default(data):
    rn, cn = [], []
    for v, m in enumerate(data):
        rn.append(m.rownames if m.rownames is not None else [f"__v{v}_r{i}" for i in range(m.shape[0])])
        cn.append(m.colnames if m.colnames is not None else [f"__v{v}_c{i}" for i in range(m.shape[1])])
    return rn, cn


def _maps(data, row_indices, col_indices):
    rn, cn = _names_or_default(data)
    out = []
    for (v, w), (iv, iw) in prep.shared_maps(row_indices, rn).items():
        out.append((L.MAP_ROW, v, w, iv, iw))
    for (v, w), (iv, iw) in prep.shared_maps(col_indices, cn).items():
        out.append((L.MAP_COL, v, w, iv, iw))
    return out


import threading  # noqa: E402

_POOLS = {}  # device tuple -> NativePool kept for the life of the process (release_pools() lets go of them)
_POOLS_LOCK = threading.Lock()


def _shared_pool(n_gpus, devices):
    """One native pool per set of GPUs for the whole process: creating a pool (contexts, streams, private memory pools)
    and above all destroying it -- every block of the memory pools goes back to the driver, 0.6 s per call on one GPU,
    measured -- were serial time inside every apply_resnmtf call; kept, the pool also keeps its device memory and the
    loaded kernels warm for the next call.  RESNMTF_KEEP_POOLS=0 goes back to one pool per call."""
    key = ("n", int(n_gpus)) if devices is None else ("d",) + tuple(int(d) for d in devices)
    with _POOLS_LOCK:
        pool = _POOLS.get(key)
        if pool is None or not pool._h.value:
            pool = NativePool(n_gpus=n_gpus, devices=devices)
            pool.call_lock = threading.Lock()  # one apply_resnmtf call at a time on a shared pool
            _POOLS[key] = pool
    return pool


def release_pools():
    """Destroys the cached pools (their contexts, streams and device memory)."""
    for pool in list(_POOLS.values()):
        pool.close()
    _POOLS.clear()


import atexit  # noqa: E402

atexit.register(release_pools)


class _SplitBisil:
    """Stands in for a resident view in obtain_biclusters(): the bisilhouette of ONE fit with its biclusters dealt over
    the GPUs of the pool (every GPU holds a copy of the view; the per-bicluster values are independent).  The
    post-processing of the six sweep fits sits between the two batches of a call, and the fit with the largest k
    carries a third of its distance work: scored whole on one GPU each, the six fits took 0.42 s on 8 GPUs against
    0.44 s for the whole sweep batch.  The values are combined in bicluster order, so the score is bit-identical to the
    single-GPU call."""

    def __init__(self, runner, view, first_gpu):
        self.runner, self.view, self.first_gpu = runner, view, int(first_gpu)

    def bisil(self, row_clustering, col_clustering, method="euclidean"):
        if method not in L.DISTANCES:
            raise ValueError("distance must be one of 'euclidean', 'manhattan' or 'cosine'.")
        rc = np.asfortranarray(np.asarray(row_clustering) > 0, dtype=np.float64)
        cc = np.asfortranarray(np.asarray(col_clustering) > 0, dtype=np.float64)
        k = rc.shape[1]
        nr, ncl = rc.sum(axis=0), cc.sum(axis=0)
        live = [j for j in range(k) if nr[j] > 0 and ncl[j] > 0]
        if not live:
            return {"bisil": 0.0, "vals": []}
        n_gpus = len(self.runner.pool)
        # distance work of bicluster j ~ |R_j| x (rows of all live clusters) x |C_j|; largest first onto the least loaded GPU
        rows_all = float(sum(nr[j] for j in live))
        cost = {j: float(nr[j]) * rows_all * float(ncl[j]) for j in live}
        # the load table is the runner's, shared by the fits whose post-processing runs at the same time: dealt fit by
        # fit against its own table, the six fits of a sweep left the busiest GPU with 0.21 s of the 0.9 s of distance
        # work on 8 GPUs (ideal 0.11 s); where a piece runs does not change its value
        mine = [[] for _ in range(n_gpus)]
        with self.runner.bisil_lock:
            load = self.runner.bisil_load
            for j in sorted(live, key=lambda j: -cost[j]):
                g = min(range(n_gpus), key=lambda g: (load[g], (g - self.first_gpu) % n_gpus))
                mine[g].append(j)
                load[g] += cost[j]
        def piece(g):
            want = np.zeros(k, dtype=np.int32)
            want[mine[g]] = 1
            with self.runner.gpu_locks[g]:  # entry points on one context are not re-entrant
                return self.runner.handle(self.view, g).bisil_part(rc, cc, want, method)[0]

        busy = [g for g in range(n_gpus) if mine[g]]
        vals = np.zeros(k, dtype=np.float64)
        # the pieces of this fit run at the same time on their GPUs (ctypes releases the interpreter lock); the pieces
        # of the other fits' host threads queue on the same per-GPU locks
        with ThreadPoolExecutor(max_workers=len(busy)) as ex:
            for part in ex.map(piece, busy):
                vals += part  # disjoint supports: exact
        total = 0.0
        for j in live:  # in bicluster order, as resnmtf_data_bisil accumulates them
            total += float(vals[j])
        return {"bisil": total / len(live), "vals": [float(vals[j]) for j in live]}


class _LockedBisil:
    """The whole bisilhouette of a fit on one GPU of the pool, under that GPU's lock."""

    def __init__(self, runner, view, gpu):
        self.runner, self.view, self.gpu = runner, view, int(gpu)

    def bisil(self, row_clustering, col_clustering, method="euclidean"):
        with self.runner.gpu_locks[self.gpu]:
            return self.runner.handle(self.view, self.gpu).bisil(row_clustering, col_clustering, method=method)


class NativeRunner:
    """The pool of one call and the unit bookkeeping of its fits."""

    def __init__(self, n_gpus=0, devices=None):
        self.keep = os.environ.get("RESNMTF_KEEP_POOLS", "1") not in ("", "0")
        self.pool = _shared_pool(n_gpus, devices) if self.keep else NativePool(n_gpus=n_gpus, devices=devices)
        self._held = False
        if self.keep:
            self.pool.call_lock.acquire()
            self._held = True
        # the post-processing of several fits runs on host threads that share the pool's contexts (JSD pair kernel,
        # bisilhouette pieces): one lock per GPU, entry points on one context are not re-entrant
        self.gpu_locks = [threading.Lock() for _ in range(len(self.pool))]
        self.bisil_lock = threading.Lock()
        self.bisil_load = [0.0] * len(self.pool)  # distance work already dealt to every GPU (_SplitBisil), per batch of fits
        for g, c in enumerate(self.pool.contexts):
            c.lock = self.gpu_locks[g]

    def close(self):
        t0 = time.perf_counter()
        if self.keep:
            try:
                if KEY_DATA in self.pool.shapes:
                    self.pool.drop(KEY_DATA)  # the views of this call; the pool itself stays
            finally:
                if self._held:
                    self._held = False
                    self.pool.call_lock.release()
        else:
            self.pool.close()
        _trace("pool released" if not self.keep else "data set dropped (pool kept)", t0)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def put(self, views, prep_values):
        """Uploads the views (NamedMatrix list) under KEY_DATA; with ``prep_values`` through make_non_neg_inner and
        matrix_normalisation on the device (same warning as the host version, R/utils.r:20-27)."""
        t0 = time.perf_counter()
        neg = self.pool.put_host(KEY_DATA, [m.x for m in views], prep=prep_values)
        _trace("upload + prep on the first GPU", t0)
        if neg:
            warnings.warn("Matrix is not non-negative. Has been made non-negative.")

    def handle(self, view, gpu=0):
        return self.pool.get(KEY_DATA, gpu, view)

    # ---- fits --------------------------------------------------------------------------------------------------------
    def run_fits(self, specs, phi, xi, psi, n_iters, num_repeats, spurious, distance, no_clusts, max_iters=0,
                 err_mode=L.ERR_AUTO, impl=L.IMPL_AUTO):
        """The res_nmtf_inner calls described by ``specs`` (see api.run_fits), as one batch of units + post-processing.
        A spec may carry ``sub = (rows per view, cols per view)``: the fit runs on that sub-sample of the data set."""
        from .api import _fit_post

        need_shuffles = bool(spurious) and not no_clusts
        units, core_at, shuffle_at = [], [], []
        for sp in specs:
            data = sp["data"]
            V = len(data)
            explicit = sp.get("init_f") is not None and sp.get("init_s") is not None and sp.get("init_g") is not None
            k_vec = ([int(np.asarray(f).shape[1]) for f in sp["init_f"]] if explicit
                     else [int(k) for k in sp["k_vec"]])
            base = dict(key=KEY_DATA)
            if sp.get("sub") is not None:
                base["rows"], base["cols"] = sp["sub"]
            core = dict(base, k=k_vec, phi=phi, xi=xi, psi=psi, maps=_maps(data, sp["row_indices"], sp["col_indices"]),
                        n_iters=n_iters, max_iters=max_iters, err_mode=err_mode, impl=impl)
            if explicit:
                core.update(init_f=sp["init_f"], init_s=sp["init_s"], init_g=sp["init_g"])
            else:
                core["noise"] = _noise(sp["rng"], k_vec)
            core_at.append(len(units))
            units.append(core)
            mine = []
            if need_shuffles:
                # one repeat of obtain_shuffled_f (R/obtain_bicl.r:33-40): shuffle every view, re-prep, refit with
                # k = n_clusts for all views, no restrictions, to convergence
                for srng in sp["rng"].spawn(int(num_repeats)):
                    seed = int(srng.integers(0, 2 ** 63 - 1))
                    mine.append(len(units))
                    units.append(dict(base, k=[k_vec[0]] * V, shuffle_seed=seed, renormalise=True,
                                      noise=_noise(srng, [k_vec[0]] * V), n_iters=None, max_iters=max_iters))
            shuffle_at.append(mine)
        t0 = time.perf_counter()
        done = self.pool.run(units)
        _trace(f"batch of units ({len(specs)} fits{' + shuffled refits' if need_shuffles else ''})", t0, done)
        t0 = time.perf_counter()

        def post(si):
            sp = specs[si]
            res = done[core_at[si]]
            core = {"output_f": res["output_f"], "output_s": res["output_s"], "output_g": res["output_g"],
                    "total_err": res["total_err"], "lambda": res["lambda"], "mu": res["mu"],
                    "counters": {"iters": res["iters"], "gpu": res["gpu"], "seconds": res["seconds"]}}
            gpu = si % len(self.pool)
            want_bisil = sp.get("want_bisil", True) and not no_clusts
            split = len(self.pool) > 1 and os.environ.get("RESNMTF_SPLIT_BISIL", "1") not in ("", "0")
            resident = ([(_SplitBisil(self, v, gpu) if split else _LockedBisil(self, v, gpu)) for v in range(len(sp["data"]))]
                        if (want_bisil and sp.get("sub") is None) else [None] * len(sp["data"]))
            return _fit_post(core, sp["data"], n_iters, num_repeats, spurious, distance, no_clusts, rng=sp["rng"],
                             ctx=self.pool.contexts[gpu],
                             shuffled_f=[done[i]["output_f"] for i in shuffle_at[si]] if need_shuffles else None,
                             resident=resident, want_bisil=sp.get("want_bisil", True))

        with self.bisil_lock:
            self.bisil_load[:] = [0.0] * len(self.pool)
        if len(specs) == 1:
            out = [post(0)]
        else:
            # the fits with the most biclusters carry the most distance work: they start first (with fewer host threads than
            # fits -- one or two GPUs -- the k = 8 fit of a sweep was otherwise the last to start)
            first = sorted(range(len(specs)), key=lambda si: -int(np.asarray(done[core_at[si]]["output_f"][0]).shape[1]))
            out = [None] * len(specs)
            with ThreadPoolExecutor(max_workers=max(1, min(len(specs), 2 * len(self.pool)))) as ex:
                for si, res in zip(first, ex.map(post, first)):
                    out[si] = res
        _trace("post-processing (JSD thresholds, binarise, bisilhouette)", t0)
        return out

    # ---- stability analysis (R/stability_analysis.r:302-338) -----------------------------------------------------------
    def stability_check(self, data, results, k, phi, xi, psi, n_iters, spurious, num_repeats, no_clusts, distance,
                        sample_rate=0.9, n_stability=5, stab_thres=0.6, remove_unstable=True, rng=None, max_iters=0):
        from .stability import draw_subsample, number_biclusters, relevance_results

        if number_biclusters(results) == 0:
            print("No biclusters detected!")
            return results
        k = int(np.atleast_1d(k)[0])
        n_views = len(data)
        shapes = [m.shape for m in data]
        dim_1 = shapes[0]
        rng = np.random.default_rng() if rng is None else rng
        child_rngs = rng.spawn(int(n_stability))

        def sums(i, rows, cols):  # column / row sums of a sub-sample, gathered and reduced on the device
            sub = self.handle(i).subsample(rows, cols)
            try:
                return sub.sums()
            finally:
                sub.close()

        t0 = time.perf_counter()
        samples = [draw_subsample(sums, shapes, dim_1, n_views, sample_rate, child_rngs[i])
                   for i in range(int(n_stability))]
        _trace("stability: sub-sample draws + sums on the device", t0)
        if any(smp is None for smp in samples):
            return results
        specs = []
        for i, (row_s, col_s) in enumerate(samples):
            new_data = [ResidentView((len(row_s[v]), len(col_s[v])), [data[v].rownames[r] for r in row_s[v]],
                                     [data[v].colnames[c] for c in col_s[v]]) for v in range(n_views)]
            reordered = prep.reorder_data(new_data, n_views, [m.rownames for m in new_data],
                                          [m.colnames for m in new_data])
            specs.append(dict(data=new_data, sub=(row_s, col_s), row_indices=reordered["row_indices"],
                              col_indices=reordered["col_indices"], k_vec=[k] * n_views, rng=child_rngs[i],
                              want_bisil=False))  # the resample fits are read for their clusters only (:268-276)
        new_results = self.run_fits(specs, phi, xi, psi, n_iters, num_repeats, spurious, distance, False,
                                    max_iters=max_iters)
        relevance = np.zeros((n_views, k))
        for (row_s, col_s), new_res in zip(samples, new_results):  # in repeat order, as the reference accumulates them
            for i in range(n_views):
                relevance[i, :] += relevance_results(new_res["row_clusters"][i], new_res["col_clusters"][i],
                                                     results["row_clusters"][i][row_s[i], :],
                                                     results["col_clusters"][i][col_s[i], :])
        relevance = relevance / n_stability
        if not remove_unstable:
            return {"res": results, "relevance": relevance}
        for i in range(n_views):
            unstable = relevance[i, :] < stab_thres
            results["row_clusters"][i][:, unstable] = 0.0
            results["col_clusters"][i][:, unstable] = 0.0
        return results
