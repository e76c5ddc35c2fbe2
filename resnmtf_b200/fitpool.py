"""Placement of the independent fits of one ``apply_resnmtf`` call on the visible GPUs (first row).

The reference intends a ``foreach %dopar%`` over the k values of the sweep (R/main.r:288-299; unreachable there) and
runs everything else serially: the ``num_repeats`` shuffled refits behind every fit (R/obtain_bicl.r:31-42) and the
``n_stability`` resample fits with their own shuffled refits (R/stability_analysis.r:302-338) -- 66 convergence loops
per default call, all independent given their inputs.  Here every one of them is a *unit*: a callable that takes a
``Worker`` (one GPU: one library context, the views resident on that GPU) and returns factors.  ``FitPool.run`` hands
the units of a phase to one host thread per GPU, longest first, each thread pulling the next unit when it is free;
ctypes and torch release the interpreter lock inside the library / CUDA calls.  Every unit draws from its own child
generator, so what a unit returns does not depend on where or when it ran: the result of a call is bit-identical for
any number of GPUs.  No collective is involved; the only inter-GPU traffic is the one-off copy of the views.
"""
from __future__ import annotations

import os
import threading
import time

import numpy as np

from .device import DeviceData
from .prep import NamedMatrix


class ResidentMatrix(NamedMatrix):
    """A view that exists on the GPUs only (a sub-sample gathered on the device): shape and names, no host values."""

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
        return ResidentMatrix(self._shape, self.rownames, self.colnames)


class Worker:
    """One GPU of the pool: a library context plus what is resident on that GPU for the current call."""

    def __init__(self, pool, ctx, index):
        self.pool = pool
        self.ctx = ctx
        self.index = index
        self.views = {}    # data key -> [device tensor per view], p x n row-major (== column-major n x p)
        self.handles = {}  # data key -> [DeviceData per view] (library layout, shared by the fits of that data)
        self.eig = {}      # data key -> {view: top singular triplets}, for data fitted more than once (the k-sweep)

    def get_views(self, key):
        if key not in self.views:
            self.views[key] = self.pool.loaders[key](self)
        return self.views[key]

    def get_handles(self, key):
        """The views of ``key`` in the library's layout, created once per GPU and shared by every fit of that data."""
        if key not in self.handles:
            if key in self.pool.loaders:
                import torch

                xs = self.get_views(key)
                torch.cuda.current_stream(xs[0].device).synchronize()  # the library reads on its own stream
                self.handles[key] = [DeviceData.from_device(self.ctx, x.data_ptr(), x.shape[1], x.shape[0]) for x in xs]
            else:
                self.handles[key] = [DeviceData(self.ctx, m.x) for m in self.pool.host_data[key]]
        return self.handles[key]

    def eig_cache(self, key):
        return self.eig.setdefault(key, {})

    def clear(self):
        for hs in self.handles.values():
            for h in hs:
                h.close()
        self.handles.clear()
        self.views.clear()
        self.eig.clear()


class FitPool:
    def __init__(self, contexts):
        self.workers = [Worker(self, c, i) for i, c in enumerate(contexts)]
        self.loaders = {}    # data key -> callable(worker) -> [device tensor per view]
        self.host_data = {}  # data key -> [NamedMatrix] (views that live on the host)
        # RESNMTF_TRACE=1: wall time of every unit and phase, printed when the pool closes (tools/ksweep_wall.py)
        self.trace = [] if os.environ.get("RESNMTF_TRACE", "0") not in ("", "0") else None

    def __len__(self):
        return len(self.workers)

    # ---- data ---------------------------------------------------------------------------------------
    def place_host(self, key, data):
        self.host_data[key] = data

    def place_resident(self, key, data):
        """Uploads the host views once (to the first GPU) and copies them from there to the other GPUs."""
        import torch

        from .api import torch_device

        self.host_data[key] = data
        w0 = self.workers[0]
        dev0 = torch_device(w0.ctx.device)
        first = [torch.from_numpy(np.ascontiguousarray(m.x.T)).to(dev0) for m in data]
        w0.views[key] = first
        for w in self.workers[1:]:
            w.views[key] = [t.to(torch_device(w.ctx.device)) for t in first]
        for w in self.workers:
            torch.cuda.synchronize(w.views[key][0].device)
        self.loaders[key] = lambda worker: worker.views[key]

    def place_gather(self, key, base_key, rows, cols):
        """Views of ``key`` = rows / columns ``rows[v]`` / ``cols[v]`` of the resident views of ``base_key``, gathered
        on whichever GPU asks for them (the sub-samples of the stability analysis never visit the host)."""
        def load(worker):
            import torch

            out = []
            for v, xt in enumerate(worker.get_views(base_key)):
                r = torch.from_numpy(np.asarray(rows[v], dtype=np.int64)).to(xt.device)
                c = torch.from_numpy(np.asarray(cols[v], dtype=np.int64)).to(xt.device)
                out.append(xt.index_select(0, c).index_select(1, r).contiguous())
            return out

        self.loaders[key] = load

    # ---- work ---------------------------------------------------------------------------------------
    def run(self, tasks):
        """``tasks``: [(cost, callable(worker) -> result)].  Returns the results in task order.  Units are started
        longest first; with one GPU they simply run one after the other on the calling thread."""
        order = sorted(range(len(tasks)), key=lambda i: (-float(tasks[i][0]), i))
        results = [None] * len(tasks)
        if self.trace is not None:
            tasks = [(c, self._timed(fn)) for c, fn in tasks]
            t_phase = time.perf_counter()
            try:
                return self._run(tasks, order, results)
            finally:
                self.trace.append(("phase", len(tasks), time.perf_counter() - t_phase))
        return self._run(tasks, order, results)

    def _timed(self, fn):
        def run(worker):
            t0 = time.perf_counter()
            try:
                return fn(worker)
            finally:
                self.trace.append(("unit", worker.index, time.perf_counter() - t0))

        return run

    def _run(self, tasks, order, results):
        n_threads = min(len(self.workers), len(tasks))
        if n_threads <= 1:
            for i in order:
                results[i] = tasks[i][1](self.workers[0])
            return results
        lock = threading.Lock()
        cursor = [0]
        errors = []

        def loop(worker):
            while True:
                with lock:
                    if errors or cursor[0] >= len(order):
                        return
                    i = order[cursor[0]]
                    cursor[0] += 1
                try:
                    results[i] = tasks[i][1](worker)
                except BaseException as exc:  # noqa: BLE001 - re-raised on the calling thread
                    with lock:
                        errors.append(exc)
                    return

        threads = [threading.Thread(target=loop, args=(w,), daemon=True) for w in self.workers[:n_threads]]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results

    def report(self):
        """Per phase: wall time, number of units, busy time per GPU (RESNMTF_TRACE)."""
        lines, busy = [], {}
        for kind, a, b in self.trace or []:
            if kind == "unit":
                busy[a] = busy.get(a, 0.0) + b
            else:
                per = " ".join(f"gpu{w}={busy.get(w, 0.0):.2f}" for w in sorted(busy))
                lines.append(f"  phase: {a:3d} units, {b:6.2f} s wall, busy {per}")
                busy = {}
        return "\n".join(lines)

    def close(self):
        if self.trace:
            print(self.report())
            self.trace = []
        for w in self.workers:
            w.clear()
        self.loaders.clear()
        self.host_data.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
