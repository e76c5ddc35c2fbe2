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

    def place_resident(self, key, data, prep=False):
        """Uploads the host views once (to the first GPU) and copies them from there to the other GPUs.  ``prep``: the
        views are raw; make_non_neg_inner and matrix_normalisation (R/utils.r:20-27, 86-88) run on the first GPU
        before the copies (same warning as the host version)."""
        import warnings

        import torch

        from .api import torch_device

        self.host_data[key] = data
        w0 = self.workers[0]
        dev0 = torch_device(w0.ctx.device)
        first = []
        for m in data:
            xt = torch.from_numpy(np.ascontiguousarray(np.asarray(m.x, dtype=np.float64).T)).to(dev0)  # p x n
            if prep:
                col_min = xt.min(dim=1, keepdim=True).values
                if bool((col_min < 0).any()):
                    warnings.warn("Matrix is not non-negative. Has been made non-negative.")
                xt = xt + torch.abs(torch.clamp(col_min, max=0.0))
                xt = xt / xt.sum(dim=1, keepdim=True)
            first.append(xt)
        w0.views[key] = first
        torch.cuda.synchronize(dev0)

        def copy_from_first(worker):  # runs on the worker's own thread, so the GPU-to-GPU copies overlap
            out = [t.to(torch_device(worker.ctx.device)) for t in first]
            torch.cuda.synchronize(out[0].device)
            return out

        self.loaders[key] = copy_from_first

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
        """``tasks``: [(cost, callable(worker) -> result)] or [(cost, callable, deps, label[, affinity])] where ``deps`` lists the
        indices of the tasks whose results must exist first (a task with deps reads them from the list this method
        returns -- it is handed over as ``worker.results`` -- e.g. the post-processing of a fit waits for its core
        and its shuffled refits).  Returns the results in task order.  A free GPU takes the ready task with the
        largest cost; with one GPU the tasks simply run one after the other on the calling thread."""
        tasks = [(t[0], t[1], tuple(t[2]) if len(t) > 2 else (), t[3] if len(t) > 3 else "unit",
                  t[4] if len(t) > 4 else None) for t in tasks]
        order = sorted(range(len(tasks)), key=lambda i: (-float(tasks[i][0]), i))
        results = [None] * len(tasks)
        finished = [False] * len(tasks)
        t_phase = time.perf_counter()

        def execute(i, worker):
            worker.results = results
            t0 = time.perf_counter()
            try:
                results[i] = tasks[i][1](worker)
            finally:
                if self.trace is not None:
                    self.trace.append(("unit", worker.index, time.perf_counter() - t0, tasks[i][3]))

        owner = {}  # affinity key -> index of the worker that took the first task of that key

        def next_ready(pending, worker):
            """Highest-priority ready task for this worker.  Tasks that share an affinity key (the fits of one data
            set: whoever runs the first of them builds the library-layout copy of the views and the SVD triplets on
            its GPU) go to the worker that owns the key as long as it is busy with them; another worker takes one
            only when nothing else is ready for it."""
            fallback = None
            for pos, i in enumerate(pending):
                if not all(finished[d] for d in tasks[i][2]):
                    continue
                key = tasks[i][4]
                if key is None or owner.get(key, worker.index) == worker.index:
                    if key is not None:
                        owner[key] = worker.index
                    return pending.pop(pos)
                if fallback is None:
                    fallback = pos
            if fallback is not None:
                return pending.pop(fallback)
            return None

        pending = list(order)
        n_threads = min(len(self.workers), len(tasks))
        try:
            if n_threads <= 1:
                while pending:
                    i = next_ready(pending, self.workers[0])
                    if i is None:
                        raise RuntimeError("FitPool.run: circular task dependencies")
                    execute(i, self.workers[0])
                    finished[i] = True
                return results
            cond = threading.Condition()
            errors = []
            running = [0]

            def loop(worker):
                while True:
                    with cond:
                        while True:
                            if errors or not pending:
                                return
                            i = next_ready(pending, worker)
                            if i is not None:
                                running[0] += 1
                                break
                            if running[0] == 0:
                                errors.append(RuntimeError("FitPool.run: circular task dependencies"))
                                cond.notify_all()
                                return
                            cond.wait()
                    try:
                        execute(i, worker)
                    except BaseException as exc:  # noqa: BLE001 - re-raised on the calling thread
                        with cond:
                            errors.append(exc)
                            running[0] -= 1
                            cond.notify_all()
                        return
                    with cond:
                        finished[i] = True
                        running[0] -= 1
                        cond.notify_all()

            threads = [threading.Thread(target=loop, args=(w,), daemon=True) for w in self.workers[:n_threads]]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            if errors:
                raise errors[0]
            return results
        finally:
            if self.trace is not None:
                self.trace.append(("phase", len(tasks), time.perf_counter() - t_phase, ""))

    def report(self):
        """Per phase: wall time, number of units, busy time per GPU, mean unit time per kind (RESNMTF_TRACE)."""
        lines, busy, kinds = [], {}, {}
        for kind, a, b, label in self.trace or []:
            if kind == "unit":
                busy[a] = busy.get(a, 0.0) + b
                kinds.setdefault(label, []).append(b)
            else:
                per = " ".join(f"gpu{w}={busy.get(w, 0.0):.2f}" for w in sorted(busy))
                mean = " ".join(f"{lb}:{len(v)}x{sum(v) / len(v):.3f}s" for lb, v in sorted(kinds.items()))
                lines.append(f"  phase: {a:3d} units, {b:6.2f} s wall, busy {per} | {mean}")
                busy, kinds = {}, {}
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
