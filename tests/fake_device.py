"""CPU stand-ins for the device objects, so that the HOST logic around the C ABI -- the unit decomposition of
apply_resnmtf (resnmtf_b200/fitpool.py, api.run_fits), the stability analysis over the pool, the resident-data code
paths (torch tensors, here on the CPU) -- runs in the `-m "not gpu"` suite.  Test infrastructure only: the fake fit is
the oracle's loop (oracle/resnmtf_oracle.py), the fake pair kernel is the host restatement of jsd_calc.  Nothing here
is reachable from the package."""
from __future__ import annotations

import contextlib
import ctypes as C

import numpy as np

from oracle import resnmtf_oracle as O


class FakeContext:
    def __init__(self, device=0):
        self.device = int(device)
        self.fits = 0  # how many fits ran on this context (placement checks)


def _from_address(ptr, n, p):
    """The n x p column-major matrix at host address ``ptr`` (a CPU tensor's data_ptr)."""
    buf = (C.c_double * (int(n) * int(p))).from_address(int(ptr))
    return np.ctypeslib.as_array(buf).reshape(int(p), int(n)).T.copy(order="F")


class FakeData:
    def __init__(self, ctx, x):
        self.ctx = ctx
        self.x = np.asfortranarray(x, dtype=np.float64)
        self.shape = self.x.shape

    @classmethod
    def from_device(cls, ctx, dev_ptr, n, p, ld=None):
        return cls(ctx, _from_address(dev_ptr, n, p))

    def close(self):
        pass


class FakeFit:
    """DeviceFit's interface on top of the oracle loop (no restrictions: the orchestration tests do not couple views)."""

    def __init__(self, ctx, n, p, k):
        self.ctx, self.n, self.p, self.k = ctx, list(n), list(p), list(k)
        self.n_views = len(self.n)
        self.x = [None] * self.n_views
        self.init = [None] * self.n_views
        self.out = None

    def set_options(self, err_mode=0, impl=0):
        pass

    def set_data(self, v, x):
        self.x[v] = np.asfortranarray(x, dtype=np.float64)

    def attach_data(self, v, data):
        assert data.ctx is self.ctx, "data handle of another context"
        self.x[v] = data.x

    def set_data_device(self, v, dev_ptr, ld):
        self.x[v] = _from_address(dev_ptr, self.n[v], self.p[v])

    def set_factors(self, v, f, s, g, lam=None, mu=None):
        self.init[v] = (np.array(f), np.array(s), np.array(g))

    def set_restrictions(self, phi=None, xi=None, psi=None):
        for m in (phi, xi, psi):
            assert m is None or not np.any(m), "the fake fit does not couple views"

    def set_shared_map(self, kind, v, w, idx_v, idx_w):
        pass

    def run(self, n_iters=None, tol=1.0e-6, max_iters=0):
        self.ctx.fits += 1
        V = self.n_views
        rn, cn = O.default_names(self.x)
        z = np.zeros((V, V))
        self.out = O.res_nmtf_loop(self.x, O.shared_names(rn), O.shared_names(cn), rn, cn,
                                   [i[0] for i in self.init], [i[1] for i in self.init], [i[2] for i in self.init],
                                   self.k, z, z, z, n_iters=n_iters, max_iters=max_iters or None)
        self.normalised = False
        return len(self.out["All_Error"])

    def errors(self):
        return np.asarray(self.out["All_Error"])

    def normalise(self):
        self.normalised = True

    def get_factors(self, v):
        o = self.out
        if self.normalised:
            return o["output_f"][v], o["output_s"][v], o["output_g"][v], o["lambda"][v], o["mu"][v]
        return o["raw_f"][v], o["raw_s"][v], o["raw_g"][v], o["lambda"][v], o["mu"][v]

    def counters(self):
        return {"iterations": len(self.out["All_Error"])}

    def close(self):
        pass


def fake_jsd_pairs(ctx, vecs, bw, vmax, pair_a, pair_b):
    from resnmtf_b200.bicluster import jsd_calc

    vecs = np.asarray(vecs)
    return np.array([jsd_calc(vecs[:, a], vecs[:, b]) for a, b in zip(pair_a, pair_b)])


class _NoStream:
    def synchronize(self):
        pass


def install(monkeypatch, n_devices=1, resident=False):
    """Patches the package to run on the fakes; returns the contexts.  ``resident=True`` also enables the
    resident-data route with CPU tensors standing in for device tensors."""
    import torch

    from resnmtf_b200 import api, device, fitpool

    contexts = [FakeContext(d) for d in range(n_devices)]
    monkeypatch.setattr(api, "DeviceFit", FakeFit)
    monkeypatch.setattr(fitpool, "DeviceData", FakeData)
    monkeypatch.setattr(api, "default_context", lambda: contexts[0])
    monkeypatch.setattr(api, "device_contexts", lambda first: contexts)
    monkeypatch.setattr(device, "default_context", lambda: contexts[0])
    monkeypatch.setattr(device, "device_contexts", lambda first: contexts)
    monkeypatch.setattr(device, "jsd_pairs", fake_jsd_pairs)
    monkeypatch.setattr(api, "_torch_cuda", (lambda: torch) if resident else (lambda: None))
    monkeypatch.setattr(api, "torch_device", lambda index: torch.device("cpu"))
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: _NoStream())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda dev=None: None)
    return contexts
