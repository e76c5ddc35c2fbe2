"""Host-side dense helpers of the library (rn_dense.h: tred2 / tql2 symmetric eigensolver, Cholesky, triangular
inverse -- the Rayleigh-Ritz step and the block orthonormalisation of the subspace iteration behind
resnmtf_data_svd_topk) against LAPACK.  Compiled here with g++; no GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = tmp_path_factory.mktemp("dense") / "libdense_shim.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(out),
                    os.path.join(ROOT, "tests", "native", "dense_shim.cpp")], check=True)
    return C.CDLL(str(out))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("n,kind", [(1, "random"), (2, "random"), (7, "random"), (64, "random"), (64, "clustered"),
                                    (64, "rank3"), (150, "random")])
def test_symmetric_eigensolver_matches_lapack(shim, n, kind):
    rng = np.random.default_rng(n)
    if kind == "random":
        a = rng.standard_normal((n, n))
        a = a + a.T
    elif kind == "clustered":  # a well separated leading value and a tight cluster, like the Gram matrix of shuffled data
        q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        a = (q * np.r_[100.0, 1.0 + 1e-9 * rng.standard_normal(n - 1)]) @ q.T
        a = 0.5 * (a + a.T)
    else:
        b = rng.standard_normal((n, 3))
        a = b @ b.T
    a = np.asfortranarray(a)
    w = np.empty(n)
    v = np.empty((n, n), order="F")
    assert shim.t_sym_eig(n, _p(a), _p(w), _p(v)) == 0
    w_ref = np.linalg.eigvalsh(a)
    scale = max(1.0, np.abs(w_ref).max())
    assert np.all(np.diff(w) >= 0)
    np.testing.assert_allclose(w, w_ref, rtol=0, atol=5e-14 * scale * max(n, 8))
    np.testing.assert_allclose(v.T @ v, np.eye(n), rtol=0, atol=5e-14 * max(n, 8))
    np.testing.assert_allclose(a @ v, v * w[None, :], rtol=0, atol=1e-13 * scale * max(n, 8))


@pytest.mark.parametrize("n", [1, 5, 64])
def test_cholesky_and_triangular_inverse(shim, n):
    rng = np.random.default_rng(100 + n)
    y = rng.standard_normal((4 * n + 3, n))
    g = np.asfortranarray(y.T @ y)
    r = np.empty((n, n), order="F")
    ri = np.empty((n, n), order="F")
    assert shim.t_cholesky_upper(n, _p(g), _p(r)) == 0
    np.testing.assert_allclose(r.T @ r, g, rtol=1e-13, atol=1e-13 * np.abs(g).max())
    assert np.allclose(np.tril(r, -1), 0.0)
    shim.t_upper_inverse(n, _p(r), _p(ri))
    np.testing.assert_allclose(r @ ri, np.eye(n), rtol=0, atol=1e-12)
    q = y @ ri  # Cholesky QR: orthonormal columns
    np.testing.assert_allclose(q.T @ q, np.eye(n), rtol=0, atol=1e-12)
    bad = np.asfortranarray(-np.eye(n))
    assert shim.t_cholesky_upper(n, _p(bad), _p(r)) == 1
