"""The reference's integration tests (tests/testthat/test-resnmtf.R:53-184) re-expressed against the
Python mirror of apply_resnmtf / res_nmtf_inner running on the device."""
import warnings

import numpy as np
import pytest

from helpers import rel_err
from oracle import resnmtf_oracle as O
from resnmtf_b200 import synth
from resnmtf_b200.api import apply_resnmtf, res_nmtf_inner
from resnmtf_b200.prep import NamedMatrix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def block_data():
    views, rc = synth.block_views(2, seed=21)
    return views, rc


def sizes_ok(results):
    for v in range(2):
        assert sorted(results["row_clusters"][v].sum(0)) == [60, 60, 60]
        assert sorted(results["col_clusters"][v].sum(0)) == [60, 60, 60]


def test_negative_matrix_warns(ctx, block_data):
    """test-resnmtf.R:53-58."""
    with pytest.warns(UserWarning, match="Matrix is not non-negative. Has been made non-negative."):
        apply_resnmtf([-block_data[0][0]], k_max=4, spurious=False, stability=False,
                      rng=np.random.default_rng(0), ctx=ctx)


def test_fixed_k_no_stability_no_spurious(ctx, block_data):
    """test-resnmtf.R:98-118."""
    res = apply_resnmtf(block_data[0], k_val=3, spurious=False, stability=False,
                        rng=np.random.default_rng(1), ctx=ctx)
    np.testing.assert_allclose(res["output_f"][0].sum(0), np.ones(3), atol=1e-12)
    np.testing.assert_allclose(res["output_g"][0].sum(0), np.ones(3), atol=1e-12)
    recon = res["output_f"][0] @ res["output_s"][0] @ res["output_g"][0].T
    assert np.mean(recon.sum(0) - 1.0) < 1e-3
    assert len(res["output_f"]) == 2 and res["output_f"][0].shape == (180, 3)
    sizes_ok(res)
    assert set(res) >= {"output_f", "output_s", "output_g", "Error", "All_Error", "bisil", "row_clusters",
                        "col_clusters", "lambda", "mu"}


def test_fixed_k_with_spurious_removal(ctx, block_data):
    """test-resnmtf.R:63-72."""
    res = apply_resnmtf(block_data[0], k_val=3, stability=False, rng=np.random.default_rng(2), ctx=ctx)
    sizes_ok(res)


def test_fixed_k_with_stability_no_spurious(ctx, block_data):
    """test-resnmtf.R:86-96."""
    res = apply_resnmtf(block_data[0], k_val=3, spurious=False, rng=np.random.default_rng(3), ctx=ctx)
    sizes_ok(res)


def test_k_sweep_selects_three(ctx, block_data):
    """test-resnmtf.R:123-135: the only test that pins the bisilhouette-selected k."""
    res = apply_resnmtf(block_data[0], k_max=5, spurious=False, stability=False,
                        rng=np.random.default_rng(4), ctx=ctx)
    assert res["output_f"][0].shape == (180, 3)
    sizes_ok(res)


def test_restrictions_partial_overlap(ctx, block_data):
    """test-resnmtf.R:140-184."""
    rn = [[f"row_{i}" for i in range(1, 181)],
          [f"row_{i}" for i in range(1, 121)] + [f"row_{i}" for i in range(181, 241)]]
    cn = [[f"col_{i}" for i in range(1, 181)],
          [f"col_{i}" for i in range(1, 121)] + [f"col_{i}" for i in range(181, 241)]]
    data = [NamedMatrix(x, rn[v], cn[v]) for v, x in enumerate(block_data[0])]
    rest = np.zeros((2, 2))
    rest[0, 1] = 1000.0
    res = apply_resnmtf(data, k_val=3, phi=rest, psi=rest, spurious=False, stability=False,
                        rng=np.random.default_rng(5), ctx=ctx)
    f, g = res["output_f"], res["output_g"]
    assert np.mean(np.abs(f[0][120:180] - f[1][120:180])) > np.mean(np.abs(f[0][:120] - f[1][:120]))
    assert np.mean(np.abs(g[0][120:180] - g[1][120:180])) > np.mean(np.abs(g[0][:120] - g[1][:120]))
    sizes_ok(res)


def test_res_nmtf_inner_matches_oracle_end_to_end(ctx, block_data):
    """Explicit inits through res_nmtf_inner (the route the reference supports, R/update_steps.r:49-60):
    identical binary bicluster matrices and All_Error within 1e-9."""
    data = [synth.prep(x) for x in block_data[0]]
    rng = np.random.default_rng(6)
    k = 3
    noise = [np.abs(np.sqrt(0.05) * rng.standard_normal((k, k))) for _ in range(2)]
    f0, s0, g0, _, _ = O.init_mats_inner(data, [k, k], noise)
    rn, cn = O.default_names(data)
    ri, ci = O.shared_names(rn), O.shared_names(cn)
    z = np.zeros((2, 2))
    ref = O.res_nmtf_loop(data, ri, ci, rn, cn, f0, s0, g0, [k, k], z, z, z)
    named = [NamedMatrix(x, rn[v], cn[v]) for v, x in enumerate(data)]
    res = res_nmtf_inner(named, ri, ci, f0, s0, g0, [k, k], z, z, z, spurious=False, ctx=ctx)
    assert len(res["All_Error"]) == len(ref["All_Error"])
    np.testing.assert_allclose(res["All_Error"], ref["All_Error"], rtol=1e-9)
    rows_o, cols_o, _ = O.binarise(ref["output_f"], ref["output_g"], ref["output_s"])
    for v in range(2):
        assert np.array_equal(res["row_clusters"][v], rows_o[v])
        assert np.array_equal(res["col_clusters"][v], cols_o[v])
    assert np.isclose(res["Error"], ref["Error"], rtol=1e-9)


def test_device_svd_initialisation_matches_lapack():
    """The GPU route of the SVD initialisation (Gram matrix of the smaller side + eigendecomposition)
    against the full LAPACK svd the reference calls (R/update_steps.r:92), on a planted and on a shuffled (nearly
    degenerate) spectrum, both orientations."""
    from resnmtf_b200 import api

    rng = np.random.default_rng(17)
    x = synth.prep(synth.planted_view(1300, 700, 4, rng, row_prob=0.2, col_prob=0.2, height=5.0, sigma=1.0)[0])
    shuffled = rng.permutation(x.ravel()).reshape(x.shape)
    for m in (x, shuffled, np.ascontiguousarray(x.T)):
        out = api._svd_topk_device(np.asfortranarray(m), 6, -1)
        assert out is not None, "torch CUDA path unavailable on the GPU box"
        u, d, vt = np.linalg.svd(m, full_matrices=False)
        assert rel_err(out[1], d[:6]) <= 1e-11
        assert np.max(np.abs(out[0] - np.abs(u[:, :6]))) <= 1e-10
        assert np.max(np.abs(out[2] - np.abs(vt[:6].T))) <= 1e-10


def test_device_shuffle_refits_path():
    """Views of >= 250k entries take the device-resident shuffle refits (permutation, re-normalisation, SVD
    initialisation and fit on the GPU).  The reference's own assertions still hold (three planted
    blocks of 200 rows / columns survive the spurious-bicluster removal: test-resnmtf.R:63-118 scaled up), the run is
    reproducible from the seed, and the shuffled factors have the reference's shape and normalisation."""
    from resnmtf_b200 import api
    from resnmtf_b200.device import default_context

    views, _ = synth.block_views(2, block=200, n_blocks=3, seed=4)
    outs = []
    for _ in range(2):
        res = apply_resnmtf(views, k_val=3, stability=False, spurious=True, rng=np.random.default_rng(12), max_iters=500)
        outs.append(res)
        for v in range(2):
            assert sorted(res["row_clusters"][v].sum(axis=0).tolist()) == [200.0, 200.0, 200.0]
            assert sorted(res["col_clusters"][v].sum(axis=0).tolist()) == [200.0, 200.0, 200.0]
    for v in range(2):
        assert np.array_equal(outs[0]["output_f"][v], outs[1]["output_f"][v])
        assert np.array_equal(outs[0]["row_clusters"][v], outs[1]["row_clusters"][v])
    prepped = [synth.prep(x) for x in views]
    f_mess = api.shuffled_fits_device(prepped, 3, 2, np.random.default_rng(1), default_context(), max_iters=500)
    assert f_mess is not None and len(f_mess) == 2 and len(f_mess[0]) == 2
    for rep in f_mess:
        for f in rep:
            assert f.shape == (600, 3) and np.all(f >= 0) and np.allclose(f.sum(axis=0), 1.0, atol=1e-12)


@pytest.mark.parametrize("method", ["euclidean", "manhattan", "cosine"])
def test_device_bisilhouette_matches_host(method):
    """The GPU bisilhouette against the host restatement of the same definition: overlapping and
    empty clusters, a cluster with a single row, all three distances."""
    from resnmtf_b200 import bicluster as B

    rng = np.random.default_rng(23)
    x = synth.prep(synth.planted_view(400, 150, 3, rng, row_prob=0.3, col_prob=0.3)[0])
    rc = (rng.random((400, 5)) < 0.25).astype(float)
    cc = (rng.random((150, 5)) < 0.3).astype(float)
    rc[:, 3] = 0.0            # empty row cluster
    rc[:, 4] = 0.0
    rc[17, 4] = 1.0           # single-row cluster
    host = B.bisilhouette(x, rc, cc, method=method)
    dev = B.bisilhouette_device(x, rc, cc, method=method)
    assert dev is not None
    assert len(dev["vals"]) == len(host["vals"])
    assert np.allclose(dev["vals"], host["vals"], rtol=1e-10, atol=1e-12)
    assert abs(dev["bisil"] - host["bisil"]) <= 1e-12


def test_resident_route_with_restrictions_spurious_and_stability():
    """Two matrix-sized views (600 x 600, the resident-data route) with partially shared rows and columns, phi and psi
    (test-resnmtf.R:140-184 scaled up), spurious-bicluster removal and stability analysis: the sub-samples are
    gathered on the device, their shared-name maps rebuilt from the sub-sampled names, the JSD thresholds come from
    the pair kernel.  The planted blocks survive, shared rows end up closer than unshared ones, and the call is
    reproducible from its seed."""
    views, _ = synth.block_views(2, block=200, n_blocks=3, seed=41)
    rn = [[f"row_{i}" for i in range(1, 601)],
          [f"row_{i}" for i in range(1, 401)] + [f"row_{i}" for i in range(601, 801)]]
    cn = [[f"col_{i}" for i in range(1, 601)],
          [f"col_{i}" for i in range(1, 401)] + [f"col_{i}" for i in range(601, 801)]]
    data = [NamedMatrix(x, rn[v], cn[v]) for v, x in enumerate(views)]
    rest = np.zeros((2, 2))
    rest[0, 1] = 1000.0
    outs = []
    for _ in range(2):
        res = apply_resnmtf(data, k_val=3, phi=rest, psi=rest, spurious=True, stability=True, n_stability=3,
                            num_repeats=3, rng=np.random.default_rng(6), max_iters=1000)
        outs.append(res)
        for v in range(2):
            assert sorted(res["row_clusters"][v].sum(0)) == [200, 200, 200]
            assert sorted(res["col_clusters"][v].sum(0)) == [200, 200, 200]
        f = res["output_f"]
        assert np.mean(np.abs(f[0][400:600] - f[1][400:600])) > np.mean(np.abs(f[0][:400] - f[1][:400]))
    for key in ("output_f", "output_g", "row_clusters", "col_clusters"):
        for a, b in zip(outs[0][key], outs[1][key]):
            assert np.array_equal(a, b), key


def test_device_svd_initialisation_filtered_route_matches_lapack():
    """Gram matrices of order >= 1024 take the Chebyshev-filtered subspace iteration (api._topk_eig_filtered) instead
    of the dense eigensolver.  Against LAPACK's full svd of the view itself (what the reference calls,
    R/update_steps.r:92): singular values to 1e-11, |U_k| and |V_k| to 1e-9 -- the Gram route squares the condition of
    the bulk vectors whichever solver reads the Gram matrix -- on a planted and on a shuffled spectrum."""
    from resnmtf_b200 import api

    rng = np.random.default_rng(19)
    x = synth.prep(synth.planted_view(2600, 1100, 4, rng, row_prob=0.2, col_prob=0.2, height=5.0, sigma=1.0)[0])
    shuffled = rng.permutation(x.ravel()).reshape(x.shape)
    shuffled = shuffled / shuffled.sum(axis=0)[None, :]
    for m in (x, shuffled):
        out = api._svd_topk_device(np.asfortranarray(m), 6, -1)
        assert out is not None
        u, d, vt = np.linalg.svd(m, full_matrices=False)
        assert rel_err(out[1], d[:6]) <= 1e-11
        assert np.max(np.abs(out[0] - np.abs(u[:, :6]))) <= 1e-9
        assert np.max(np.abs(out[2] - np.abs(vt[:6].T))) <= 1e-9


def test_matrix_sized_route_runs_without_torch():
    """The default route of apply_resnmtf on matrix-sized views -- prep, SVD initialisation, shuffles, sub-samples, the
    fits, JSD thresholds, bisilhouette -- goes through the C ABI only: a fresh interpreter finishes the reference's
    default call (k sweep + spurious-bicluster removal + stability analysis) without ever importing torch."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {root!r})\n"
        "from resnmtf_b200 import synth\n"
        "from resnmtf_b200.api import apply_resnmtf\n"
        "views, _ = synth.block_views(1, block=200, n_blocks=3, seed=5)\n"
        "res = apply_resnmtf(views, k_min=3, k_max=4, num_repeats=3, n_stability=3, rng=np.random.default_rng(9),\n"
        "                    max_iters=600)\n"
        "assert 'torch' not in sys.modules, 'torch was imported on the default route'\n"
        "sizes = res['row_clusters'][0].sum(0)\n"
        "assert set(sizes) <= {0.0, 200.0} and (sizes == 200).sum() >= 3, sizes  # planted blocks, whatever k was selected\n"
        "print('selected k', res['output_f'][0].shape[1], 'bisil', res['bisil'])\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "selected k" in out.stdout


def test_native_route_is_reproducible_and_matches_the_first_version_statistically(monkeypatch):
    """Same seed, same result (every unit has its own generator and the library's sums have fixed orders); the first
    version of the matrix-sized steps (RESNMTF_ROUTE=torch: other permutation generator, library eigensolver) finds the
    same biclusters on planted data."""
    views, _ = synth.block_views(1, block=200, n_blocks=3, seed=6)
    a = apply_resnmtf(views, k_val=3, num_repeats=3, n_stability=3, rng=np.random.default_rng(3), max_iters=800)
    b = apply_resnmtf(views, k_val=3, num_repeats=3, n_stability=3, rng=np.random.default_rng(3), max_iters=800)
    for key in ("output_f", "output_g", "output_s", "row_clusters", "col_clusters"):
        assert np.array_equal(a[key][0], b[key][0])
    assert a["bisil"] == b["bisil"]
    monkeypatch.setenv("RESNMTF_ROUTE", "torch")
    c = apply_resnmtf(views, k_val=3, num_repeats=3, n_stability=3, rng=np.random.default_rng(3), max_iters=800)
    assert sorted(c["row_clusters"][0].sum(0)) == sorted(a["row_clusters"][0].sum(0)) == [200, 200, 200]
    assert abs(c["bisil"] - a["bisil"]) < 1e-6
