"""The native fan-out (resnmtf_pool_* / resnmtf_batch_run): units run by the library's own worker threads -- no Python
threads -- against the same fits driven step by step through DeviceFit, and against the oracle."""
import numpy as np
import pytest

from helpers import RTOL, rel_err
from oracle import resnmtf_oracle as O
from resnmtf_b200 import _lib as L
from resnmtf_b200 import synth
from resnmtf_b200.device import DeviceData, DeviceFit
from resnmtf_b200.native import NativePool

pytestmark = pytest.mark.gpu


def views2(seed=1):
    vs, _ = synth.block_views(2, block=40, n_blocks=3, seed=seed)
    return [synth.prep(x) for x in vs]


def test_explicit_init_unit_equals_the_device_fit_and_the_oracle(ctx):
    data = views2()
    k = 3
    rng = np.random.default_rng(2)
    inits = [synth.random_factors(120, 120, k, rng) for _ in range(2)]
    f0, s0, g0 = [i[0] for i in inits], [i[1] for i in inits], [i[2] for i in inits]
    phi = O.init_rest_mats(np.array([[0, 200.0], [0, 0]]), 2)
    ids = np.arange(120, dtype=np.int32)
    maps = [(L.MAP_ROW, 0, 1, ids, ids), (L.MAP_ROW, 1, 0, ids, ids)]
    with NativePool(1) as pool:
        pool.put_host(7, data)
        out = pool.run([dict(key=7, k=[k, k], init_f=f0, init_s=s0, init_g=g0, phi=phi, maps=maps, n_iters=6)])[0]
    names = [[f"r{i}" for i in range(120)]] * 2
    cn = [[f"a{i}" for i in range(120)], [f"b{i}" for i in range(120)]]
    z = np.zeros((2, 2))
    ref = O.res_nmtf_loop(data, O.shared_names(names), O.shared_names(cn), names, cn, f0, s0, g0, [k, k], phi, z, z,
                          n_iters=6)
    assert out["iters"] == 6
    assert rel_err(out["total_err"], ref["All_Error"]) <= RTOL
    for v in range(2):
        assert rel_err(out["output_f"][v], ref["output_f"][v]) <= RTOL
        assert rel_err(out["output_s"][v], ref["output_s"][v]) <= RTOL
        assert rel_err(out["output_g"][v], ref["output_g"][v]) <= RTOL
        assert rel_err(out["lambda"][v], ref["lambda"][v]) <= RTOL
        assert rel_err(out["mu"][v], ref["mu"][v]) <= RTOL


def test_svd_initialised_unit_equals_the_reference_initialisation(ctx):
    """A unit without explicit factors: SVD initialisation on the device with the caller's noise draw
    (R/update_steps.r:92-105), then the loop to convergence -- against the oracle started from LAPACK's svd."""
    data = views2(seed=3)
    k = 3
    rng = np.random.default_rng(4)
    noise = [np.abs(np.sqrt(0.05) * rng.standard_normal((k, k))) for _ in range(2)]
    with NativePool(1) as pool:
        pool.put_host(1, data)
        out = pool.run([dict(key=1, k=[k, k], noise=noise, n_iters=None, max_iters=2000)])[0]
    f0, s0, g0, _, _ = O.init_mats_inner(data, [k, k], noise)
    z = np.zeros((2, 2))
    names_r, names_c = O.default_names(data)
    ref = O.res_nmtf_loop(data, O.shared_names(names_r), O.shared_names(names_c), names_r, names_c, f0, s0, g0, [k, k],
                          z, z, z, n_iters=None)
    assert out["iters"] == len(ref["All_Error"])
    assert rel_err(out["total_err"], ref["All_Error"]) <= 1e-8
    rows_d, cols_d, _ = O.binarise(out["output_f"], out["output_g"], out["output_s"])
    rows_o, cols_o, _ = O.binarise(ref["output_f"], ref["output_g"], ref["output_s"])
    for v in range(2):
        assert rel_err(out["output_f"][v], ref["output_f"][v]) <= 1e-7
        assert np.array_equal(rows_d[v], rows_o[v]) and np.array_equal(cols_d[v], cols_o[v])


def test_shuffled_and_subsampled_units_equal_the_same_steps_by_hand(ctx):
    """derive = SUBSAMPLE | SHUFFLE inside a unit is exactly subsample() -> shuffle() -> fit on the data handles."""
    x = synth.prep(synth.planted_view(600, 260, 3, np.random.default_rng(5), 0.3, 0.3)[0])
    rng = np.random.default_rng(6)
    rows = np.sort(rng.permutation(600)[:540])
    cols = np.sort(rng.permutation(260)[:234])
    k = 4
    noise = [np.abs(np.sqrt(0.05) * rng.standard_normal((k, k)))]
    with NativePool(1) as pool:
        pool.put_host(3, [x])
        outs = pool.run([
            dict(key=3, k=[k], noise=noise, rows=[rows], cols=[cols], n_iters=15),
            dict(key=3, k=[k], noise=noise, rows=[rows], cols=[cols], shuffle_seed=99, n_iters=15),
            dict(key=3, k=[k], noise=noise, shuffle_seed=99, n_iters=15),
        ])
        c0 = pool.contexts[0]
        base = pool.get(3, 0, 0)
        sub = base.subsample(rows, cols)
        # same key derivation as the unit: view v is keyed by seed + golden * (v + 1)
        key = (99 + 0x9E3779B97F4A7C15 * 1) & (2 ** 64 - 1)
        chain = [(sub, outs[0]), (sub.shuffle(key), outs[1]), (base.shuffle(key), outs[2])]
        for dd, got in chain:
            u, d, v = dd.svd_topk(k)
            s0 = np.abs(np.diag(d)) + noise[0]
            csf, csg = u.sum(0), v.sum(0)
            s0 = s0 * (csf * csg)[None, :]
            f0, g0 = u / csf[None, :], v / csg[None, :]
            fit = DeviceFit(c0, [dd.shape[0]], [dd.shape[1]], [k])
            fit.attach_data(0, dd)
            fit.set_factors(0, f0, s0, g0, f0.sum(0), g0.sum(0))
            fit.run(15)
            errs = fit.errors()
            fit.normalise()
            f, s, g, _, _ = fit.get_factors(0)
            fit.close()
            assert got["output_f"][0].shape == f.shape
            # the unit forms the initial factors in C++ (sequential column sums), this test in NumPy (pairwise sums):
            # same numbers to rounding, so the 15 sweeps agree to well inside the parity bar rather than bit for bit
            assert rel_err(got["output_f"][0], f) <= 1e-10 and rel_err(got["output_g"][0], g) <= 1e-10
            assert rel_err(got["total_err"], errs) <= 1e-10


def test_batch_is_independent_of_the_number_of_gpus_and_of_the_unit_order(ctx):
    """A k-sweep's worth of units (fits + shuffled refits): same numbers from one worker and from all visible GPUs,
    and in any submission order."""
    x = synth.prep(synth.planted_view(800, 300, 3, np.random.default_rng(8), 0.3, 0.3)[0])
    rng = np.random.default_rng(9)
    units = []
    for k in (3, 4, 5):
        noise = [np.abs(np.sqrt(0.05) * rng.standard_normal((k, k)))]
        units.append(dict(key=0, k=[k], noise=noise, n_iters=None, max_iters=300))
        for r in range(2):
            units.append(dict(key=0, k=[k], noise=noise, shuffle_seed=1000 * k + r, n_iters=None, max_iters=300))
    results = []
    for n_gpus, order in ((1, list(range(len(units)))), (0, list(reversed(range(len(units)))))):
        with NativePool(n_gpus) as pool:
            pool.put_host(0, [x])
            out = pool.run([units[i] for i in order])
            res = [None] * len(units)
            for pos, i in enumerate(order):
                res[i] = out[pos]
            results.append(res)
            if n_gpus == 0 and len(pool) > 1:
                assert len({r["gpu"] for r in out}) > 1
    for a, b in zip(*results):
        assert a["iters"] == b["iters"]
        assert np.array_equal(a["output_f"][0], b["output_f"][0])
        assert np.array_equal(a["output_s"][0], b["output_s"][0])
        assert np.array_equal(a["total_err"], b["total_err"])


def test_a_failing_unit_reports_its_own_status(ctx):
    x = views2()[0]
    with NativePool(1) as pool:
        pool.put_host(0, [x])
        with pytest.raises(L.ResnmtfError, match="unit 1"):
            pool.run([dict(key=0, k=[3], n_iters=2), dict(key=0, k=[40], n_iters=2)])  # k > 16


def test_two_workers_with_graph_capturing_units_and_a_view_copy(ctx):
    """Two worker threads (a pool of two contexts, here on ONE GPU): units with k > 8 run the two-pass kernels, whose
    sweep is captured into a CUDA graph by the worker, while the other worker copies the view from the first context
    (resnmtf_data_copy) and runs its own units.  Regression: the copy used to synchronise the source context's stream,
    which fails while the other thread has it in capture (found on 8 GPUs).  Results equal the one-worker pool's."""
    rng = np.random.default_rng(9)
    x = synth.prep(synth.planted_view(700, 300, 3, rng, 0.3, 0.3)[0])
    ks = [9, 3, 10, 4, 11, 5, 9, 6]
    units = lambda: [dict(key=3, k=[k], noise=[np.abs(np.sqrt(0.05) * np.random.default_rng(k).standard_normal((k, k)))],  # noqa: E731
                          n_iters=None, max_iters=400) for k in ks]
    with NativePool(devices=[ctx.device, ctx.device]) as pool:
        assert len(pool) == 2
        pool.put_host(3, [x])
        two = pool.run(units())
    with NativePool(devices=[ctx.device]) as pool:
        pool.put_host(3, [x])
        one = pool.run(units())
    assert {u["gpu"] for u in two} == {0, 1}
    for a, b in zip(two, one):
        assert a["iters"] == b["iters"] and np.array_equal(a["total_err"], b["total_err"])
        assert np.array_equal(a["output_f"][0], b["output_f"][0]) and np.array_equal(a["output_g"][0], b["output_g"][0])


@pytest.mark.parametrize("overlap", ["1", "0"])
def test_home_svd_is_computed_by_the_home_worker_while_the_others_start(ctx, overlap, monkeypatch):
    """Two workers (two contexts on ONE GPU), a k-sweep's worth of units: the SVD triplets of the resident view are the
    first job of the home GPU's worker while the other worker already runs shuffled refits (its copy of the view is made
    before the triplets exist and adopts them later); RESNMTF_POOL_SVD_OVERLAP=0 computes them before the fan-out.
    Same numbers as the one-worker pool either way."""
    monkeypatch.setenv("RESNMTF_POOL_SVD_OVERLAP", overlap)
    x = synth.prep(synth.planted_view(900, 350, 3, np.random.default_rng(18), 0.3, 0.3)[0])
    rng = np.random.default_rng(19)
    units = []
    for k in (5, 4, 3):
        noise = [np.abs(np.sqrt(0.05) * rng.standard_normal((k, k)))]
        for r in range(2):
            units.append(dict(key=0, k=[k], noise=noise, shuffle_seed=77 * k + r, n_iters=None, max_iters=200))
        units.append(dict(key=0, k=[k], noise=noise, n_iters=None, max_iters=200))
    with NativePool(devices=[ctx.device, ctx.device]) as pool:
        pool.put_host(0, [x])
        two = pool.run(units)
    with NativePool(devices=[ctx.device]) as pool:
        pool.put_host(0, [x])
        one = pool.run(units)
    assert {u["gpu"] for u in two} == {0, 1}
    for a, b in zip(two, one):
        assert a["iters"] == b["iters"] and np.array_equal(a["total_err"], b["total_err"])
        assert np.array_equal(a["output_f"][0], b["output_f"][0]) and np.array_equal(a["output_s"][0], b["output_s"][0])
