"""Host logic of the multi-GPU placement of apply_resnmtf (independent fits): the unit decomposition, the
pool, the stability analysis over the pool and the resident-data route, run on CPU stand-ins for the device objects
(tests/fake_device.py: the oracle's loop behind DeviceFit's interface).  What must hold: the result of a call does
not depend on the number of GPUs, and the reference's own assertions (test-resnmtf.R:63-135) still hold."""
import threading

import numpy as np
import pytest

import fake_device
from resnmtf_b200 import synth
from resnmtf_b200.api import apply_resnmtf, res_nmtf_inner
from resnmtf_b200.fitpool import FitPool


def _same(a, b, keys=("output_f", "output_s", "output_g", "row_clusters", "col_clusters")):
    for key in keys:
        for x, y in zip(a[key], b[key]):
            assert np.array_equal(x, y), key


def test_pool_runs_every_unit_once_longest_first_and_keeps_task_order():
    pool = FitPool([fake_device.FakeContext(d) for d in range(3)])
    seen, lock = [], threading.Lock()

    def unit(i):
        def run(worker):
            with lock:
                seen.append((i, worker.index))
            return i * i

        return run

    costs = [1.0, 5.0, 3.0, 5.0, 0.5, 2.0, 4.0]
    out = pool.run([(c, unit(i)) for i, c in enumerate(costs)])
    assert out == [i * i for i in range(len(costs))]
    assert sorted(i for i, _ in seen) == list(range(len(costs)))
    one = FitPool([fake_device.FakeContext(0)])
    order = []
    one.run([(c, (lambda w, i=i: order.append(i))) for i, c in enumerate(costs)])
    assert order == [1, 3, 6, 2, 5, 0, 4]  # longest first, ties in task order


def test_pool_reraises_a_unit_failure():
    pool = FitPool([fake_device.FakeContext(d) for d in range(2)])

    def bad(worker):
        raise ValueError("unit failed")

    with pytest.raises(ValueError, match="unit failed"):
        pool.run([(1.0, lambda w: 1), (2.0, bad), (1.0, lambda w: 2)])


@pytest.mark.parametrize("resident", [False, True])
def test_k_sweep_with_spurious_removal_is_independent_of_the_gpu_count(monkeypatch, resident):
    """k sweep + shuffled refits (test-resnmtf.R:123-135 with the spurious-bicluster test switched on): 1 vs 3 GPUs.
    ``resident``: 420 x 600 views (>= 250k entries) take the resident-data route -- initialisation from the resident
    tensor, device shuffles, one JSD batch per view -- with CPU tensors standing in."""
    block = 200 if resident else 60
    views, _ = synth.block_views(1, block=block, n_blocks=3, seed=5)
    if resident:
        views = [np.asfortranarray(views[0][:420, :])]
    kw = dict(k_min=3, k_max=4, spurious=True, stability=False, num_repeats=2, max_iters=60)
    outs = []
    for n_dev in (1, 3):
        with monkeypatch.context() as mp:
            ctxs = fake_device.install(mp, n_dev, resident=resident)
            outs.append(apply_resnmtf(views, rng=np.random.default_rng(3), **kw))
            assert sum(c.fits for c in ctxs) == 2 * (1 + 2)  # 2 k values x (fit + 2 shuffled refits)
            if n_dev == 3:
                assert all(c.fits > 0 for c in ctxs)
    _same(outs[0], outs[1])
    assert outs[0]["bisil"] == outs[1]["bisil"]
    assert outs[0]["output_f"][0].shape[1] == 3  # the sweep selects the planted k
    n_rows = views[0].shape[0]
    sizes = sorted(outs[0]["row_clusters"][0].sum(axis=0).tolist())
    assert sizes == sorted([float(n_rows - 2 * block), float(block), float(block)])
    assert sorted(outs[0]["col_clusters"][0].sum(axis=0).tolist()) == [float(block)] * 3


@pytest.mark.parametrize("resident", [False, True])
def test_stability_analysis_is_independent_of_the_gpu_count(monkeypatch, resident):
    """Fixed k with spurious removal and stability (test-resnmtf.R:63-118): the resample fits and their shuffled
    refits are units like any other; 1 vs 2 GPUs give the same surviving biclusters."""
    block = 180 if resident else 60
    views, _ = synth.block_views(2, block=block, n_blocks=3, seed=7)
    kw = dict(k_val=3, spurious=True, stability=True, n_stability=2, num_repeats=2, max_iters=50)
    outs = []
    for n_dev in (1, 2):
        with monkeypatch.context() as mp:
            ctxs = fake_device.install(mp, n_dev, resident=resident)
            outs.append(apply_resnmtf(views, rng=np.random.default_rng(11), **kw))
            assert sum(c.fits for c in ctxs) == (1 + 2) + 2 * (1 + 2)
    _same(outs[0], outs[1])
    for v in range(2):
        assert sorted(outs[0]["row_clusters"][v].sum(axis=0).tolist()) == [float(block)] * 3
        assert sorted(outs[0]["col_clusters"][v].sum(axis=0).tolist()) == [float(block)] * 3


def test_res_nmtf_inner_serial_route_equals_the_pool_route(monkeypatch):
    """res_nmtf_inner on its own (the reference's serial order: fit, then the shuffled refits inside
    obtain_biclusters) returns what the same fit returns as units of a 2-GPU pool."""
    views, _ = synth.block_views(1, block=60, n_blocks=3, seed=9)
    data = [synth.prep(views[0])]
    from resnmtf_b200 import prep

    named = prep.give_names(data, 1)
    idx = prep.reorder_data(named["data"], 1, named["row_names"], named["col_names"])
    outs = []
    for n_dev in (1, 2):
        with monkeypatch.context() as mp:
            fake_device.install(mp, n_dev)
            if n_dev == 1:
                outs.append(res_nmtf_inner(named["data"], idx["row_indices"], idx["col_indices"], k_vec=[3],
                                           num_repeats=2, rng=np.random.default_rng(21), max_iters=60))
            else:
                from resnmtf_b200 import api

                with FitPool(api.device_contexts(None)) as pool:
                    pool.place_host("data", named["data"])
                    spec = dict(key="data", data=named["data"], row_indices=idx["row_indices"],
                                col_indices=idx["col_indices"], k_vec=[3], rng=np.random.default_rng(21))
                    outs.append(api.run_fits(pool, [spec], None, None, None, None, 2, True, "euclidean", False,
                                             max_iters=60)[0])
    _same(outs[0], outs[1])
    assert outs[0]["bisil"] == outs[1]["bisil"]


@pytest.mark.parametrize("kind", ["planted", "shuffled", "blocks", "rank deficient"])
def test_filtered_subspace_iteration_matches_the_dense_eigensolver(kind):
    """api._topk_eig_filtered (the top-16 eigenpairs of a fit's Gram matrix) against numpy's dense
    solver on the spectra the fits meet: planted blocks over a noise bulk, a shuffled view (one dominant pair, then a
    nearly degenerate bulk edge), the reference's block test data (three nearly equal large pairs), and a matrix of
    rank 5 (order 1050: five pairs, then the null space)."""
    import torch

    from resnmtf_b200 import api

    rng = np.random.default_rng(31)
    if kind in ("planted", "shuffled"):
        x = synth.prep(synth.planted_view(5000, 1100, 5, rng, row_prob=0.2, col_prob=0.2, height=5.0, sigma=1.0)[0])
        if kind == "shuffled":
            x = rng.permutation(x.ravel()).reshape(x.shape)
            x = x / x.sum(axis=0)[None, :]
    elif kind == "blocks":
        x = synth.prep(synth.block_views(1, block=350, n_blocks=3, seed=4)[0][0])
    else:
        x = rng.random((1500, 5)) @ rng.random((5, 1050))
    gram = torch.from_numpy(np.ascontiguousarray(x.T @ x))
    w, v = np.linalg.eigh(gram.numpy())
    w, v = w[::-1][:16], v[:, ::-1][:, :16]
    out = api._topk_eig_filtered(torch, gram, 16)
    if kind == "rank deficient":  # the five non-zero pairs are found; the rest of the block lies in the null space
        wt, vt = api._topk_eigh(torch, gram, 16)
        assert np.allclose(wt.numpy()[:5], w[:5], rtol=1e-12)
        assert np.max(np.abs(wt.numpy()[5:])) <= 1e-12 * w[0]
        assert np.max(np.abs(np.abs(vt.numpy()[:, :5]) - np.abs(v[:, :5]))) <= 1e-10
        return
    assert out is not None
    wi, vi = out[0].numpy(), out[1].numpy()
    assert np.max(np.abs(wi - w) / w[0]) <= 1e-14
    # residual at the rounding level of the matrix norm, like the dense solver's; orthonormal to rounding
    assert np.max(np.linalg.norm(gram.numpy() @ vi - vi * wi[None, :], axis=0)) <= 5e-15 * w[0]
    assert np.max(np.abs(vi.T @ vi - np.eye(16))) <= 1e-13
    # the vectors themselves agree as far as their conditioning (rounding level of the norm over the gap) allows
    gaps = np.minimum(np.abs(np.diff(w, prepend=np.inf)), np.abs(np.diff(np.append(w, w[-1] - (w[-2] - w[-1])))))
    bound = 200 * np.finfo(float).eps * w[0] / gaps
    assert np.all(np.max(np.abs(np.abs(vi) - np.abs(v)), axis=0) <= np.maximum(bound, 1e-12))


def test_pool_keeps_tasks_of_one_affinity_key_on_one_gpu_and_lets_idle_gpus_steal():
    """The fits of one data set share what the first of them builds on its GPU (library-layout copy, SVD triplets):
    they stay on the GPU that took the first one while the others have other work; a GPU with nothing else to do
    takes one rather than idle."""
    import time

    ran, lock = [], threading.Lock()

    def unit(name, seconds):
        def run(worker):
            time.sleep(seconds)
            with lock:
                ran.append((name, worker.index))

        return run

    pool = FitPool([fake_device.FakeContext(d) for d in range(3)])
    tasks = [(2.0, unit(f"fit{i}", 0.03), (), "fit", "data") for i in range(4)]
    tasks += [(1.0, unit(f"refit{i}", 0.03), (), "shuffled refit") for i in range(24)]
    pool.run(tasks)
    assert len(ran) == 28
    assert len({w for name, w in ran if name.startswith("fit")}) == 1
    # nothing but affinity tasks: the other GPUs take them instead of waiting for the owner
    ran.clear()
    pool.run([(1.0, unit(f"fit{i}", 0.05), (), "fit", "data") for i in range(6)])
    assert len({w for _, w in ran}) == 3


def test_fixed_k_with_no_clusts_and_default_stability_returns_the_factors(monkeypatch, capsys):
    """apply_resnmtf(k_val = 3, no_clusts = TRUE) with the default stability = TRUE: the no_clusts result list holds
    only the three factor lists (R/main.r:115-120), number_biclusters() of a list without row_clusters is 0
    (R/stability_analysis.r:94-99) and stability_check prints "No biclusters detected!" and hands the factors back
    (:308-312) -- it must not fail on the missing key."""
    views, _ = synth.block_views(1, block=60, n_blocks=3, seed=13)
    with monkeypatch.context() as mp:
        fake_device.install(mp, 1)
        out = apply_resnmtf(views, k_val=3, no_clusts=True, rng=np.random.default_rng(4), max_iters=40)
    assert sorted(out.keys()) == ["output_f", "output_g", "output_s"]
    assert out["output_f"][0].shape == (180, 3)
    assert "No biclusters detected!" in capsys.readouterr().out
