"""Row-sharded view (NCCL all-reduce of [X'F | F'F | colSums(F)] per sweep) against the unsharded oracle: the whole
code path on ONE GPU through a communicator of one rank, and over 2 GPUs (those tests need >= 2 visible GPUs: run with
gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from resnmtf_b200 import _lib as L  # noqa: E402

pytestmark = pytest.mark.gpu
WORLD = 2


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def problem(k):
    from resnmtf_b200 import synth

    rng = np.random.default_rng(23)
    n, p = 1000, 300  # 16 panels -> 8 + 8, ragged last panel
    x = synth.prep(synth.planted_view(n, p, 4, rng, 0.3, 0.3)[0])
    f, s, g = synth.random_factors(n, p, k, rng)
    return x, f, s, g


def worker(rank, world, port, out_dir, impl, err_mode, k):
    import torch
    import torch.distributed as dist

    from helpers import rel_err
    from oracle import resnmtf_oracle as O
    from resnmtf_b200 import sharding
    from resnmtf_b200.device import Context, DeviceFit

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = Context(rank)
    ids = [Context.comm_id_create() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.join(ids[0], rank, world)
    x, f, s, g = problem(k)
    b, e = sharding.row_shards(x.shape[0], world)[rank]
    z = np.zeros((1, 1))
    states = []
    ref = O.res_nmtf_loop([x], None, None, [None], [None], [f], [s], [g], [k], z, z, z, n_iters=6,
                          trace=lambda t, cf, cs, cg, cl, cm, er: states.append(
                              (cf[0].copy(), cs[0].copy(), cg[0].copy(), cl[0].copy(), cm[0].copy(), er.copy())))
    fit = DeviceFit(ctx, [e - b], [x.shape[1]], [k])
    fit.set_options(err_mode=err_mode, impl=impl)
    fit.set_data(0, x[b:e])
    fit.set_factors(0, f[b:e], s, g)  # lambda/mu default to the GLOBAL colSums (all-reduced)
    worst = 0.0
    for t in range(6):
        fit.step()
        fl, sl, gl, lam, mu = fit.get_factors(0)
        errs, _ = fit.view_errors()
        cf, cs, cg, cl, cm, oerr = states[t]
        for a, r in ((fl, cf[b:e]), (sl, cs), (gl, cg), (lam, cl), (mu, cm), (errs, oerr)):
            worst = max(worst, rel_err(a, r))
    fit.normalise()
    fn, sn, gn, _, _ = fit.get_factors(0)
    worst_n = max(rel_err(fn, ref["output_f"][0][b:e]), rel_err(sn, ref["output_s"][0]),
                  rel_err(gn, ref["output_g"][0]))
    fit.close()
    ctx.close()
    np.save(os.path.join(out_dir, f"worst{rank}.npy"), np.array([worst, worst_n]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("impl,err_mode,k", [(L.IMPL_AUTO, L.ERR_AUTO, 8), (L.IMPL_TMA, L.ERR_AUTO, 5),
                                             (L.IMPL_DMMA, L.ERR_DIRECT, 3), (L.IMPL_DFMA, L.ERR_DIRECT, 9),
                                             (L.IMPL_DFMA, L.ERR_ALGEBRAIC, 4)])
def test_sharded_code_path_on_a_one_rank_communicator(impl, err_mode, k):
    """The WHOLE row-sharded code path on a 1-GPU box: a context that joined a communicator of ONE rank runs
    rn_g_step (T to HBM) -> rn_pack_ff -> ncclAllReduce -> stand-alone rn_g_epilogue, the all-reduced ||X||^2, column
    sums and direct residual, exactly as an 8-rank run does -- against the oracle sweep by sweep (1e-9), and the
    fused one-pass kernel must NOT be picked (the view is sharded)."""
    from helpers import rel_err
    from oracle import resnmtf_oracle as O
    from resnmtf_b200.device import Context, DeviceFit

    x, f, s, g = problem(k)
    z = np.zeros((1, 1))
    states = []
    ref = O.res_nmtf_loop([x], None, None, [None], [None], [f], [s], [g], [k], z, z, z, n_iters=6,
                          trace=lambda t, cf, cs, cg, cl, cm, er: states.append(
                              (cf[0].copy(), cs[0].copy(), cg[0].copy(), cl[0].copy(), cm[0].copy(), er.copy())))
    with Context(0) as ctx:
        ctx.join(Context.comm_id_create(), 0, 1)
        fit = DeviceFit(ctx, [x.shape[0]], [x.shape[1]], [k])
        fit.set_options(err_mode=err_mode, impl=impl)
        fit.set_data(0, x)
        fit.set_factors(0, f, s, g)
        worst = 0.0
        for t in range(6):
            fit.step()
            fl, sl, gl, lam, mu = fit.get_factors(0)
            errs, _ = fit.view_errors()
            cf, cs, cg, cl, cm, oerr = states[t]
            for a, r in ((fl, cf), (sl, cs), (gl, cg), (lam, cl), (mu, cm), (errs, oerr)):
                worst = max(worst, rel_err(a, r))
        c = fit.counters()
        assert c["impl"] != L.IMPL_FUSED  # a sharded view never takes the one-GPU one-pass kernel
        assert c["kernel_launches"] >= 5  # F step, G stream, pack, all-reduce, G epilogue (+ residual passes)
        fit.normalise()
        fn, sn, gn, _, _ = fit.get_factors(0)
        worst_n = max(rel_err(fn, ref["output_f"][0]), rel_err(sn, ref["output_s"][0]), rel_err(gn, ref["output_g"][0]))
        fit.close()
    assert worst <= 1e-9 and worst_n <= 1e-9, (worst, worst_n)


@pytest.mark.skipif(L.device_count() < WORLD, reason="needs 2 GPUs")
@pytest.mark.parametrize("impl,err_mode,k", [(L.IMPL_TMA, L.ERR_AUTO, 5), (L.IMPL_DMMA, L.ERR_DIRECT, 3),
                                             (L.IMPL_DFMA, L.ERR_DIRECT, 9)])
def test_row_sharded_view_matches_oracle(tmp_path, impl, err_mode, k):
    import torch.multiprocessing as mp

    mp.spawn(worker, args=(WORLD, free_port(), str(tmp_path), impl, err_mode, k), nprocs=WORLD, join=True)
    for r in range(WORLD):
        worst = np.load(os.path.join(tmp_path, f"worst{r}.npy"))
        assert worst[0] <= 1e-9 and worst[1] <= 1e-9, (r, worst)


@pytest.mark.skipif(L.device_count() < WORLD, reason="needs 2 GPUs")
def test_k_sweep_placed_on_two_gpus_equals_serial_sweep():
    """apply_resnmtf's k-sweep with use_parallel on >= 2 GPUs (one host thread and one context per GPU, fits
    placed longest-first) returns exactly what the single-GPU sweep returns: same selected k, same factors."""
    from resnmtf_b200 import synth
    from resnmtf_b200.api import apply_resnmtf
    from resnmtf_b200.device import Context

    views, _ = synth.block_views(2, seed=31)
    ctx = Context(0)
    kw = dict(k_max=5, spurious=False, stability=False, ctx=ctx)
    par = apply_resnmtf(views, use_parallel=True, rng=np.random.default_rng(9), **kw)
    ser = apply_resnmtf(views, use_parallel=False, rng=np.random.default_rng(9), **kw)
    assert par["output_f"][0].shape == ser["output_f"][0].shape == (180, 3)
    for key in ("output_f", "output_s", "output_g", "row_clusters", "col_clusters"):
        for a, b in zip(par[key], ser[key]):
            assert np.array_equal(a, b), key
    assert par["bisil"] == ser["bisil"]


@pytest.mark.skipif(L.device_count() < WORLD, reason="needs 2 GPUs")
def test_stability_repeats_placed_on_two_gpus_equal_serial():
    """The n_stability resample fits of apply_resnmtf (R/stability_analysis.r:302-338) dealt to one context per
    GPU give exactly the single-GPU result (per-repeat child generators): same surviving biclusters."""
    from resnmtf_b200 import synth
    from resnmtf_b200.api import apply_resnmtf
    from resnmtf_b200.device import Context

    views, _ = synth.block_views(2, seed=33)
    ctx = Context(0)
    kw = dict(k_val=3, spurious=True, stability=True, n_stability=4, ctx=ctx, max_iters=500)
    par = apply_resnmtf(views, use_parallel=True, rng=np.random.default_rng(11), **kw)
    ser = apply_resnmtf(views, use_parallel=False, rng=np.random.default_rng(11), **kw)
    for key in ("output_f", "row_clusters", "col_clusters"):
        for a, b in zip(par[key], ser[key]):
            assert np.array_equal(a, b), key
    for v in range(2):
        assert sorted(par["row_clusters"][v].sum(axis=0).tolist()) == [60.0, 60.0, 60.0]


@pytest.mark.skipif(L.device_count() < WORLD, reason="needs 2 GPUs")
def test_resident_route_units_on_two_gpus_equal_one_gpu(monkeypatch):
    """Matrix-sized view (resident-data route): k sweep + spurious removal + stability as units on 2 GPUs (views copied
    GPU to GPU, sub-samples gathered where a unit runs) equal the 1-GPU call bit for bit."""
    from resnmtf_b200 import synth
    from resnmtf_b200.api import apply_resnmtf
    from resnmtf_b200.device import Context

    views, _ = synth.block_views(1, block=200, n_blocks=3, seed=43)
    ctx = Context(0)
    kw = dict(k_min=3, k_max=4, spurious=True, stability=True, n_stability=2, num_repeats=2, ctx=ctx, max_iters=1000)
    outs = []
    for cap in ("1", "2"):
        monkeypatch.setenv("RESNMTF_MAX_GPUS", cap)
        outs.append(apply_resnmtf(views, rng=np.random.default_rng(13), **kw))
    assert outs[0]["output_f"][0].shape == (600, 3)
    for key in ("output_f", "output_s", "output_g", "row_clusters", "col_clusters"):
        for a, b in zip(outs[0][key], outs[1][key]):
            assert np.array_equal(a, b), key
    assert outs[0]["bisil"] == outs[1]["bisil"]
    assert sorted(outs[0]["row_clusters"][0].sum(axis=0).tolist()) == [200.0, 200.0, 200.0]
