"""world_size-2 gloo test (CPU) of the multi-GPU host logic:
  * row-sharded view: each rank holds a panel range of X and the matching rows of F; one all-reduce of
    [X'F | F'F | colSums(F) | ||X||^2 | residual] per sweep; G, S, lambda, mu replicated.  The result must
    equal the unsharded oracle (this is the algebra the NCCL path of DESIGN.md section 6 implements);
  * independent fits placed on ranks with assign_fits, results gathered on rank 0."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import resnmtf_oracle as O  # noqa: E402
from resnmtf_b200 import sharding, synth  # noqa: E402

WORLD = 2


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def problem():
    rng = np.random.default_rng(17)
    n, p, k = 333, 70, 4  # 6 panels -> 3 + 3; the last panel is ragged
    x = synth.prep(synth.planted_view(n, p, 3, rng, 0.3, 0.3)[0])
    f, s, g = synth.random_factors(n, p, k, rng)
    return x, f, s, g


def sharded_sweeps(rank, world, x, f, s, g, n_iters):
    """The row-sharded sweep with the reference's update rules (uncoupled single view)."""
    b, e = sharding.row_shards(x.shape[0], world)[rank]
    xs, fs = x[b:e], f[b:e].copy()
    lam, mu = f.sum(0), g.sum(0)  # replicated (global colSums of the inits)
    k = s.shape[0]
    xn = torch.tensor([np.sum(xs * xs)])
    dist.all_reduce(xn)
    errs = []
    for _ in range(n_iters):
        # F step: row-local
        num = (xs @ g) @ s.T
        den = (fs @ s) @ ((g.T @ g) @ s.T)
        ratio = num / (den + 0.5 * lam[None, :])
        ratio[np.isnan(ratio)] = 1.0
        fs = np.abs(fs * ratio)
        # the one exchange step: [X'F | F'F | colSums(F)]
        buf = torch.from_numpy(np.concatenate([(xs.T @ fs).ravel(), (fs.T @ fs).ravel(), fs.sum(0)]))
        dist.all_reduce(buf)
        buf = buf.numpy()
        p = x.shape[1]
        t = buf[:p * k].reshape(p, k)
        ftf = buf[p * k:p * k + k * k].reshape(k, k)
        csf = buf[p * k + k * k:]
        # G, S, lambda, mu: replicated
        num = t @ s
        den = (g @ s.T) @ (ftf @ s)
        ratio = num / (den + 0.5 * mu[None, :])
        ratio[np.isnan(ratio)] = 1.0
        g = np.abs(g * ratio)
        a = t.T @ g
        den = (ftf @ s) @ (g.T @ g)
        ratio = a / den
        ratio[np.isnan(ratio)] = 1.0
        s = np.abs(s * ratio)
        lam, mu = csf * lam, g.sum(0) * mu
        res = torch.tensor([np.sum((xs - (fs @ s) @ g.T) ** 2)])
        dist.all_reduce(res)
        errs.append(float(res.item() / xn.item()))
    return fs, s, g, lam, mu, np.array(errs), (b, e)


def worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, f, s, g = problem()
    fs, s2, g2, lam, mu, errs, (b, e) = sharded_sweeps(rank, world, x, f, s, g, 6)
    # independent fits: each rank runs the fits assign_fits gives it, rank 0 gathers the errors
    ks = [3, 4, 5, 6, 7, 8]
    where, _ = sharding.assign_fits([sharding.fit_cost([x.shape], k) for k in ks], world)
    mine = {}
    z = np.zeros((1, 1))
    for i, k in enumerate(ks):
        if where[i] != rank:
            continue
        rng = np.random.default_rng(100 + k)
        fk, sk, gk = synth.random_factors(x.shape[0], x.shape[1], k, rng)
        r = O.res_nmtf_loop([x], None, None, [None], [None], [fk], [sk], [gk], [k], z, z, z, n_iters=3)
        mine[k] = r["All_Error"]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), fs=fs, s=s2, g=g2, lam=lam, mu=mu, errs=errs,
             rows=np.array([b, e]))
    if rank == 0:
        merged = {}
        for d in gathered:
            merged.update(d)
        np.savez(os.path.join(out_dir, "fits.npz"), **{str(k): v for k, v in merged.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_sweeps_and_fit_placement(tmp_path):
    mp.spawn(worker, args=(WORLD, free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    x, f, s, g = problem()
    z = np.zeros((1, 1))
    ref = O.res_nmtf_loop([x], None, None, [None], [None], [f], [s], [g], [s.shape[0]], z, z, z, n_iters=6)
    f_parts = []
    for r in range(WORLD):
        d = np.load(os.path.join(tmp_path, f"rank{r}.npz"))
        f_parts.append(d["fs"])
        np.testing.assert_allclose(d["s"], ref["raw_s"][0], rtol=1e-10)
        np.testing.assert_allclose(d["g"], ref["raw_g"][0], rtol=1e-10)
        np.testing.assert_allclose(d["lam"], ref["lambda"][0], rtol=1e-10)
        np.testing.assert_allclose(d["mu"], ref["mu"][0], rtol=1e-10)
        np.testing.assert_allclose(d["errs"], ref["All_Error"], rtol=1e-10)
    np.testing.assert_allclose(np.concatenate(f_parts), ref["raw_f"][0], rtol=1e-10)
    fits = np.load(os.path.join(tmp_path, "fits.npz"))
    assert sorted(int(k) for k in fits.files) == [3, 4, 5, 6, 7, 8]
    for k in (3, 8):
        rng = np.random.default_rng(100 + k)
        fk, sk, gk = synth.random_factors(x.shape[0], x.shape[1], k, rng)
        r = O.res_nmtf_loop([x], None, None, [None], [None], [fk], [sk], [gk], [k], z, z, z, n_iters=3)
        np.testing.assert_allclose(fits[str(k)], r["All_Error"], rtol=1e-12)


def test_row_shards_cover_panels():
    for n, w in [(333, 2), (64, 4), (1, 3), (2_000_000, 8), (250_000, 8)]:
        sh = sharding.row_shards(n, w)
        assert sh[0][0] == 0 and sh[-1][1] == n
        for (b0, e0), (b1, e1) in zip(sh, sh[1:]):
            assert e0 == b1 and (b0 % 64 == 0 or b0 == n)  # empty tail shards start at n
        sizes = [(e - b + 63) // 64 for b, e in sh]
        assert max(sizes) - min(sizes) <= 1


def test_assign_fits_balances_longest_first():
    costs = [sharding.fit_cost([(20000, 4000)], k) for k in range(3, 9)] * 6  # 36 sweep fits
    where, load = sharding.assign_fits(costs, 8)
    assert len(where) == 36 and set(where) == set(range(8))
    assert max(load) / (sum(load) / 8) < 1.15
