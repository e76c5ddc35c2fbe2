"""Generates the golden fixtures under tests/golden/ from the CPU oracle (oracle/resnmtf_oracle.py).

The reference is pure R and cannot run in this image (no R, no rpy2), and its own tests contain no numeric
vectors, so these fixtures are produced by the oracle after it has been pinned against the reference's
property tests (tests/test_oracle_reference_properties.py).  They freeze the oracle's numbers so that (a) a
later change to the oracle is noticed and (b) the GPU box can be checked against committed values.
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import resnmtf_oracle as O  # noqa: E402
from resnmtf_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N_SWEEPS = 10


def case_readme_toy():
    """BASELINE.json configs[0]: 2 views 100x50 and 100x30 with shared rows, phi[1,2] = 200, k = 3."""
    rng = np.random.default_rng(synth.config_seed(1, 0))
    x1, rows, _ = synth.planted_view(100, 50, 3, rng, row_prob=0.5, col_prob=0.4, sigma=1.0)
    x2, _, _ = synth.planted_view(100, 30, 3, rng, row_prob=0.5, col_prob=0.4, sigma=0.01, rows=rows)
    data = [synth.prep(x1), synth.prep(x2)]
    rn = [[f"row_{i}" for i in range(1, 101)]] * 2
    cn = [[f"col_{i}" for i in range(1, 51)], [f"col_{i}" for i in range(51, 81)]]
    phi = np.zeros((2, 2))
    phi[0, 1] = 200.0
    return dict(data=data, k=[3, 3], rn=rn, cn=cn, phi=O.init_rest_mats(phi, 2), xi=np.zeros((2, 2)),
                psi=np.zeros((2, 2)), seed=11)


def case_single_view():
    rng = np.random.default_rng(5)
    x, _, _ = synth.planted_view(90, 61, 3, rng, row_prob=0.3, col_prob=0.3)
    return dict(data=[synth.prep(x)], k=[4], rn=None, cn=None, phi=np.zeros((1, 1)), xi=np.zeros((1, 1)),
                psi=np.zeros((1, 1)), seed=12)


def case_three_views_all_restrictions():
    """phi, psi and xi all non-zero, partially overlapping and permuted names."""
    rng = np.random.default_rng(6)
    n, p = 70, 48
    data = [synth.prep(synth.planted_view(n, p, 3, rng, 0.35, 0.35)[0]) for _ in range(3)]
    rn = [[f"r{i}" for i in range(n)],
          [f"r{i}" for i in rng.permutation(n)],
          [f"r{i}" for i in range(20, n)] + [f"s{i}" for i in range(20)]]
    cn = [[f"c{i}" for i in range(p)],
          [f"c{i}" for i in range(10, p)] + [f"d{i}" for i in range(10)],
          [f"c{i}" for i in rng.permutation(p)]]
    phi = np.zeros((3, 3)); phi[0, 1] = 200.0; phi[1, 2] = 50.0
    psi = np.zeros((3, 3)); psi[0, 2] = 100.0; psi[0, 1] = 30.0
    xi = np.zeros((3, 3)); xi[0, 1] = 50.0; xi[0, 2] = 20.0
    return dict(data=data, k=[3, 3, 3], rn=rn, cn=cn, phi=O.init_rest_mats(phi, 3), xi=O.init_rest_mats(xi, 3),
                psi=O.init_rest_mats(psi, 3), seed=13)


def build(case):
    data, k = case["data"], case["k"]
    V = len(data)
    dn_r, dn_c = O.default_names(data)
    rn = case["rn"] or dn_r
    cn = case["cn"] or dn_c
    rng = np.random.default_rng(case["seed"])
    noise = [np.abs(np.sqrt(0.05) * rng.standard_normal((kk, kk))) for kk in k]
    f0, s0, g0, _, _ = O.init_mats_inner(data, k, noise)  # the reference's SVD initialisation
    ri, ci = O.shared_names(rn), O.shared_names(cn)
    out = {"n_views": np.array(V), "k": np.array(k)}
    for v in range(V):
        out[f"x{v}"] = data[v]
        out[f"rn{v}"] = np.array(rn[v])
        out[f"cn{v}"] = np.array(cn[v])
        out[f"f0_{v}"], out[f"s0_{v}"], out[f"g0_{v}"] = f0[v], s0[v], g0[v]
    out["phi"], out["xi"], out["psi"] = case["phi"], case["xi"], case["psi"]
    errs = []
    fixed = O.res_nmtf_loop(data, ri, ci, rn, cn, f0, s0, g0, k, case["phi"], case["xi"], case["psi"],
                            n_iters=N_SWEEPS, trace=lambda t, cf, cs, cg, cl, cm, e: errs.append(e.copy()))
    out["sweep_errors"] = np.array(errs)
    for v in range(V):
        out[f"f{N_SWEEPS}_{v}"] = fixed["raw_f"][v]
        out[f"s{N_SWEEPS}_{v}"] = fixed["raw_s"][v]
        out[f"g{N_SWEEPS}_{v}"] = fixed["raw_g"][v]
        out[f"lam{N_SWEEPS}_{v}"] = fixed["lambda"][v]
        out[f"mu{N_SWEEPS}_{v}"] = fixed["mu"][v]
    conv = O.res_nmtf_loop(data, ri, ci, rn, cn, f0, s0, g0, k, case["phi"], case["xi"], case["psi"])
    out["all_error"] = conv["All_Error"]
    rows, cols, rel = O.binarise(conv["output_f"], conv["output_g"], conv["output_s"])
    for v in range(V):
        out[f"of_{v}"], out[f"os_{v}"], out[f"og_{v}"] = conv["output_f"][v], conv["output_s"][v], conv["output_g"][v]
        out[f"rows_{v}"], out[f"cols_{v}"] = rows[v].astype(np.uint8), cols[v].astype(np.uint8)
    return out


if __name__ == "__main__":
    for name, fn in (("readme_toy", case_readme_toy), ("single_view", case_single_view),
                     ("three_views", case_three_views_all_restrictions)):
        out = build(fn())
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "sweeps to converge:", len(out["all_error"]), "file KB:", os.path.getsize(path) // 1024)
