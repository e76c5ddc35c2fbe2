"""Per-kernel timing of the update sweep on one view (default: the C2 shape 20000 x 4000), via the C ABI's
resnmtf_fit_profile (CUDA events between launches).  Prints achieved algorithmic GB/s per streaming pass.
Usage: python tools/kernel_bench.py [--n N --p P --ks 3,5,8 --impls 1,2 --iters 10]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import _lib as L  # noqa: E402
from resnmtf_b200 import synth  # noqa: E402
from resnmtf_b200.device import Context, DeviceFit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--p", type=int, default=4000)
    ap.add_argument("--ks", default="3,5,8")
    ap.add_argument("--impls", default="1,2")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--err", type=int, default=L.ERR_ALGEBRAIC)
    args = ap.parse_args()
    rng = np.random.default_rng(1)
    t0 = time.time()
    x = np.asfortranarray(rng.random((args.p, args.n)).T)  # column-major n x p
    x /= x.sum(axis=0)[None, :]
    print(f"data {args.n}x{args.p} ready in {time.time() - t0:.1f}s", flush=True)
    ctx = Context()
    xbytes = 8.0 * args.n * args.p
    for k in [int(s) for s in args.ks.split(",")]:
        f, s, g = synth.random_factors(args.n, args.p, k, rng)
        for impl in [int(s_) for s_ in args.impls.split(",")]:
            fit = DeviceFit(ctx, [args.n], [args.p], [k])
            fit.set_options(err_mode=args.err, impl=impl)
            fit.set_data(0, x)
            fit.set_factors(0, f, s, g)
            fit.run(3)  # warm-up (graph path)
            prof = fit.profile(args.iters)
            fit.set_factors(0, f, s, g)
            fit.run(args.iters)
            c = fit.counters()
            line = f"k={k} impl={ {1: 'DFMA', 2: 'DMMA', 3: 'TMA', 4: 'FUSED'}.get(c['impl'], c['impl']) }"
            for name in ("f_step", "g_stream", "fused_step", "residual", "finish"):
                ms = prof[name]["ms"] / max(1, args.iters)
                line += f" | {name} {ms * 1e3:8.1f} us"
                if ms > 0 and name in ("f_step", "g_stream", "fused_step"):
                    line += f" ({xbytes / ms * 1e-6:7.0f} GB/s)"
            it_ms = c["device_ms"] / args.iters
            line += f" || graph: {it_ms * 1e3:8.1f} us/iter, {c['alg_bytes_per_iter'] / it_ms * 1e-6:7.0f} GB/s alg"
            print(line, flush=True)
            fit.close()
    ctx.close()


if __name__ == "__main__":
    main()
