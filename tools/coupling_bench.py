"""Cost of the coupling terms on the fused path: two 50000x5000 views, k = 5, with no coupling / phi / psi / xi.
Usage: python tools/coupling_bench.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import _lib as L  # noqa: E402
from small_bench import run, sym  # noqa: E402
from resnmtf_b200.device import Context  # noqa: E402

ctx = Context()
shapes = [(50000, 5000)] * 2
for impl, nm in ((L.IMPL_TMA, "two-pass"), (L.IMPL_FUSED, "fused")):
    run(ctx, f"{nm}: uncoupled", shapes, 5, iters=20, impl=impl)
    run(ctx, f"{nm}: phi", shapes, 5, phi=sym(2, 200.0), iters=20, impl=impl)
    run(ctx, f"{nm}: psi", shapes, 5, psi=sym(2, 200.0), iters=20, impl=impl)
    run(ctx, f"{nm}: xi", shapes, 5, xi=sym(2, 50.0), iters=20, impl=impl)
ctx.close()
