"""Row-sharded single view (the BASELINE configs[4] structure): every rank holds a row shard of one view resident on
its GPU; per update-iteration the ranks all-reduce [X'F | F'F | colSums(F)] over NCCL.  Data are generated on the
device (a 40 GB shard is not worth a host pass).  Launch with torchrun, one rank per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/sharded_bench.py [--rows-per-rank 250000 --cols 20000 --k 8 --iters 20]"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnmtf_b200 import _lib as L  # noqa: E402
from resnmtf_b200.device import Context, DeviceFit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows-per-rank", type=int, default=250000)
ap.add_argument("--cols", type=int, default=20000)
ap.add_argument("--k", type=int, default=8)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
ctx = Context(local)
if world > 1:
    ids = [Context.comm_id_create() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.join(ids[0], rank, world)
n, p, k = a.rows_per_rank, a.cols, a.k
gen = torch.Generator(device="cuda")
gen.manual_seed(1234 + rank)
xt = torch.rand((p, n), generator=gen, device="cuda", dtype=torch.float64)  # row-major p x n == column-major n x p
colsum = xt.sum(dim=1, keepdim=True)
if world > 1:
    cs = colsum.cpu()
    dist.all_reduce(cs)
    colsum = cs.cuda()
xt /= colsum  # L1 column normalisation over the WHOLE view
torch.cuda.synchronize()
rng = np.random.default_rng(7)
g0 = rng.random((p, k)) + 0.05
g0 /= g0.sum(0)[None, :]
s0 = np.abs(np.diag(rng.random(k) + 0.5)) + 0.05
f0 = np.random.default_rng(100 + rank).random((n, k)) + 0.05
f0 /= (f0.sum(0)[None, :] * world)
fit = DeviceFit(ctx, [n], [p], [k])
fit.set_options(err_mode=L.ERR_ALGEBRAIC, impl=L.IMPL_AUTO)
fit.set_data_device(0, xt.data_ptr(), n)
del xt
torch.cuda.empty_cache()
fit.set_factors(0, np.asfortranarray(f0), np.asfortranarray(s0), np.asfortranarray(g0))
fit.run(3)
if world > 1:
    dist.barrier()
fit.run(a.iters)
c = fit.counters()
ms = torch.tensor([c["device_ms"]])
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    it_ms = float(ms.item()) / a.iters
    gb = world * 16.0 * n * p / (it_ms * 1e-3) * 1e-9
    print(f"row-sharded view {world * n}x{p} over {world} GPU(s), k={k}, impl {c['impl']}: {it_ms:.3f} ms per "
          f"update-iteration = {1e3 / it_ms:.1f} it/s, {gb:.0f} GB/s of two-pass X bytes in aggregate "
          f"({gb / world:.0f} per GPU), err {fit.errors()[-1]:.6f}", flush=True)
fit.close()
ctx.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
