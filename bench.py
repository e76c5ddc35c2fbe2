#!/usr/bin/env python
"""bench.py -- headline benchmark of the ResNMTF update loop on B200.

Metric (BASELINE.json): update-iterations/s of the F/S/G multiplicative-update sweep, with the achieved
HBM bandwidth against the measured peak.  Workload at one GPU = BASELINE.json configs[1]: a single view
20000 x 4000 (FP64, 640 MB, larger than the 126 MB L2) fitted for every k of the sweep 3..8.

A "step" is ITERS_PER_STEP update-iterations (update_matrices sweep + error, R/main.r:56-80) of each of the six
k-fits = one res_nmtf_inner(n_iters = ITERS_PER_STEP) call per k of the sweep, so that the timed region of the
driver's default run (--steps 20) lasts about a second and the clock sampler sees it.

  value      update-iterations/s with inputs already resident in HBM (device-timed, CUDA events on the
             library's stream, max over ranks)
  e2e        the same through the reference-facing call with HOST buffers, every step: H2D of X (640 MB from
             PAGEABLE memory, which is what R hands over; the pinned figure sits beside it) and of the initial
             factors, ITERS_PER_STEP sweeps per k, D2H of the factors and the error history; host-timed
  roofline   the dominant kernel against the measured HBM copy peak.  Default path: rn_fused_step, the one-pass
             kernel that does the F step and the G step with a single read of X (algorithmic bytes per launch
             8(np + 2nk + 3pk): X once, F read + written, G read as fragments + read + written); with
             RESNMTF_IMPL=3 the two-pass TMA kernels (rn_f_step_tma / rn_g_step_tma, one read of X each)
  cpu_baseline / --impl reference: the NumPy restatement of the reference's own operation sequence
             (3 GEMM passes over X + materialised X_hat, oracle/resnmtf_oracle.py) on the host cores; a step of
             that arm is a bounded sample of the workload: ONE update-iteration of one k (k cycles 3..8)
  sharded    BASELINE configs[4] structure: one view ROW-SHARDED over the N ranks of the run through the library's
             own NCCL communicator (resnmtf_ctx_join), one 250000 x 20000 shard (40 GB) per GPU generated on the
             device; at N = 1 the same code path on a communicator of one rank
  toy        BASELINE configs[0] (README toy, 64 KB of X): update-iterations/s of the persistent single-CTA loop
  k_sweep_wall   wall time of ONE default apply_resnmtf call (k sweep 3..8 + spurious-bicluster removal + stability:
             66 fits) on the bench view through the public API, on the N GPUs of the run (strong scaling: the work is
             fixed) -- first call and steady state -- measured by rank 0 after the timed region, when the other ranks
             have released their GPUs (--no-ksweep skips it)

N > 1 (torchrun, one rank per GPU): the k-sweep / resample fits of the reference are independent, so every
rank runs its own six fits on its own GPU with no data-path collective (weak scaling); value is the sum
over ranks divided by the max time.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS, N_COLS = 20000, 4000
K_SWEEP = (3, 4, 5, 6, 7, 8)
ITERS_PER_STEP = 60
METRIC = "resnmtf_update_iterations_per_second"
UNIT = "update-iterations/s"
WORKLOAD = ("configs[1]: single view 20000x4000 FP64, k sweep 3..8 "
            f"({ITERS_PER_STEP} update-iterations of each k per step)")
SHARD_ROWS, SHARD_COLS, SHARD_K = 250000, 20000, 8


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the row-sharded leg (configs[4] shard per GPU)")
    ap.add_argument("--no-toy", action="store_true", help="skip the configs[0] leg (README toy, persistent single-CTA loop)")
    ap.add_argument("--no-ksweep", action="store_true",
                    help="skip the wall time of the default apply_resnmtf call (BASELINE metric, third part)")
    return ap.parse_args()


def bench_config():
    """The `config` object of the JSON line -- identical for both arms."""
    return {"workload": WORKLOAD, "rows": N_ROWS, "cols": N_COLS, "k_sweep": list(K_SWEEP),
            "update_iterations_per_step": len(K_SWEEP) * ITERS_PER_STEP, "fits_per_gpu": len(K_SWEEP),
            "l2": "inputs larger than L2 (640 MB view; no flush needed)",
            "multi_gpu": "independent k-sweep fits per rank, no data-path collective"}


# ----------------------------------------------------------------------------------------------------
# synthetic workload (planted biclusters + |N(0,1)| noise, reference prep)
# ----------------------------------------------------------------------------------------------------


def make_workload(rank=0):
    from resnmtf_b200 import synth

    rng = np.random.default_rng(synth.config_seed(2, 0) + 1000 * rank)
    x, _, _ = synth.planted_view(N_ROWS, N_COLS, 5, rng, row_prob=0.2, col_prob=0.2, height=5.0, sigma=1.0)
    x = synth.prep(x)
    inits = {k: synth.random_factors(N_ROWS, N_COLS, k, rng) for k in K_SWEEP}
    return x, inits


def toy_leg(iters=20000):
    """BASELINE configs[0] (the README toy: 2 views 100 x 50 and 100 x 30 with shared rows, phi[1,2] = 200, k = 3): its
    64 KB of X are launch-latency-bound, so the whole loop runs as ONE persistent CTA (RESNMTF_IMPL_SMALL); reported as
    update-iterations/s over `iters` fixed sweeps (R/main.r:83-108), one launch per 4096 sweeps (the size of the
    error-history buffer the host drains)."""
    from resnmtf_b200 import _lib as L
    from resnmtf_b200 import synth
    from resnmtf_b200.device import Context, DeviceFit

    try:
        rng = np.random.default_rng(synth.config_seed(1, 0))
        x1, rows, _ = synth.planted_view(100, 50, 3, rng, row_prob=0.5, col_prob=0.4, sigma=1.0)
        x2, _, _ = synth.planted_view(100, 30, 3, rng, row_prob=0.5, col_prob=0.4, sigma=0.01, rows=rows)
        data = [synth.prep(x1), synth.prep(x2)]
        phi = np.zeros((2, 2))
        phi[0, 1] = phi[1, 0] = 200.0
        ctx = Context()
        fit = DeviceFit(ctx, [100, 100], [50, 30], [3, 3])
        fit.set_options(err_mode=L.ERR_ALGEBRAIC)
        for v in range(2):
            fit.set_data(v, data[v])
            fit.set_factors(v, *synth.random_factors(100, data[v].shape[1], 3, rng))
        fit.set_restrictions(phi, None, None)
        idx = np.arange(100, dtype=np.int32)
        fit.set_shared_map(L.MAP_ROW, 0, 1, idx, idx)
        fit.set_shared_map(L.MAP_ROW, 1, 0, idx, idx)
        fit.run(200)
        t0 = time.perf_counter()
        fit.run(iters)
        wall = time.perf_counter() - t0
        c = fit.counters()
        out = {"workload": "configs[0]: README toy, 2 views 100x50 + 100x30, shared rows, phi = 200, k = 3, fixed sweeps",
               "update_iterations_per_s": iters / (c["device_ms"] * 1e-3), "us_per_update_iteration": c["device_ms"] * 1e3 / iters,
               "wall_update_iterations_per_s": iters / wall, "iters": iters, "launches": int(c["kernel_launches"]),
               "impl": int(c["impl"]), "bound": "latency (64 KB of X): one persistent CTA, no roofline"}
        fit.close()
        ctx.close()
        return out
    except Exception as exc:  # noqa: BLE001 - the headline line must still be printed
        return {"error": f"{type(exc).__name__}: {exc}"}


def ksweep_wall(x, n_gpus):
    """BASELINE.json metric, third part ("k-sweep wall time"): wall time of ONE default apply_resnmtf call on the bench
    view -- k sweep 3..8 with bisilhouette selection, spurious-bicluster removal and stability analysis, the reference's
    defaults: 66 convergence loops -- through the public API from the host matrix, on ``n_gpus`` GPUs of this process.
    One small untimed call first (CUDA contexts), then two timed calls: the first is what a user sees, the second is
    the steady state."""
    from resnmtf_b200 import synth
    from resnmtf_b200.api import apply_resnmtf

    try:
        os.environ["RESNMTF_MAX_GPUS"] = str(int(n_gpus))
        small, _, _ = synth.planted_view(1200, 600, 3, np.random.default_rng(1), row_prob=0.3, col_prob=0.3)
        apply_resnmtf([small], k_min=3, k_max=3 + int(n_gpus), spurious=False, stability=False,
                      rng=np.random.default_rng(2), max_iters=50)
        runs, res = [], None
        for _ in range(2):
            t0 = time.perf_counter()
            res = apply_resnmtf([x], k_min=3, k_max=8, rng=np.random.default_rng(5), max_iters=5000)
            runs.append(time.perf_counter() - t0)
        return {"seconds": runs[0], "first_call_seconds": runs[0], "steady_seconds": runs[1], "n_gpus": int(n_gpus),
                "fits": 66, "scaling": "strong",
                "workload": "one default apply_resnmtf call on the bench view: k sweep 3..8 + spurious-bicluster "
                            "removal (5 shuffled refits per fit) + stability analysis (5 resamples), host matrix in, "
                            "result list out",
                "selected_k": int(res["output_f"][0].shape[1]), "bisil": float(res["bisil"]),
                "biclusters_kept": int((res["row_clusters"][0].sum(axis=0) > 0).sum())}
    except Exception as exc:  # noqa: BLE001 - the headline line must still be printed
        return {"error": f"{type(exc).__name__}: {exc}"}


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle's reference-cost restatement, bounded sample
# ----------------------------------------------------------------------------------------------------


def cpu_reference_rate(x, inits, steps, warmup):
    """update-iterations/s of the NumPy/BLAS restatement on ALL host cores (the BLAS thread count is set explicitly:
    torchrun exports OMP_NUM_THREADS=1).  One step = ONE update-iteration (update_matrices + calculate_error) of one k
    on the full view, k cycling through the sweep -- a bounded sample of the GPU arm's step."""
    from oracle import resnmtf_oracle as O

    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits

        limiter = threadpool_limits(limits=cores)
    except Exception:  # pragma: no cover
        limiter = None
    z = np.zeros((1, 1))
    names_r, names_c = O.default_names([x])
    ri, ci = O.shared_names(names_r), O.shared_names(names_c)
    norms = np.array([np.linalg.norm(x, "fro") ** 2])
    state = {}
    for k in K_SWEEP:
        f, s, g = inits[k]
        state[k] = ([f.copy()], [s.copy()], [g.copy()], [f.sum(0)], [g.sum(0)])

    def one_step(i):
        k = K_SWEEP[i % len(K_SWEEP)]
        cf, cs, cg, cl, cm = state[k]
        cf, cs, cg, cl, cm = O.update_matrices([x], cf, cs, cg, cl, cm, z, z, z, ri, ci, names_r, names_c)
        O.calculate_error([x], cf, cs, cg, norms)
        state[k] = (cf, cs, cg, cl, cm)

    for i in range(warmup):
        one_step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        one_step(i)
    dt = time.perf_counter() - t0
    try:
        from threadpoolctl import threadpool_info

        used = max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:  # pragma: no cover
        used = cores
    del limiter
    return {
        "value": steps / dt, "unit": UNIT, "cores": int(used), "kind": "port",
        "sample": f"{steps} steps, each ONE update-iteration of one k (cycling 3..8) on the full 20000x4000 view "
                  f"(NumPy/OpenBLAS restatement of the R operation sequence incl. X_hat; not R)",
        "seconds": dt, "steps": steps,
    }


# ----------------------------------------------------------------------------------------------------
# clocks: NVML polled every 5 ms from a thread (nvidia-smi's 100 ms loop cannot see a sub-second region)
# ----------------------------------------------------------------------------------------------------


class ClockSampler:
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
               (0x4, "sw_power_cap"))

    def __init__(self, cuda_index):
        self.rows, self.stop_flag, self.thread, self.handle, self.nvml = [], False, None, None, None
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            self.nvml = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
        except Exception:
            self.handle = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                self.rows.append((sm, mx, rs, pw))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.handle is None:
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({nm for _, _, rs, _ in self.rows for bit, nm in self.REASONS if rs & bit})
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": float(max(r[1] for r in self.rows)),
                "reasons": reasons, "samples": len(self.rows), "power_w_max": float(max(r[3] for r in self.rows)),
                "how": "NVML polled every 5 ms during the timed region"}


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram bytes per launch (k = 8) of `kernel` from the committed ncu capture, if there is one."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(f"{kernel}_dram_bytes_per_launch")
    except Exception:
        return None


def sharded_leg(torch, dist, local_rank, rank, world, iters=20):
    """One view row-sharded over the ranks of the run through the LIBRARY's communicator (resnmtf_ctx_join ->
    ncclCommInitRank; per sweep one ncclAllReduce of [X'F | F'F | colSums(F)] between the G stream and the stand-alone
    G epilogue), BASELINE configs[4] shard shape per GPU, data generated on the device.  Every rank first runs its
    shard on a communicator of ONE rank (same kernels, same launches) -- the weak-scaling base -- then all ranks join
    one communicator."""
    from resnmtf_b200 import _lib as L
    from resnmtf_b200.device import Context, DeviceFit

    free, _ = torch.cuda.mem_get_info()
    n, p, k = SHARD_ROWS, SHARD_COLS, SHARD_K
    while 2.2 * 8.0 * n * p > free and n > 4096:  # the view lives twice for a moment (generator output + library layout)
        n //= 2
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    xt = torch.rand((p, n), generator=gen, device="cuda", dtype=torch.float64)  # row-major p x n == column-major n x p
    colsum = xt.sum(dim=1, keepdim=True)
    if world > 1:
        dist.all_reduce(colsum)
    xt /= colsum  # L1 column normalisation over the WHOLE view (R/utils.r:86-88)
    torch.cuda.synchronize()
    rng = np.random.default_rng(7)
    g0 = rng.random((p, k)) + 0.05
    g0 /= g0.sum(0)[None, :]
    s0 = np.abs(np.diag(rng.random(k) + 0.5)) + 0.05
    f0 = np.random.default_rng(100 + rank).random((n, k)) + 0.05

    def run(n_ranks, my_rank, comm_id):
        ctx = Context(local_rank)
        ctx.join(comm_id, my_rank, n_ranks)
        fit = DeviceFit(ctx, [n], [p], [k])
        fit.set_options(err_mode=L.ERR_ALGEBRAIC, impl=L.IMPL_AUTO)
        fit.set_data_device(0, xt.data_ptr(), n)
        fit.set_factors(0, np.asfortranarray(f0 / (f0.sum(0)[None, :] * n_ranks)), np.asfortranarray(s0),
                        np.asfortranarray(g0))
        fit.run(3)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        fit.run(iters)
        c = fit.counters()
        err = float(fit.errors()[-1])
        fit.close()
        ctx.close()
        ms = torch.tensor([c["device_ms"]], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / iters, int(c["kernel_launches"]) // iters, int(c["impl"]), err

    one_ms, launches, impl, _ = run(1, 0, Context.comm_id_create())
    all_ms, err = one_ms, None
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(Context.comm_id_create()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, src=0)
        all_ms, launches, impl, err = run(world, rank, bytes(idt.cpu().numpy().tobytes()))
    del xt
    torch.cuda.empty_cache()
    gbs = 16.0 * n * p / (all_ms * 1e-3) * 1e-9
    return {"workload": f"configs[4] structure: single view {world * n}x{p} row-sharded over {world} rank(s), "
                        f"{n}x{p} (={8e-9 * n * p:.0f} GB) per GPU, k={k}, fixed sweeps, data generated on the device",
            "n_ranks": world, "rows_per_rank": n, "cols": p, "k": k, "iters": iters,
            "ms_per_update_iteration": all_ms, "update_iterations_per_s": 1e3 / all_ms,
            "gbs_per_gpu_two_pass_bytes": gbs, "launches_per_update_iteration": launches, "impl": impl,
            "one_rank_ms_per_update_iteration": one_ms, "weak_scaling_efficiency": one_ms / all_ms,
            "collective": "ncclAllReduce(sum, f64) of p*k + k*k + k doubles per sweep on the library's own communicator "
                          "(resnmtf_ctx_join), enqueued on the fit's stream inside the iteration graph",
            "err_last": err}


def run_gpu(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:  # leave the init lines of the communicators (torch's and the library's) in the log
        os.environ.setdefault("NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
    import torch
    import torch.distributed as dist

    from resnmtf_b200 import _lib as L
    from resnmtf_b200.device import Context, DeviceData, DeviceFit

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L.require_device()

    x, inits = make_workload(rank)  # pageable host memory, column-major: what R hands over
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).pin_memory()  # pinned twin; memory == column-major n x p
    x_pinned = xt.numpy().T
    ctx = Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident fits: one per k, all attached to ONE upload of the view ------------------------------
    data = DeviceData(ctx, x_pinned)
    fits = {}
    for k in K_SWEEP:
        fit = DeviceFit(ctx, [N_ROWS], [N_COLS], [k])
        fit.set_options(err_mode=L.ERR_AUTO, impl=L.IMPL_AUTO)
        fit.attach_data(0, data)
        fit.set_factors(0, *inits[k])
        fits[k] = fit

    def do_steps(n):
        launches = 0
        for _ in range(n):
            for k in K_SWEEP:  # the fits are independent, so the order is free
                fits[k].run(ITERS_PER_STEP)
                launches += fits[k].counters()["kernel_launches"]
        return launches

    warm = max(args.warmup, 3)
    do_steps(warm)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    launches = do_steps(args.steps)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    iters_rank = args.steps * len(K_SWEEP) * ITERS_PER_STEP

    # ---- roofline of the dominant kernel, measured live (CUDA events between launches) --------------
    prof_iters = 20
    kern = {"rn_f_step_tma": [0.0, 0.0], "rn_g_step_tma": [0.0, 0.0], "rn_fused_step": [0.0, 0.0]}  # bytes, ms
    per_k = {}
    for k in K_SWEEP:
        pr = fits[k].profile(prof_iters)
        for name, cls, nbytes in (
                ("rn_f_step_tma", "f_step", 8.0 * (N_ROWS * N_COLS + 2 * N_ROWS * k + N_COLS * k)),
                ("rn_g_step_tma", "g_stream", 8.0 * (N_ROWS * N_COLS + N_ROWS * k + 2 * N_COLS * k)),
                ("rn_fused_step", "fused_step", 8.0 * (N_ROWS * N_COLS + 2 * N_ROWS * k + 3 * N_COLS * k))):
            if pr[cls]["ms"] > 0.0:
                kern[name][0] += prof_iters * nbytes
                kern[name][1] += pr[cls]["ms"]
                per_k.setdefault(name, {})[str(k)] = 1e3 * pr[cls]["ms"] / prof_iters
    peak, peak_src = measured_peak()
    ran = {nm: bm for nm, bm in kern.items() if bm[1] > 0.0}
    dom = max(ran, key=lambda nm: ran[nm][1])  # the kernel the step spends most of its time in
    achieved = ran[dom][0] / ran[dom][1] * 1e-6
    alg_bytes_iter = sum(fits[k].counters()["alg_bytes_per_iter"] for k in K_SWEEP)  # one sweep of each k

    isolated = {"achieved": achieved, "us_per_launch": 1e3 * ran[dom][1] / (prof_iters * len(K_SWEEP)),
                "us_per_launch_by_k": per_k.get(dom),
                "how": "CUDA events between individual launches, no graph (resnmtf_fit_profile)"}
    us_per_launch = isolated["us_per_launch"]
    if len(ran) == 1 and launches == iters_rank:
        # the timed region is nothing but launches of this kernel (one per update-iteration): its average launch
        # duration is the region's CUDA-event time over the launches -- how the kernel runs in production, incl. the
        # overlap of consecutive launches (programmatic dependent launch)
        us_per_launch = 1e3 * ms / launches
        region_bytes = args.steps * ITERS_PER_STEP * sum(
            8.0 * (N_ROWS * N_COLS + 2 * N_ROWS * k + 3 * N_COLS * k) for k in K_SWEEP)
        achieved = region_bytes / ms * 1e-6
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "peak_source": peak_src, "traffic": ncu_traffic(dom),
        "us_per_launch": us_per_launch, "isolated": isolated,
        "other_kernels": {nm: {"achieved": bm[0] / bm[1] * 1e-6} for nm, bm in ran.items() if nm != dom},
        "whole_step_achieved": alg_bytes_iter * ITERS_PER_STEP * args.steps / ms * 1e-6,
        "alg_bytes_per_step": alg_bytes_iter * ITERS_PER_STEP,
        "note": ("one-pass kernel: X is read once per update-iteration (B_min of SURVEY 8d); the co-limiter is the "
                 "FP64 tensor pipe (DMMA)" if dom == "rn_fused_step"
                 else "two-pass kernels: X is read once per kernel, twice per iteration"),
    }
    if dom == "rn_fused_step":
        fma = 2.0 * N_ROWS * N_COLS * 8 * prof_iters * len(K_SWEEP)
        roofline["fp64_mma"] = {"achieved_tfma_per_s": fma / (ran[dom][1] * 1e-3) * 1e-12, "chip_peak_tfma_per_s": 18.5,
                                "peak_source": "tools/microbench.cu (DESIGN.md section 4), not MEASURED_PEAKS.json"}

    # ---- e2e: reference-facing fit calls with HOST buffers, X uploaded every step ----------------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 5))

        def e2e_step(xhost):
            # what apply_resnmtf's k-sweep does (R/main.r:279-287): the host view goes to the device once per call,
            # then one res_nmtf_inner-style fit per k with its own host inits and host results
            h2d = d2h = 0
            dd = DeviceData(ctx, xhost)
            h2d += xhost.nbytes
            for k in K_SWEEP:
                f0, s0, g0 = inits[k]
                fit = DeviceFit(ctx, [N_ROWS], [N_COLS], [k])
                fit.attach_data(0, dd)
                fit.set_factors(0, f0, s0, g0)
                fit.run(ITERS_PER_STEP)
                errs = fit.errors()
                fit.normalise()
                f, s, g, lam, mu = fit.get_factors(0)
                fit.close()
                h2d += f0.nbytes + s0.nbytes + g0.nbytes
                d2h += f.nbytes + s.nbytes + g.nbytes + lam.nbytes + mu.nbytes + errs.nbytes
            dd.close()
            return h2d, d2h

        def e2e_time(xhost):
            e2e_step(xhost)  # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                h2d, d2h = e2e_step(xhost)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return world * e2e_steps * len(K_SWEEP) * ITERS_PER_STEP / dt, dt, h2d, d2h

        v_page, dt_page, h2d, d2h = e2e_time(x)
        v_pin, dt_pin, _, _ = e2e_time(x_pinned)
        e2e = {"value": v_page, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "seconds": dt_page, "host_memory": "pageable (what R hands over)",
               "pinned": {"value": v_pin, "seconds": dt_pin},
               "call": "every step: data_create(host X, 640 MB H2D) + per k of the sweep fit_create + attach_data + "
                       f"set_factors(host) + run(n_iters={ITERS_PER_STEP}) + normalise + get_factors/get_errors"}

    # ---- reduce over ranks -------------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        ln = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        launches = int(ln.item())
    value = world * iters_rank / (ms * 1e-3)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_rate(x, inits, steps=12, warmup=1)

    for f in fits.values():
        f.close()
    data.close()
    ctx.close()
    del xt, x_pinned

    sharded = None
    if not args.no_sharded:
        try:
            sharded = sharded_leg(torch, dist, local_rank, rank, world)
        except Exception as exc:  # noqa: BLE001 - the headline line must still be printed
            sharded = {"error": f"{type(exc).__name__}: {exc}"}

    if world > 1:  # the ranks are done with each other: the call below is one process over all the GPUs
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    toy = None if args.no_toy else toy_leg()
    ksweep = None if args.no_ksweep else ksweep_wall(x, world)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(),
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
        "cpu_baseline": cpu, "sharded": sharded, "toy": toy, "k_sweep_wall": ksweep, "impl": "b200",
    }
    print(json.dumps(line), flush=True)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (R is absent from this image,
    so this is the oracle's NumPy/BLAS restatement of the same operation sequence), rank 0 only, all host cores.
    --steps / --warmup are honoured; a step of this arm is a bounded sample of the GPU arm's step (ONE
    update-iteration of one k on the full view, k cycling), the value is normalised to update-iterations/s."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    x, inits = make_workload(0)
    res = cpu_reference_rate(x, inits, steps=max(1, args.steps), warmup=max(0, args.warmup))
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / max(res["steps"], 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(),
        "impl": "reference",
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
