"""Host-side partitioning for the multi-GPU paths.

* ``assign_fits``: the independent fits of one ``apply_resnmtf`` call (each k of the sweep, the shuffled
  refits, the stability resamples; R/main.r:270-299, R/obtain_bicl.r:31-42, R/stability_analysis.r:302-338)
  are placed on ranks longest-first; no collective is involved.
* ``row_shards``: a single view too large for one GPU is split into contiguous ranges of 64-row panels
  (the unit of the device layout); rank r owns rows [begin, end).  Per sweep the ranks all-reduce
  [X'F | F'F | colSums(F)] once (p*k + k*k + k doubles) and everything after is computed redundantly.
"""
from __future__ import annotations

PANEL_ROWS = 64


def row_shards(n_rows, world):
    """[(begin, end)] per rank: whole panels, sizes differing by at most one panel, empty tail ranks allowed."""
    panels = (int(n_rows) + PANEL_ROWS - 1) // PANEL_ROWS
    q, rem = divmod(panels, int(world))
    out, start = [], 0
    for r in range(int(world)):
        cnt = q + (1 if r < rem else 0)
        begin = min(start * PANEL_ROWS, n_rows)
        end = min((start + cnt) * PANEL_ROWS, n_rows)
        out.append((begin, end))
        start += cnt
    return out


def fit_cost(shapes, k, iters=1.0):
    """Relative cost of a fit: bytes streamed per sweep (2 passes over every view) times sweeps."""
    return float(iters) * sum(2.0 * n * p + 4.0 * (n + p) * k for n, p in shapes)


def assign_fits(costs, world):
    """Longest-processing-time-first placement.  Returns (rank of every fit, load per rank)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * int(world)
    where = [0] * len(costs)
    for i in order:
        r = min(range(int(world)), key=lambda j: (load[j], j))
        where[i] = r
        load[r] += costs[i]
    return where, load
