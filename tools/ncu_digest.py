"""Digest of an .ncu-rep (read here on the CPU box): headline metrics, stall-reason shares and the hottest SASS
instructions of the first kernel in the report.  Usage: python tools/ncu_digest.py gpurun_out/prof.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = page(rep, "raw")
    hdr, units, r = rows[0], rows[1], rows[2]
    get = lambda k: next((r[i] for i, h in enumerate(hdr) if h == k), None)
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
            "launch__cluster_size", "launch__registers_per_thread", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
            "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "smsp__issue_active.avg.per_cycle_active", "lts__t_bytes.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
    print("kernel:", get("Kernel Name"))
    for k in keys:
        u = next((units[i] for i, h in enumerate(hdr) if h == k), "")
        print(f"  {k} = {get(k)} {u}")
    st = []
    for i, h in enumerate(hdr):
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try:
                st.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1.0
    print("stalls:", ", ".join(f"{h} {100 * v / tot:.1f}%" for v, h in sorted(st, reverse=True)[:8]))
    src = page(rep, "source", ("--print-source", "sass"))
    h2 = src[1]
    idx = {h: i for i, h in enumerate(h2)}
    body = []
    for row in src[2:]:
        if len(row) < len(h2) or row[0] in ("Kernel Name", "Address"):
            break
        body.append(row)
    S, E = idx["Warp Stall Sampling (All Samples)"], idx["Instructions Executed"]
    stall_cols = [h for h in h2 if h.startswith("stall_") and "Not" not in h]
    total = sum(int(x[S] or 0) for x in body) or 1
    print(f"{len(body)} SASS instructions, {total} samples")
    hot = sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:top]
    for i in sorted(hot):
        x = body[i]
        m = sorted(((int(x[idx[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        print(f"{i:5d} {int(x[S] or 0):6d} {100 * int(x[S] or 0) / total:5.1f}% exec={x[E]:>8} {x[idx['Source']][:60]:60s} {m}")


if __name__ == "__main__":
    main()
