#!/usr/bin/env python
"""Text summary of an `ncu --set full` report for profiles/: per captured launch, the metrics the roofline argument uses
(duration, tensor-pipe %, DRAM bytes and %, L2 / L1 throughput %, occupancy limiters, registers).  Runs here (no GPU):

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--stalls] > profiles/r02_ncu_full_x.txt

--stalls adds, for the report's first launch, the warp-state sampling totals (why warps did not issue) and the SASS
instructions with the most samples / the most shared-memory wavefronts (ncu --page source --print-source sass).
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max"]


STALLS = ["barrier", "branch_resolving", "dispatch_stall", "lg_throttle", "long_scoreboard", "math_pipe_throttle", "membar", "mio_throttle",
          "no_instructions", "not_selected", "selected", "short_scoreboard", "sleeping", "tex_throttle", "wait"]


def stalls(rep, hdr, units, row):
    ix = {h: i for i, h in enumerate(hdr)}
    tot = 0.0
    vals = []
    for k in STALLS:
        m = "smsp__pcsamp_warps_issue_stalled_%s" % k
        if m in ix:
            try:
                v = float(row[ix[m]])
            except ValueError:
                v = 0.0
            vals.append((v, k)); tot += v
    print("\n  warp-state samples of the first launch (%d): %s" % (tot, "  ".join("%s %.1f%%" % (k, 100 * v / tot) for v, k in sorted(vals, reverse=True) if v / max(tot, 1) >= 0.01)))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not starts:
        return
    h = rows[starts[0]]
    end = starts[1] - 2 if len(starts) > 1 else len(rows)
    jx = {c: i for i, c in enumerate(h)}
    data = [r for r in rows[starts[0] + 1:end] if len(r) == len(h)]

    def f(r, k):
        try:
            return float(r[jx[k]])
        except (ValueError, KeyError):
            return 0.0
    n = sum(f(r, "# Samples") for r in data) or 1.0
    print("  SASS instructions with the most samples (share of samples, times executed, instruction):")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:14]:
        print("    %5.1f%%  x%-9d %s" % (100 * f(r, "# Samples") / n, f(r, "Instructions Executed"), r[jx["Source"]].strip()[:110]))
    wf = sum(f(r, "L1 Wavefronts Shared") for r in data)
    ex = sum(f(r, "L1 Wavefronts Shared Excessive") for r in data)
    if wf:
        print("  shared-memory wavefronts (LSU): %d, of which excessive (bank conflicts) %d = %.1f%%" % (wf, ex, 100 * ex / wf))
        for r in sorted(data, key=lambda r: -f(r, "L1 Wavefronts Shared"))[:4]:
            print("    wavefronts %-10d ideal %-10d %s" % (f(r, "L1 Wavefronts Shared"), f(r, "L1 Wavefronts Shared Ideal"), r[jx["Source"]].strip()[:100]))


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("# %s  (ncu --set full --clock-control none; metrics from --page raw)" % rep.split("/")[-1])
    for r in rows[2:]:
        print("\nkernel: %s   grid %s block %s" % (r[ix["Kernel Name"]], r[ix.get("Grid Size", 0)], r[ix.get("Block Size", 0)]))
        for w in WANT:
            if w in ix:
                print("  %-66s %18s %s" % (w, r[ix[w]], units[ix[w]]))
    if "--stalls" in sys.argv[2:] and len(rows) > 2:
        stalls(rep, hdr, units, rows[2])


if __name__ == "__main__":
    main()
