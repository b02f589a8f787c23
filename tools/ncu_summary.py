#!/usr/bin/env python
"""Text summary of an `ncu --set full` report for profiles/: per captured launch, the metrics the roofline argument uses
(duration, tensor-pipe %, DRAM bytes and %, L2 / L1 throughput %, occupancy limiters, registers).  Runs here (no GPU):

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/r02_ncu_full_x.txt
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


if __name__ == "__main__":
    main()
