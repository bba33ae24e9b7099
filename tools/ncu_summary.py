#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): one block of the metrics that matter per captured kernel.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls]
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==== %s" % r[hdr.index("Kernel Name")][:110])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-70s %s %s" % (w, r[i], units[i]))
        if "--stalls" in sys.argv:
            st = [(float(r[i]), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
            for v, h in sorted(st, reverse=True)[:8]:
                print("  stall %-80s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    main()
