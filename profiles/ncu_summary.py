#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key metrics, stall reasons and the hottest SASS lines.

usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--sass N]
"""
import csv
import io
import subprocess
import sys


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep = sys.argv[1]
    nsass = int(sys.argv[sys.argv.index("--sass") + 1]) if "--sass" in sys.argv else 25
    rows = page(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    for k, d in enumerate(data):
        name = d[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("== launch %d: %s" % (k, name[:70]))
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-66s %s %s" % (w, d[i], units[i]))
        stalls = [(h, i) for i, h in enumerate(hdr)
                  if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
        stalls.sort(key=lambda x: -float(d[x[1]] or 0))
        print("  top stall reasons (warps stalled per issue-active cycle):")
        for h, i in stalls[:7]:
            print("    %-40s %s" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), d[i]))
    src = page(rep, "source", ("--print-source", "sass"))
    secs = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    if secs:
        h = src[secs[0] + 1]
        ia, isrc, isamp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
        body = src[secs[0] + 2:(secs[1] if len(secs) > 1 else None)]
        tot = sum(int(r[ia]) for r in body if len(r) > ia)
        tsamp = sum(int(r[isamp]) for r in body if len(r) > isamp)
        print("== SASS: %d warp instructions, %d lines, %d samples; hottest by stall samples:" % (tot, len(body), tsamp))
        for r in sorted(body, key=lambda r: -int(r[isamp]))[:nsass]:
            print("   %6.2f%% samples  %10s exec  %s" % (100.0 * int(r[isamp]) / max(tsamp, 1), r[ia], r[isrc].strip()[:100]))


if __name__ == "__main__":
    main()
