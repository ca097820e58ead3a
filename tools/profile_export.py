#!/usr/bin/env python
"""Export the judged summaries of an .ncu-rep into profiles/ (read here, no GPU):

    python tools/profile_export.py gpurun_out/prof_c3.ncu-rep profiles/r01_step3d_tiled_r4_fast_packed

writes <prefix>_details.csv (the full `--page details` table) and
<prefix>_raw_selected.csv (the raw metrics the roofline and stall analysis in
DESIGN.md quote: duration, DRAM bytes, launch geometry, issue / pipe
utilisation, sampled stall reasons)."""
import csv
import subprocess
import sys

KEEP = ("Kernel Name", "gpu__time_duration", "dram__bytes", "dram__cycles_active",
        "gpu__dram_throughput", "launch__", "sm__throughput", "sm__warps_active",
        "smsp__inst_executed.sum", "smsp__issue_active", "smsp__warps_eligible",
        "smsp__pcsamp_warps_issue_stalled", "sm__inst_executed_pipe_fma.",
        "sm__inst_executed_pipe_alu.", "sm__inst_executed_pipe_lsu.",
        "sm__pipe_fma_cycles_active", "sm__pipe_fmaheavy_cycles_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate", "lts__throughput", "l1tex__throughput",
        "smsp__thread_inst_executed.sum")


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"],
                         capture_output=True, text=True).stdout
    return [r for r in csv.reader(out.splitlines()) if r]


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    with open(prefix + "_details.csv", "w", newline="") as f:
        csv.writer(f, quoting=csv.QUOTE_ALL).writerows(page(rep, "details"))
    rows = page(rep, "raw")
    cols = [i for i, h in enumerate(rows[0]) if h.startswith(KEEP)]
    with open(prefix + "_raw_selected.csv", "w", newline="") as f:
        csv.writer(f).writerows([[r[i] for i in cols] for r in rows])
    hdr = rows[0]
    get = lambda name: rows[2][hdr.index(name)]          # noqa: E731
    print("%s: %s %s, dram read %s %s + write %s %s" % (
        get("Kernel Name")[:60], get("gpu__time_duration.sum"),
        rows[1][hdr.index("gpu__time_duration.sum")], get("dram__bytes_read.sum"),
        rows[1][hdr.index("dram__bytes_read.sum")], get("dram__bytes_write.sum"),
        rows[1][hdr.index("dram__bytes_write.sum")]))


if __name__ == "__main__":
    main()
