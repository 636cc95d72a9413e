#!/usr/bin/env python3
"""Turns an .ncu-rep (ncu --set full) into the markdown summary committed under profiles/.
    python tools/ncu_summary.py gpurun_out/prof_scan.ncu-rep profiles/r01_ncu_cosine_scan_bulk.md "title" "notes..."
"""
import csv
import io
import subprocess
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_hmma.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    notes = sys.argv[4:] 
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {k: (v, u) for k, u, v in zip(hdr, units, vals)}
    with open(out, "w") as f:
        f.write("# %s\n\n" % title)
        for n in notes:
            f.write(n + "\n\n")
        f.write("| metric | value | unit |\n|---|---|---|\n")
        for k in KEEP:
            if k in d and d[k][0] not in ("", "n/a"):
                f.write("| %s | %s | %s |\n" % (k, d[k][0], d[k][1]))
    tr = sum(float(d[k][0]) * SCALE.get(d[k][1], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in d)
    print("dram traffic bytes per launch:", tr, " duration:", d.get("gpu__time_duration.sum"))


if __name__ == "__main__":
    main()
