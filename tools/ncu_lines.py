#!/usr/bin/env python3
"""tools/ncu_lines.py <report.ncu-rep> — warp instructions executed and stall samples per CUDA source line (ncu source
page with --import-source), top lines first.  Run on the CPU box."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if "# Samples" in r)
h = rows[hdr]
i_src, i_samp, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
lines = {}
cur = None
tot_ex = tot_s = 0
for r in rows[hdr + 1:]:
    if len(r) <= i_ex:
        continue
    src = r[i_src]
    try:
        ex, sa = int(r[i_ex]), int(r[i_samp])
    except ValueError:
        continue
    # in cuda,sass view a CUDA line row is followed by its SASS rows; CUDA rows carry the aggregated counts
    if r[0].strip().isdigit() or not r[0].startswith("0x"):
        cur = (r[0], src.strip()[:120])
        lines[cur] = [ex, sa]
        tot_ex += ex
        tot_s += sa
print("total warp instructions %d, samples %d" % (tot_ex, tot_s))
for (ln, src), (ex, sa) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6s %6.2f%% inst %6.2f%% samp  %s" % (ln, 100.0 * ex / max(tot_ex, 1), 100.0 * sa / max(tot_s, 1), src))
