#!/usr/bin/env python3
"""Developer tool: hottest source lines (warp-stall samples) of one kernel in an .ncu-rep.
usage: ncu_hot_lines.py REPORT KERNEL_NAME [TOP]"""
import csv, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel, "--print-source",
                      "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, acc = None, []
for r in rows:
    if r and r[0] == "Line No":
        if hdr is not None:
            break  # first kernel instance only
        hdr = r
        continue
    if hdr and len(r) > 8 and r[0].isdigit():
        try:
            acc.append((int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]), int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
total = sum(a[0] for a in acc) or 1
acc.sort(reverse=True)
print("total samples", total)
for a in acc[:top]:
    print("%5.1f%% %9d  L%-5d %s" % (100.0 * a[0] / total, a[1], a[2], a[3]))
