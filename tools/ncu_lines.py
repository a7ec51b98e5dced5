#!/usr/bin/env python3
"""Per-source-line attribution from an .ncu-rep captured with --import-source on (kernels built with -lineinfo):
executed warp instructions and stall samples by file:line, and by file. Usage: ncu_lines.py report.ncu-rep [top]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur, hdr = None, None
by_line, by_file, text = collections.Counter(), collections.Counter(), {}
samples_line, samples_file = collections.Counter(), collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or cur is None or not r[0].strip().isdigit():
        continue
    try:
        n, s = int(r[iE]), int(r[iS])
    except (ValueError, IndexError):
        continue
    key = (cur, int(r[0]))
    by_line[key] += n; by_file[cur] += n; samples_line[key] += s; samples_file[cur] += s
    text[key] = r[1].strip()[:110]
tot, stot = sum(by_file.values()), sum(samples_file.values())
print("total warp instructions", tot, "samples", stot)
for f, n in by_file.most_common():
    print("%-22s %6.2f%% instr  %6.2f%% samples" % (f, 100.0 * n / tot, 100.0 * samples_file[f] / max(stot, 1)))
print()
for key, n in by_line.most_common(top):
    print("%-20s:%-4d %5.2f%% instr %5.2f%% samples | %s" % (key[0], key[1], 100.0 * n / tot, 100.0 * samples_line[key] / max(stot, 1), text[key]))
