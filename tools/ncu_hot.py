#!/usr/bin/env python
"""Hot SASS instructions (by stall samples) of an ncu report: python tools/ncu_hot.py <report.ncu-rep> [min_pct]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
isrc = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "instructions", len(data))
for k, r in enumerate(data):
    s = int(r[isamp])
    if s > tot * minpct / 100:
        st = sorted([(int(r[i]), hdr[i][6:]) for i in stall_cols if r[i] not in ("", "0")], reverse=True)[:2]
        print("%5d %7d %5.1f%% ex=%9s  %-64s %s" % (k, s, 100 * s / tot, r[iex], r[isrc].strip()[:64], st))
