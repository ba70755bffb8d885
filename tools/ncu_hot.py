#!/usr/bin/env python
"""Instruction / stall-sample distribution of an ncu --set full capture, in blocks of SASS instructions.
usage: ncu_hot.py report.ncu-rep [block]"""
import csv, subprocess, sys
rep = sys.argv[1]; blk = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
iS = hdr.index("Source"); iN = hdr.index("# Samples"); iE = hdr.index("Instructions Executed")
E = [int(r[iE]) for r in data]; N = [int(r[iN]) for r in data]
tot = sum(E); tots = sum(N)
print("warp instructions", tot, "samples", tots, "SASS instructions", len(data))
for b in range(0, len(data), blk):
    e = sum(E[b:b + blk]); n = sum(N[b:b + blk])
    if e > tot * 0.01 or n > tots * 0.01:
        print(f"[{b:4d}-{b+blk:4d}) inst {100*e/tot:5.1f}%  samples {100*n/tots:5.1f}%  avg exec {e/blk/1e6:8.2f}M  first: {data[b][iS].strip()[:60]}")
