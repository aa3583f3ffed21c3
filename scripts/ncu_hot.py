#!/usr/bin/env python
"""Top stall lines of one kernel from an .ncu-rep: python scripts/ncu_hot.py rep kernel-regex [N]"""
import csv, subprocess, sys, io, collections
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = None; out = []; nk = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        nk += 1
        if nk > 1: break
        continue
    if r and r[0] == "Address": h = r; continue
    if h and len(r) == len(h): out.append(r)
idx = {n: i for i, n in enumerate(h)}
tot = sum(int(r[idx["# Samples"]]) for r in out)
print(len(out), "SASS lines;", tot, "samples")
mix = collections.Counter()
for r in out:
    toks = r[idx["Source"]].strip().split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    mix[op.split('.')[0]] += int(r[idx["Instructions Executed"]])
print(mix.most_common(16))
# print a window with samples: line number, samples, instr
for i, r in enumerate(out):
    r.append(i)
top = sorted(out, key=lambda r: -int(r[idx["# Samples"]]))[:N]
for r in sorted(top, key=lambda r: r[-1]):
    print("%5d %5s %7s  %s" % (r[-1], r[idx["# Samples"]], r[idx["Instructions Executed"]], r[idx["Source"]].strip()[:100]))
